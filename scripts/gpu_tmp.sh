mkdir -p gpurun_out
timeout 600 python scripts/trace_glue.py 2>&1 | grep -v Warning > gpurun_out/trace_glue.txt; head -60 gpurun_out/trace_glue.txt
