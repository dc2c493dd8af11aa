#!/bin/bash
# quick ncu check of selected kernels: duration + DRAM bytes (regex in $1)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none -k regex:"$1" -c ${2:-6} --csv --log-file gpurun_out/ncu_quick.csv python scripts/profile_kernels.py --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_quick.csv")) if len(r)>10 and r[0].isdigit()]
cur=None
for r in rows:
    key=(r[0], r[4][:50])
    if key!=cur: print(); print(key, end=" "); cur=key
    print("%s=%s%s"%(r[-3].split("__")[-1][:28], r[-1], r[-2]), end=" | ")
print()
PY
