timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -s -k "single_pass" 2>&1 | grep -E "single|passed|failed|Error|error|assert" | cut -c1-230 | tail -8
