#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_rife.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step_median'], json.dumps(d['e2e'])[:120])"
