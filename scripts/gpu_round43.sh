#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rife.py -x -q 2>&1 | tail -2
for c in 1 0; do echo "CHAIN=$c"; DRBA_TC_CHAIN=$c timeout 200 python scripts/bench_conv.py block3 block4 block2 block0 2>&1 | grep x2
DRBA_TC_CHAIN=$c timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step_median'])"; done
