#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_conv_tc.py tests/test_gpu_union.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>&1 | tail -2 | cut -c1-1400
DRBA_FLOW_TERMS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>/dev/null | cut -c1-330
