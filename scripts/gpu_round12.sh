#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_gmfss.py -x -q -k "conv or block_program or windows_1080p or gmfss" 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>&1 | tail -2 | cut -c1-420; done
