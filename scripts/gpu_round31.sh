#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_union.py -x -q 2>&1 | tail -3
for n in 1 0; do echo "RES_TAP=$n"; DRBA_RES_TAP=$n timeout 300 python scripts/bench_conv.py block3 block4 2>&1 | grep x2; done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>gpurun_out/r2_bench_6.err | cut -c1-250; tail -1 gpurun_out/r2_bench_6.err | cut -c1-300
