#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_rife.py tests/test_gpu_fullsize.py tests/test_gpu_gmfss.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>gpurun_out/r2_bench_fence.err | cut -c1-330; tail -1 gpurun_out/r2_bench_fence.err | cut -c1-300
timeout 300 python scripts/bench_conv.py block3 block4 block0 2>&1 | grep -v '^{'
