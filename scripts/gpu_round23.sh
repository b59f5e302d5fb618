#!/bin/bash
timeout 300 python scripts/bench_conv.py block0.res block1.res block2.res gridnet64 gridnet128 block0 block1 block2 2>&1 | grep -v '^{'
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_gmfss.py tests/test_gpu_union.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench_3.err > gpurun_out/r2_bench_3.json; cut -c1-330 gpurun_out/r2_bench_3.json; tail -1 gpurun_out/r2_bench_3.err | cut -c1-300
