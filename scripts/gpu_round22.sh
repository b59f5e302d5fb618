#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_gmfss.py tests/test_gpu_gmflow.py tests/test_gpu_union.py -x -q 2>&1 | tail -4
for n in 1 0; do echo "NARROW=$n"; DRBA_TC_NARROW=$n timeout 300 python scripts/bench_conv.py block0.res block1.res block2.res block0 block1 block2 gridnet128 2>&1 | grep -v '^{'; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>gpurun_out/r2_bench_narrow.err | cut -c1-330; tail -1 gpurun_out/r2_bench_narrow.err | cut -c1-300
