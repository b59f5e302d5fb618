#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -5
for f in 1 0; do echo "FUSED=$f"; DRBA_FUSED_CONV0A=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>gpurun_out/r2_bench_fused$f.err | cut -c1-330; cut -c1-1500 gpurun_out/r2_bench_fused$f.err | tail -2; done
