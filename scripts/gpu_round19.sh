#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fused.py tests/test_gpu_rife.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>gpurun_out/r2_bench_v3.err | cut -c1-330; tail -1 gpurun_out/r2_bench_v3.err | cut -c1-600
