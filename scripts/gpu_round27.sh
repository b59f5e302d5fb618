#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_ops.py tests/test_gpu_union.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>gpurun_out/r2_bench_5.err | cut -c1-250; tail -1 gpurun_out/r2_bench_5.err | cut -c1-500
