#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_ops.py tests/test_gpu_fused.py tests/test_gpu_union.py -x -q 2>&1 | tail -3
B="timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs"
for cfg in "1 1" "1 0" "0 1" "0 0"; do set -- $cfg; echo "LATTICE=$1 LAST3X3=$2"; DRBA_TMP_LATTICE=$1 DRBA_LAST3X3=$2 $B 2>gpurun_out/r2_bench_l$1$2.err | cut -c1-250; tail -1 gpurun_out/r2_bench_l$1$2.err | cut -c1-900; done
