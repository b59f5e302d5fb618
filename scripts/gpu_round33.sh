#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:ifnet_assemble_v2 -s 24 -c 6 -o gpurun_out/r2_assemble_full -f python bench.py --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-reference --no-other-configs > gpurun_out/r2_assemble_ncu.log 2>&1
tail -2 gpurun_out/r2_assemble_ncu.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py tests/test_gpu_rife.py tests/test_gpu_union.py tests/test_gpu_gmfss.py -x -q 2>&1 | tail -2
