#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -1
timeout 300 python scripts/bench_conv.py block3 block4 block0 2>&1 | grep x2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step_median'])"
cd drba_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_TC_TRACE=1 -c conv_tc.cu -o build/conv_tc_T.o
mkdir -p ../../ab
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libT.so $(ls build/*.o | grep -v 'conv_tc\|ifnet_fused_\|ifnet_tc_v\|sg_v') build/conv_tc_T.o -lcuda
cd ../..
DRBA_B200_LIB=$PWD/ab/libT.so timeout 200 python scripts/trace_conv.py block4.program.x2 2>&1 | sed -n 3,5p | cut -c1-330
