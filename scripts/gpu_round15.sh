#!/bin/bash
set -e
cd drba_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_TC_TRACE=1 -c conv_tc.cu -o build/conv_tc_T.o
mkdir -p ../../ab
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libT.so $(ls build/*.o | grep -v 'conv_tc\|ifnet_fused_\|ifnet_tc_v\|sg_v') build/conv_tc_T.o -lcuda
cd ../..
DRBA_B200_LIB=$PWD/ab/libT.so python scripts/trace_conv.py block4.program.x2 block0.program.x2 2>&1 | tee gpurun_out/r2_program_trace.txt
