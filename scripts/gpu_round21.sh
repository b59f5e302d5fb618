#!/bin/bash
for t in 1 0; do echo "ALT_TAIL=$t"; DRBA_TC_ALT_TAIL=$t timeout 300 python scripts/bench_conv.py block3 block4 2>&1 | grep -v '^{' | grep x2; done
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>/dev/null | cut -c1-330
cd drba_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_TC_TRACE=1 -c conv_tc.cu -o build/conv_tc_T.o
mkdir -p ../../ab
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libT.so $(ls build/*.o | grep -v 'conv_tc\|ifnet_fused_') build/conv_tc_T.o -lcuda
cd ../..
DRBA_B200_LIB=$PWD/ab/libT.so python scripts/trace_conv.py block4.program.x2 2>&1 | cut -c1-330 | tee gpurun_out/r2_program_trace_b.txt
