#!/bin/bash
mkdir -p ab
for b in 4 5 6 8; do
  (cd drba_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_GATHER_MINB=$b -c splat_gather.cu -o build/sg_v.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libG.so $(ls build/*.o | grep -v 'splat_gather\|sg_v\|_T.o\|ifnet_fused_\|ifnet_tc_v') build/sg_v.o -lcuda && rm build/sg_v.o)
  echo "MINB=$b"; DRBA_B200_LIB=$PWD/ab/libG.so timeout 300 python scripts/bench_splat2.py 2>&1 | grep per_target | cut -c60-200
done
