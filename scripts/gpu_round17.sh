#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs"
echo "PROD=16"; $B 2>gpurun_out/r2_bench_f16.err | cut -c1-330; tail -1 gpurun_out/r2_bench_f16.err | cut -c1-420
mkdir -p ab
for n in 12 20; do
  (cd drba_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_FU_PROD=$n -c ifnet_fused.cu -o build/ifnet_fused_$n.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libF$n.so $(ls build/*.o | grep -v 'ifnet_fused\|_T.o') build/ifnet_fused_$n.o -lcuda)
  echo "PROD=$n"; DRBA_B200_LIB=$PWD/ab/libF$n.so $B 2>gpurun_out/r2_bench_f$n.err | cut -c1-330; tail -1 gpurun_out/r2_bench_f$n.err | cut -c1-420
done
