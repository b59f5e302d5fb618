#!/bin/bash
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs"
run() { $B 2>/tmp/err.txt | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'])"; grep -o '"ifnet_assemble": {"ms_per_step": [0-9.]*' /tmp/err.txt; }
echo "default (9,10)"; run
mkdir -p ab
for v in "8 9" "10 12" "12 12" "10 10"; do set -- $v
  (cd drba_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_ASM_MINB4=$1 -DDRBA_ASM_MINB1=$2 -c ifnet_tc.cu -o build/ifnet_tc_v.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libV.so $(ls build/*.o | grep -v 'ifnet_tc\|_T.o\|ifnet_fused_') build/ifnet_tc_v.o -lcuda && rm build/ifnet_tc_v.o)
  echo "variant ($1,$2)"; DRBA_B200_LIB=$PWD/ab/libV.so run
done
