#!/bin/bash
# dev helper: microbenchmark + splat tile kernel checks + short bench (run through gpurun)
cd scripts && (timeout 120 ./build/mma_shapes > ../gpurun_out/r2_mma_shapes.jsonl 2> ../gpurun_out/r2_mma_shapes.err; echo mma rc=$?; tail -17 ../gpurun_out/r2_mma_shapes.jsonl | cut -c1-60,180-330; cat ../gpurun_out/r2_mma_shapes.err); cd ..
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q -k "softsplat or address or same_key" 2>&1 | tail -5
timeout 300 python scripts/bench_splat2.py > gpurun_out/r2_splat2.jsonl 2>&1; cat gpurun_out/r2_splat2.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_splat2_launches.csv python scripts/bench_splat2.py --once > /dev/null 2>&1
grep "drba::" gpurun_out/r2_splat2_launches.csv | sed "s/(const.*gpu__time_duration.sum\",\"ns\"//" | cut -c1-120 | head -40
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>&1 | tail -3 | cut -c1-1500
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py -x -q -k "conv or block_program or windows_1080p" 2>&1 | tail -4
DRBA_TC_ALT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>/dev/null | cut -c1-400
