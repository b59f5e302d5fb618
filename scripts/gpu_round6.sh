#!/bin/bash
# dev helper (run through gpurun): splat tile kernel + full GPU suite + default bench
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "softsplat" 2>&1 | tail -3
timeout 300 python scripts/bench_splat2.py > gpurun_out/r2_splat2.jsonl 2>&1; cat gpurun_out/r2_splat2.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_splat2_launches.csv python scripts/bench_splat2.py --once > /dev/null 2>&1
grep "drba::" gpurun_out/r2_splat2_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100 | head -40
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-other-configs > gpurun_out/r2_bench_2.json 2> gpurun_out/r2_bench_2.err; cut -c1-600 gpurun_out/r2_bench_2.json; tail -2 gpurun_out/r2_bench_2.err | cut -c1-1200
