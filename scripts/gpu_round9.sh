#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "softsplat" 2>&1 | tail -3
timeout 300 python scripts/bench_splat2.py > gpurun_out/r2_splat2.jsonl 2>&1; cat gpurun_out/r2_splat2.jsonl
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_splat2_launches.csv python scripts/bench_splat2.py --once > /dev/null 2>&1
grep "drba::splat" gpurun_out/r2_splat2_launches.csv | awk -F'","' '{print substr($5,1,45), $(NF-2), $NF}' | head -40
