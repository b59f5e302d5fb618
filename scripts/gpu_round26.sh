#!/bin/bash
# round-2 measurement set: full GPU test suite, default bench line, ncu launch list + full capture of the conv programs, softsplat traffic
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-300 gpurun_out/r2_bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-reference --no-other-configs > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 14 -c 7 -o gpurun_out/r2_conv_tc_full -f python bench.py --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-reference --no-other-configs > gpurun_out/r2_conv_full_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:splat|scan' -s 6 -c 6 --csv --log-file gpurun_out/r2_softsplat_traffic.csv python scripts/run_splat_once.py 64 1152 1920 gentle 0 > /dev/null 2>&1
tail -8 gpurun_out/r2_softsplat_traffic.csv | cut -c1-200
