#!/bin/bash
# The round's measurement set, run on a B200 box with `gpurun --timeout 3600 -- scripts/gpu_measure.sh`:
# full GPU test suite, smoke(), the default bench line, the ncu launch list of the bench command, one ncu full
# capture of the conv programs of a window and the DRAM traffic of one softsplat call.  Results land in gpurun_out/;
# the summaries under profiles/ are made from them (scripts/summarize_launches.py, scripts/ncu_summary.py).
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-260 gpurun_out/bench_final.json
B="python bench.py --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-reference --no-other-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 14 -c 7 -o gpurun_out/conv_tc_full -f $B > gpurun_out/conv_full_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:splat|scan' -s 6 -c 6 --csv --log-file gpurun_out/softsplat_traffic.csv python scripts/run_splat_once.py 64 1152 1920 gentle 0 > /dev/null 2>&1
