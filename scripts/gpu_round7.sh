#!/bin/bash
# dev helper (run through gpurun): splat tile kernel, conv traces, the GMFSS / union bench configurations
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "softsplat" 2>&1 | tail -3
timeout 300 python scripts/bench_splat2.py > gpurun_out/r2_splat2.jsonl 2>&1; cat gpurun_out/r2_splat2.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_splat2_launches.csv python scripts/bench_splat2.py --once > /dev/null 2>&1
grep "drba::splat_gather" gpurun_out/r2_splat2_launches.csv | awk -F'","' '{print substr($5,1,45), $NF}'
for alt in 1 0; do echo "== trace DRBA_TC_ALT=$alt"; DRBA_TC_ALT=$alt DRBA_B200_LIB=$PWD/drba_b200/csrc/build/libT.so timeout 300 python scripts/trace_conv.py block4.res.x2 block3.res > gpurun_out/r2_conv_trace_alt$alt.txt 2>&1; head -45 gpurun_out/r2_conv_trace_alt$alt.txt; done
timeout 1200 python bench.py --config gmfss1080_scdet > gpurun_out/r2_bench_gmfss.json 2> gpurun_out/r2_bench_gmfss.err; cut -c1-1500 gpurun_out/r2_bench_gmfss.json; tail -3 gpurun_out/r2_bench_gmfss.err | cut -c1-1500
timeout 1200 python bench.py --config union4k > gpurun_out/r2_bench_union.json 2> gpurun_out/r2_bench_union.err; cut -c1-1500 gpurun_out/r2_bench_union.json; tail -3 gpurun_out/r2_bench_union.err | cut -c1-1500
