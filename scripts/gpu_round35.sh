#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; cut -c1-260 gpurun_out/r2_bench_final.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
