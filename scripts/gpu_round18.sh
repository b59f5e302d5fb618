#!/bin/bash
python scripts/run_fused_once.py 3
ncu --set full --import-source on --clock-control none -k regex:ifnet_fused -c 2 -o gpurun_out/r2_fused_v2 -f python scripts/run_fused_once.py 1 > gpurun_out/r2_fused_ncu.log 2>&1
tail -3 gpurun_out/r2_fused_ncu.log
