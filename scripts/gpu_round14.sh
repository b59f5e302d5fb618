#!/bin/bash
for d in 0 1 2 8 16 24 26; do echo "DBG=$d"; DRBA_TC_DBG=$d timeout 300 python scripts/bench_conv.py block3.res block4.res block3 block4 2>&1 | grep -v '^{'; done | tee gpurun_out/r2_conv_dbg_ab.txt
