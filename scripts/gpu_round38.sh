#!/bin/bash
for m in inkernel inkernel3 ineg; do echo $m; DRBA_VALUE_DEBUG=$m timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; done
