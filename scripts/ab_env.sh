#!/bin/bash
B="timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs"
for v in "DRBA_TC_ALT=0" "DRBA_TC_ALT=2" "DRBA_TC_STAGED=0" "DRBA_TC_STAGED=2" "DRBA_TC_PACK=0" "DRBA_FLOW_TERMS=0" "DRBA_LAST3X3=0" "DRBA_RES_TAP=0" "DRBA_TC_NARROW=0"; do
  echo -n "$v: "; env $v $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'])"; done
