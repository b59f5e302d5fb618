#!/bin/bash
# A/B timing of library variants on one box: scripts/ab_variants.sh A D ...  (ab/lib<X>.so, same C ABI)
for v in "$@"; do
  echo "== lib$v"; DRBA_B200_LIB=/root/repo/ab/lib$v.so python scripts/bench_conv.py 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read())['layers'];print({k:v['us'] for k,v in d.items()})"
  DRBA_B200_LIB=/root/repo/ab/lib$v.so python bench.py --no-cpu-baseline --steps 200 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('bench',d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"
done
