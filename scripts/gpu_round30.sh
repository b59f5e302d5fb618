#!/bin/bash
for m in nodownload noupload; do echo $m; DRBA_E2E_DEBUG=$m timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['e2e']))"; done
