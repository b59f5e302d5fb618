#!/bin/bash
DRBA_E2E_EVENTS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>gpurun_out/e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['e2e']))"; grep 'e2e ' gpurun_out/e.err | tail -6
