#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rife.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -x -q 2>&1 | tail -2
DRBA_E2E_EVENTS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>gpurun_out/e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step_median'], d['ms_per_step_max'], json.dumps(d['e2e']))"; grep 'e2e ' gpurun_out/e.err | tail -2
