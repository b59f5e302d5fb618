#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_cli.py tests/test_gpu_rife.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['e2e']))"
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['e2e']))"
