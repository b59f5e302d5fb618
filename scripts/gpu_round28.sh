#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gmfss.py -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_splat2.py 2>&1 | grep per_target
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], json.dumps(d['softsplat_roofline'])[:700])"
