#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_20.json 2> gpurun_out/bench_final_20.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_final_20.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['vs_gpu_reference'], {c:v.get('value') for c,v in d['other_configs'].items()}, d['clocks'])"
