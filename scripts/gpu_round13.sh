#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rife.py -x -q 2>&1 | tail -3
for p in 1 0; do echo "PDL=$p"; DRBA_PDL=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-other-configs 2>/dev/null | cut -c1-330; done
timeout 600 python scripts/bench_conv.py 2>&1 | grep -v '^{' | tee gpurun_out/r2_conv_layers_a.txt
