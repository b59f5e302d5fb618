#!/bin/bash
DRBA_E2E_EVENTS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-other-configs 2>&1 >/dev/null | grep 'e2e events'
