#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2_bench_2gpu.err > gpurun_out/r2_bench_2gpu.json; cut -c1-400 gpurun_out/r2_bench_2gpu.json; tail -2 gpurun_out/r2_bench_2gpu.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-500
