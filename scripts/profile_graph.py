#!/usr/bin/env python
"""Warm-cache, in-graph kernel durations of the bench workload (kineto/CUPTI activity trace).

    python scripts/profile_graph.py [--windows 8] [--out gpurun_out/graph_kernels.json]

ncu serialises launches and flushes caches, and per-launch CUDA events in eager mode measure the
host's launch rate for the ~10 us kernels; this trace gives each kernel's duration INSIDE the CUDA
graph replay that bench.py times (analysis only -- never a bench value)."""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "graph_kernels.json"))
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    from drba_b200.rife import RIFE
    dev = torch.device("cuda", 0)
    state, _ = bench.load_state()
    model = RIFE(state=state, device=dev, precision=args.precision)
    h, w = bench.net_size(bench.H_SRC, bench.W_SRC)
    frames = bench.synth_clip(8, h, w, 1000, dev)
    reuse = None
    for j in range(40):
        _, reuse = model.inference_ts_drba(frames[j % 8], frames[(j + 1) % 8], frames[(j + 2) % 8], bench.TS_PATTERN[j % 2], reuse, True)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for j in range(40, 40 + args.windows):
            _, reuse = model.inference_ts_drba(frames[j % 8], frames[(j + 1) % 8], frames[(j + 2) % 8], bench.TS_PATTERN[j % 2], reuse, True)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    t_first, t_last = None, None
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = e.name
        dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
        st = e.time_range.start
        en = e.time_range.end
        t_first = st if t_first is None else min(t_first, st)
        t_last = en if t_last is None else max(t_last, en)
        key = name.split("(")[0].replace("void ", "").replace("drba::", "")[:70]
        d = agg.setdefault(key, {"n": 0, "us": 0.0})
        d["n"] += 1
        d["us"] += dur
    tot = sum(d["us"] for d in agg.values()) or 1.0
    span = (t_last - t_first) if t_first is not None else 0.0
    out = {"windows": args.windows, "sum_kernel_us_per_window": round(tot / args.windows, 1),
           "span_us_per_window": round(span / args.windows, 1),
           "kernels": {k: {"per_window": round(v["n"] / args.windows, 1), "avg_us": round(v["us"] / v["n"], 2),
                           "us_per_window": round(v["us"] / args.windows, 1), "share": round(v["us"] / tot, 4)}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"])}}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("sum_kernel_us_per_window", "span_us_per_window")}))
    for k, v in out["kernels"].items():
        print(f"{v['share']:7.3f} {v['per_window']:7.1f} {v['avg_us']:9.2f} us  {k}")
    # every kernel of two consecutive windows in start order
    seq = [(e.time_range.start, e.name.split("(")[0].replace("void ", "").replace("drba::", "")[:48],
            e.device_time if hasattr(e, "device_time") else e.cuda_time) for e in prof.events()
           if e.device_type == torch.autograd.DeviceType.CUDA]
    seq.sort()
    nper = len(seq) // args.windows if args.windows else 0
    with open(args.out.replace(".json", "_seq.txt"), "w") as f:
        t00 = seq[0][0] if seq else 0
        for st, name, dur in seq[:2 * nper]:
            f.write(f"{(st - t00):10.1f} {dur:9.2f}  {name}\n")
    # per-launch list of the conv kernel in graph order (one window), to map layers
    conv = [(e.time_range.start, e.device_time if hasattr(e, "device_time") else e.cuda_time) for e in prof.events()
            if e.device_type == torch.autograd.DeviceType.CUDA and "conv_tc" in e.name]
    conv.sort()
    n = len(conv) // args.windows if args.windows else 0
    json.dump([round(d, 2) for _, d in conv[:2 * n]], open(args.out.replace(".json", "_conv_seq.json"), "w"))


if __name__ == "__main__":
    main()
