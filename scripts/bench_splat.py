#!/usr/bin/env python
"""Micro-benchmark of the softsplat / DRM kernels (CUDA events, L2 flushed between runs).
Prints one JSON line per case with achieved GB/s on ALGORITHMIC bytes:
    splat_bytes = N*H*W*4*[(C + 2 + (metric?1:0)) + C];  drm_rife_bytes = N*H*W*(16 + 4*maps)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from drba_b200 import drm, ops  # noqa: E402
from drba_b200.softsplat import softsplat  # noqa: E402


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def smooth_flow(h, w, amp, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lo = amp * torch.randn((1, 2, h // 16, w // 16), generator=g)
    return torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False).cuda()


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (c, h, w, mode, kind) in [(1, 1088, 1920, "avg", "smooth"), (2, 1088, 1920, "avg", "smooth"),
                                  (1, 1088, 1920, "avg", "random"), (3, 544, 960, "soft", "smooth"),
                                  (64, 544, 960, "soft", "smooth"), (64, 544, 960, "soft", "random"),
                                  (128, 272, 480, "soft", "smooth"), (64, 1152, 1920, "soft", "smooth"),
                                  (64, 1152, 1920, "soft", "gentle"), (192, 288, 480, "soft", "gentle")]:
        x = torch.randn((1, c, h, w), device="cuda")
        flow = (smooth_flow(h, w, 8.0, 1) if kind == "smooth" else smooth_flow(h, w, 2.0, 1) + 6.5 if kind == "gentle"
                else 8 * torch.randn((1, 2, h, w), device="cuda"))
        metric = torch.randn((1, 1, h, w), device="cuda") if mode == "soft" else None
        nbytes = h * w * 4 * ((c + 2 + (1 if metric is not None else 0)) + c)
        for variant in (3, 2, 1):
            ms = timeit(lambda: softsplat(x, flow, metric, mode, _variant=variant), flush=flush)
            print(json.dumps({"op": "softsplat", "C": c, "H": h, "W": w, "mode": mode, "flow": kind,
                              "variant": {3: "gather", 2: "agg_v4", 1: "scalar_atomics"}[variant], "ms": round(ms, 4),
                              "alg_GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
    h, w = 1088, 1920
    f10, f12 = smooth_flow(h, w, 8.0, 2), smooth_flow(h, w, 8.0, 3)
    ms = timeit(lambda: drm.calc_drm_rife(0.4, f10, f12, True, only="drm_t1_t01"), flush=flush)
    print(json.dumps({"op": "drm_rife(one map)", "H": h, "W": w, "ms": round(ms, 4),
                      "alg_GBps": round(h * w * 20 / ms / 1e6, 1)}), flush=True)
    ms = timeit(lambda: drm.calc_drm_rife(0.4, f10, f12, True), flush=flush)
    print(json.dumps({"op": "drm_rife(two maps)", "H": h, "W": w, "ms": round(ms, 4),
                      "alg_GBps": round(h * w * 24 / ms / 1e6, 1)}), flush=True)
    ms = timeit(lambda: ops.rife_invert_flow(f10), flush=flush)
    print(json.dumps({"op": "rife_invert_flow", "H": h, "W": w, "ms": round(ms, 4),
                      "alg_GBps": round(h * w * 16 / ms / 1e6, 1)}), flush=True)
    x = torch.randn((1, 16, h, w), device="cuda")
    ms = timeit(lambda: ops.backwarp(x, f10), flush=flush)
    print(json.dumps({"op": "backwarp C=16", "H": h, "W": w, "ms": round(ms, 4),
                      "alg_GBps": round(h * w * 4 * (2 * 16 + 2) / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
