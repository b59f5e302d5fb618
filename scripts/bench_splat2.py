#!/usr/bin/env python
"""softsplat C = 64 `soft` at 1152 x 1920 (SURVEY.md 8d's roofline size): per-variant time on three flow kinds
(gentle = what bench.py reports, smooth amp 8 = stronger divergence, random = worst case), CUDA events with an L2 flush
between runs.  --once runs each variant a single time (for an ncu launch list: per-kernel split of one call)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat  # noqa: E402


def smooth_flow(h, w, amp, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lo = amp * torch.randn((1, 2, h // 16, w // 16), generator=g)
    return torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False).cuda()


def main():
    once = "--once" in sys.argv
    c, h, w = 64, 1152, 1920
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    x = torch.randn((1, c, h, w), device="cuda")
    metric = torch.randn((1, 1, h, w), device="cuda")
    nbytes = h * w * 4 * ((c + 3) + c)
    for kind in ("gentle", "smooth8", "random"):
        flow = (smooth_flow(h, w, 2.0, 1) + 6.5 if kind == "gentle" else smooth_flow(h, w, 8.0, 1) if kind == "smooth8"
                else 8 * torch.randn((1, 2, h, w), device="cuda"))
        for variant in (4, 3):
            if once:
                softsplat(x, flow, metric, "soft", _variant=variant)
                torch.cuda.synchronize()
                continue
            for _ in range(3):
                softsplat(x, flow, metric, "soft", _variant=variant)
            ts = []
            for _ in range(9):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                softsplat(x, flow, metric, "soft", _variant=variant)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            print(json.dumps({"op": "softsplat", "C": c, "H": h, "W": w, "mode": "soft", "flow": kind,
                              "variant": {4: "gather_tile_tma", 3: "gather_per_target"}[variant], "ms": round(ms, 4),
                              "alg_GBps": round(nbytes / ms / 1e6, 1), "frac_of_6550": round(nbytes / ms / 1e6 / 6550.4, 3)}), flush=True)


if __name__ == "__main__":
    main()
