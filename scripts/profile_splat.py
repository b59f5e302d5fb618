#!/usr/bin/env python
"""Per-kernel durations (kineto) of one softsplat call per case/variant -- analysis aid."""
import sys
import torch
sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat  # noqa: E402
from scripts.bench_splat import smooth_flow  # noqa: E402
from torch.profiler import ProfilerActivity, profile


def main():
    for (c, h, w, kind) in [(64, 1152, 1920, "gentle"), (1, 1088, 1920, "random"), (64, 544, 960, "smooth")]:
        x = torch.randn((1, c, h, w), device="cuda")
        flow = (smooth_flow(h, w, 8.0, 1) if kind == "smooth" else smooth_flow(h, w, 2.0, 1) + 6.5 if kind == "gentle"
                else 8 * torch.randn((1, 2, h, w), device="cuda"))
        metric = torch.randn((1, 1, h, w), device="cuda")
        for variant in (3, 2):
            for _ in range(2):
                softsplat(x, flow, metric, "soft", _variant=variant)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                softsplat(x, flow, metric, "soft", _variant=variant)
                torch.cuda.synchronize()
            print(f"== C={c} {h}x{w} {kind} variant={variant}")
            for e in sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start):
                print(f"   {e.device_time:9.1f} us  {e.name[:90]}")


if __name__ == "__main__":
    main()
