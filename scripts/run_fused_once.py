#!/usr/bin/env python
"""One launch of the fused block-input + conv0a kernel per block (3: scale 2, 4: scale 1) at 1088 x 1920 with two jobs,
for ncu captures and CUDA-event timing (prints the time of each launch; smooth synthetic flow of a few pixels)."""
import sys

import torch

sys.path.insert(0, ".")
from drba_b200.ifnet import IFNetEngine, _BLOCKS  # noqa: E402
from drba_b200.weights import synth_ifnet_state  # noqa: E402


def main():
    H, W = 1088, 1920
    eng = IFNetEngine(synth_ifnet_state(0), "cuda", "fp16")
    g = torch.Generator(device="cpu").manual_seed(0)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1

    def job(s_prev):
        lo = 3.0 * torch.randn((1, 4, H // 32, W // 32), generator=g)
        flow = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear")[0].permute(1, 2, 0).contiguous().cuda()
        return {"img0": torch.rand((1, 3, H, W), generator=g).cuda(), "img1": torch.rand((1, 3, H, W), generator=g).cuda(),
                "f0": torch.randn((H, W, 16), generator=g).half().cuda(), "f1": torch.randn((H, W, 16), generator=g).half().cuda(),
                "ts_t": torch.rand((1, 1, H, W), generator=g).cuda(), "ts_s": 0.0, "flow": flow,
                "prev": (torch.randn((H // s_prev, W // s_prev, 16), generator=g).cuda(), 1, s_prev)}

    for bi, s in ((4, 1), (3, 2)):
        name, _, c = _BLOCKS[bi]
        jobs = [job(2 * s), job(2 * s)]
        outs = [torch.empty((H // s // 2, W // s // 2, c // 2), dtype=torch.float16, device="cuda") for _ in jobs]
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng._block_conv0a(name, jobs, outs, H, W, s)
            b.record()
            torch.cuda.synchronize()
            print(f"{name} (scale {s}, 2 jobs): {a.elapsed_time(b) * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
