#!/usr/bin/env python
"""Per-layer timing of the tcgen05 conv engine: R back-to-back launches of one layer captured in a
CUDA graph (no host launch gaps), warm caches.  Experiment knobs come from the environment
(DRBA_TC_STAGES, DRBA_TC_PDL, DRBA_TC_DBG), see csrc/conv_tc.cu."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [  # name, cin, cout, h, w, stride, res
    ("block0.res", 192, 192, 17, 30, 1, True),
    ("block1.res", 128, 128, 34, 60, 1, True),
    ("block2.res", 96, 96, 68, 120, 1, True),
    ("block3.res", 64, 64, 136, 240, 1, True),
    ("block4.res", 32, 32, 272, 480, 1, True),
    ("block4.conv0a", 64, 16, 1088, 1920, 2, False),
    ("block4.conv0b", 16, 32, 544, 960, 2, False),
    ("encode.cnn1", 16, 16, 544, 960, 1, False),
    ("gridnet64", 64, 64, 544, 960, 1, True),
    ("gridnet128", 128, 128, 272, 480, 1, True),
]


def main():
    from drba_b200.ifnet import IFNetEngine, _tc_conv3x3
    from drba_b200.weights import synth_ifnet_state
    eng = IFNetEngine(synth_ifnet_state(0), "cuda", "fp32")
    R = 40
    only = sys.argv[1:] or None
    out = {}
    for name, cin, cout, h, w, stride, res in LAYERS:
        if only and name not in only:
            continue
        g = torch.Generator(device="cpu").manual_seed(1)
        wt = torch.randn((cout, cin, 3, 3), generator=g) * (1.0 / (cin * 9)) ** 0.5
        b = torch.zeros((cout,))
        layer = _tc_conv3x3(wt, b, stride, 1, "cuda")
        oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        x = (torch.randn((h, w, cin), generator=g) * 0.5).half().cuda()
        if res:
            bufs = [x, torch.empty_like(x)]
        else:
            bufs = [x, torch.empty((oh, ow, layer.cout_pad), dtype=torch.float16, device="cuda")]

        def run():
            for i in range(R):
                if res:
                    a, o = bufs[i % 2], bufs[(i + 1) % 2]
                    eng._conv_tc(layer, a, h, w, o, oh, ow, layer.cout_pad, res=a)
                else:
                    eng._conv_tc(layer, bufs[0], h, w, bufs[1], oh, ow, layer.cout_pad)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            run()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * R)
        flops = 2.0 * 9 * cin * cout * oh * ow
        out[name] = {"us": round(us, 2), "TFLOPs": round(flops / us / 1e6, 1)}
        print(name, out[name], flush=True)
    # whole blocks as one persistent program (what the engine launches)
    from drba_b200.ifnet import _BLOCKS
    H, W = 1088, 1920
    eng16 = IFNetEngine(synth_ifnet_state(0), "cuda", "fp16")
    for nimg in (1, 2):
        for bi, s in enumerate([16, 8, 4, 2, 1]):
            name, cin, c = _BLOCKS[bi]
            if only and name not in only:
                continue
            h, w = H // s, W // s
            h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
            f16 = torch.float16
            xs = [torch.randn((h, w, 48 if bi == 0 else 64), device="cuda").half() for _ in range(nimg)]
            a = [torch.empty((h2, w2, c // 2), dtype=f16, device="cuda") for _ in range(nimg)]
            p0 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(nimg)]
            p1 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(nimg)]
            tmp = [torch.empty((h, w, 16), dtype=torch.float32, device="cuda") for _ in range(nimg)]
            steps = [(eng16.tc[f"{name}.conv0a"], h, w, xs, a, h2, w2, c // 2, None),
                     (eng16.tc[f"{name}.conv0b"], h2, w2, a, p0, h4, w4, c, None)]
            cur, nxt = p0, p1
            for i in range(8):
                steps.append((eng16.tc[f"{name}.res{i}"], h4, w4, cur, nxt, h4, w4, c, None if getattr(eng16.tc[f"{name}.res{i}"], "res_tap", False) else cur))
                cur, nxt = nxt, cur
            steps.append((eng16.tc[f"{name}.last"], h4, w4, cur, tmp, h4, w4, 16, None))
            flops = sum(nimg * 2.0 * getattr(l, "flop_taps", l.G * l.T) * l.cin_real * l.cout * oh * ow for (l, _, _, _, _, oh, ow, _, _) in steps)
            s_ = torch.cuda.Stream()
            with torch.cuda.stream(s_):
                eng16._conv_program(steps)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s_):
                for _ in range(5):
                    eng16._conv_program(steps)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            out[f"{name}.program.x{nimg}"] = {"us": round(us, 2), "TFLOPs": round(flops / us / 1e6, 1)}
            print(f"{name}.program.x{nimg}", out[f"{name}.program.x{nimg}"], flush=True)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("DRBA_TC")}, "layers": out}))


if __name__ == "__main__":
    main()
