#!/usr/bin/env python
"""Decode the clock64 pipeline trace of CTA 0 for one conv layer / program (debug aid).

The trace points are compiled out of the product library (they cost the producer / MMA warps 10-15 % on issue-bound
layers).  Build a tracing variant and point the loader at it:

    cd drba_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
        -Xcompiler -fPIC,-fvisibility=hidden -DDRBA_TC_TRACE=1 -c conv_tc.cu -o build/conv_tc_T.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libT.so \
        $(ls build/*.o | grep -v conv_tc) build/conv_tc_T.o -lcuda
    DRBA_B200_LIB=$PWD/../../ab/libT.so python scripts/trace_conv.py block4.res.x2 block4.conv0a
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from drba_b200 import _lib
    from drba_b200.ifnet import IFNetEngine, _tc_conv3x3
    from drba_b200.weights import synth_ifnet_state
    eng = IFNetEngine(synth_ifnet_state(0), "cuda", "fp32")
    cases = {"block0.res": (192, 192, 17, 30, 1, True), "block4.res": (32, 32, 272, 480, 1, True),
             "block4.conv0a": (64, 16, 1088, 1920, 2, False), "block2.res": (96, 96, 68, 120, 1, True),
             "block3.res": (64, 64, 136, 240, 1, True), "block4.res.x2": (32, 32, 544, 480, 1, True), "encode.cnn1": (16, 16, 544, 960, 1, True), "gridnet64": (64, 64, 544, 960, 1, True)}
    names = sys.argv[1:] or list(cases)
    trace = torch.zeros(4096, dtype=torch.int64, device="cuda")
    for name in names:
        cin, cout, h, w, stride, res = cases[name]
        g = torch.Generator(device="cpu").manual_seed(1)
        wt = torch.randn((cout, cin, 3, 3), generator=g) * (1.0 / (cin * 9)) ** 0.5
        layer = _tc_conv3x3(wt, torch.zeros((cout,)), stride, 1, "cuda")
        oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        x = (torch.randn((h, w, cin), generator=g) * 0.5).half().cuda()
        y = torch.empty((oh, ow, layer.cout_pad), dtype=torch.float16, device="cuda")
        nl = 3 if res else 1
        steps = [(layer, h, w, [x], [y], oh, ow, layer.cout_pad, [x] if res else None)]
        if res:
            steps += [(layer, h, w, [y], [x], oh, ow, layer.cout_pad, [y]), (layer, h, w, [x], [y], oh, ow, layer.cout_pad, [x])]
        for _ in range(3):
            eng._conv_program(steps)
        torch.cuda.synchronize()
        trace.zero_()
        _lib.lib().drba_conv_tc_debug_trace(trace.data_ptr())
        eng._conv_program(steps)
        torch.cuda.synchronize()
        _lib.lib().drba_conv_tc_debug_trace(None)
        t = trace.cpu().tolist()
        t0 = t[4090]
        rel = lambda v: (v - t0) if v else None
        print(f"== {name}: start 0, after setup+griddep {rel(t[4091])}, end {rel(t[4092])} (cycles; ~1.9 GHz)")
        for li in range(nl):
            base = li * 256
            print(f" layer {li}: per tile of CTA 0: K iteration 0/1 (prod_prewait, prod_issued, mma_full, mma_issued) | "
                  f"epilogue group 0 (acc_full, drained) | group 1")
            for tl in range(10):
                b = base + tl * 16
                if t[b + 1] == 0:
                    break
                print("   ", tl, [rel(t[b + k]) for k in range(4)], [rel(t[b + 4 + k]) for k in range(4)],
                      "|", [rel(t[b + 8]), rel(t[b + 9])], "|", [rel(t[b + 10]), rel(t[b + 11])])
            print(f"   barrier_in {rel(t[base + 202])} barrier_out {rel(t[base + 203])}")


if __name__ == "__main__":
    main()
