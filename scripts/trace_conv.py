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


def trace_program(name, trace):
    """blockN.program.xK: the whole IFBlock program (conv0a, conv0b, 8 ResConvs, lastconv) for K images at 1088 x 1920;
    one line per layer: start (previous barrier out), first TMA issued, first / last traced MMA issue, last traced drain,
    barrier in / out -- all in cycles from kernel start."""
    from drba_b200 import _lib
    from drba_b200.ifnet import IFNetEngine, _BLOCKS
    from drba_b200.weights import synth_ifnet_state
    bname, _, xk = name.split(".")
    bi, nimg = int(bname[-1]), int(xk[1:])
    eng16 = IFNetEngine(synth_ifnet_state(0), "cuda", "fp16")
    H, W = 1088, 1920
    s = [16, 8, 4, 2, 1][bi]
    _, cin, c = _BLOCKS[bi]
    h, w = H // s, W // s
    h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
    f16 = torch.float16
    xs = [torch.randn((h, w, 48 if bi == 0 else 64), device="cuda").half() for _ in range(nimg)]
    a = [torch.empty((h2, w2, c // 2), dtype=f16, device="cuda") for _ in range(nimg)]
    p0 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(nimg)]
    p1 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(nimg)]
    tch = 8 if bi == 4 else 16
    tmp = [torch.empty((h, w, tch), dtype=torch.float32, device="cuda") for _ in range(nimg)]
    steps = [(eng16.tc[f"{bname}.conv0a"], h, w, xs, a, h2, w2, c // 2, None),
             (eng16.tc[f"{bname}.conv0b"], h2, w2, a, p0, h4, w4, c, None)]
    cur, nxt = p0, p1
    for i in range(8):
        steps.append((eng16.tc[f"{bname}.res{i}"], h4, w4, cur, nxt, h4, w4, c, None if getattr(eng16.tc[f"{bname}.res{i}"], "res_tap", False) else cur))
        cur, nxt = nxt, cur
    steps.append((eng16.tc[f"{bname}.last"], h4, w4, cur, tmp, h4, w4, tch, None))
    for _ in range(3):
        eng16._conv_program(steps)
    torch.cuda.synchronize()
    trace.zero_()
    _lib.lib().drba_conv_tc_debug_trace(trace.data_ptr())
    eng16._conv_program(steps)
    torch.cuda.synchronize()
    _lib.lib().drba_conv_tc_debug_trace(None)
    t = trace.cpu().tolist()
    t0 = t[4090]
    rel = lambda v: (v - t0) if v else None
    print(f"== {name}: after setup+griddep {rel(t[4091])}, end {rel(t[4092])} cycles")
    prev_out = rel(t[4091])
    for li in range(len(steps)):
        base = li * 256
        tiles = [tl for tl in range(10) if t[base + tl * 16 + 1]]
        if not tiles:
            print(f" layer {li}: CTA 0 had no tile")
            continue
        last = tiles[-1]
        drains = [rel(v) for tl in tiles for v in (t[base + tl * 16 + 9], t[base + tl * 16 + 11]) if v]
        print(f" layer {li:2d}: start {prev_out} first_tma {rel(t[base + 1])} first_full {rel(t[base + 2])} mma_issued[0] {rel(t[base + 3])} "
              f"mma_issued[{last}] {rel(t[base + last * 16 + 3])} last_drain {max(drains) if drains else None} "
              f"barrier_in {rel(t[base + 202])} barrier_out {rel(t[base + 203])} | loop_top {rel(t[base + 208])} consts_done {rel(t[base + 205])} "
              f"tile0_prewait {rel(t[base + 0])} tile0_ring_ok {rel(t[base + 207])} |"
              f"  mma_issue_times {[rel(t[base + tl * 16 + 3]) for tl in tiles]}")
        if t[base + 203]:
            prev_out = rel(t[base + 203])


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
    for name in [n for n in names if ".program" in n]:
        trace_program(name, trace)
    for name in [n for n in names if ".program" not in n]:
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
