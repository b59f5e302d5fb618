#!/usr/bin/env python
"""CPU study for the round-2 plan: what would fp8 (e4m3) operands cost in accuracy?

Runs the fp32 oracle restatement of IFNet 4.26-heavy (oracle/ifnet.py) on a synthetic clip with the trained weights,
once as is, once with the conv operands rounded to fp16 (what the tcgen05 engine computes today: fp16 operands, fp32
accumulation) and once with the operands of the ResConv layers rounded to e4m3 (per-output-channel weight scales,
one activation scale per layer taken from the tensor's absolute maximum), and reports the PSNR of the interpolated
frame against the fp32 run.  No GPU, nothing of the product path is involved.

usage: python scripts/study_fp8_resconv.py [H W]      (default 256 448)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ifnet as O  # noqa: E402

MODE = {"conv": "fp32", "res": "fp32"}
_conv2d = F.conv2d


def _q16(x):
    return x.half().float()


def _q8_act(x):
    s = x.abs().max().clamp_min(1e-12) / 448.0          # e4m3 max normal
    return (x / s).to(torch.float8_e4m3fn).float() * s


def _q8_w(w):
    s = w.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12) / 448.0
    return (w / s).to(torch.float8_e4m3fn).float() * s


def conv2d(x, w, b=None, stride=1, padding=0, *a, **k):
    is_res = w.shape[0] == w.shape[1] and w.shape[2] == 3 and stride == 1 and w.shape[0] >= 32      # the ResConvs
    mode = MODE["res"] if is_res else MODE["conv"]
    if mode == "fp16":
        x, w = _q16(x), _q16(w)
    elif mode == "fp8":
        x, w = _q8_act(x), _q8_w(w)
    return _conv2d(x, w, b, stride, padding, *a, **k)


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()


def main():
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (256, 448)
    state, wsrc = bench.load_state()
    sd = {k: v.float() for k, v in state.items()}
    frames = bench.synth_clip(2, h, w, 3, "cpu")
    x = torch.cat(frames, 1)
    scale_list = [16, 8, 4, 2, 1]
    F.conv2d = conv2d
    out = {"weights": wsrc, "size": [h, w], "timestep": 0.5}
    runs = {}
    with torch.inference_mode():
        for name, conv_mode, res_mode in (("fp32", "fp32", "fp32"), ("fp16_all", "fp16", "fp16"), ("fp16_conv_fp8_res", "fp16", "fp8")):
            MODE["conv"], MODE["res"] = conv_mode, res_mode
            merged, flows = O.ifnet_forward(sd, x, 0.5, scale_list)
            runs[name] = (merged, flows[-1])
    ref_img, ref_flow = runs["fp32"]
    for name in ("fp16_all", "fp16_conv_fp8_res"):
        img, flow = runs[name]
        out[name] = {"psnr_vs_fp32_dB": round(psnr(img, ref_img), 2),
                     "flow_mean_abs_err_px": round(float((flow - ref_flow).abs().mean()), 4),
                     "flow_max_abs_err_px": round(float((flow - ref_flow).abs().max()), 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
