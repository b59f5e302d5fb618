#!/usr/bin/env python
"""GMFSS_UNION 4K (3840x2160 -> 2304x3840 net input, scale 0.5; BASELINE.json configs[3]) and GMFSS 1080p
(configs[2]) DRBA-window timing on B200: CUDA events over steady windows, trained weights if present."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(model, frames, windows):
    reuse = None
    ts = bench.TS_PATTERN
    for j in range(4):      # every (timestamp pattern, reuse) combination once: graph captures stay outside the timed region
        _, reuse = model.inference_ts_drba(frames[j % 4], frames[(j + 1) % 4], frames[(j + 2) % 4], ts[j % 2], reuse, True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nout = 0
    a.record()
    for j in range(4, 4 + windows):
        out, reuse = model.inference_ts_drba(frames[j % 4], frames[(j + 1) % 4], frames[(j + 2) % 4], ts[j % 2], reuse, True)
        nout += len(out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return {"windows": windows, "ms_per_window": round(ms / windows, 2), "output_frames_per_s": round(nout / ms * 1e3, 2)}


def main():
    dev = torch.device("cuda", 0)
    out = {}
    from drba_b200.gmfss import GMFSS
    from drba_b200.weights import find_gmfss_weights
    w = find_gmfss_weights()
    if w and os.path.isfile(os.path.join(w, "flownet.pkl")):
        H, W = bench.net_size(1080, 1920, 1.0, 64)
        frames = bench.synth_clip(4, H, W, 3, dev)
        out["gmfss_1080p"] = dict(net_input=[H, W], **run(GMFSS(weights=w, device=dev), frames, 6))
        del frames
    wu = os.path.join(ROOT, "baseline/_ref/weights/train_log_gmfss_union")
    if os.path.isfile(os.path.join(wu, "rife.pkl")):
        from drba_b200.gmfss_union import GMFSS_UNION
        H, W = bench.net_size(2160, 3840, 0.5, 128)
        frames = bench.synth_clip(4, H, W, 4, dev)
        out["gmfss_union_4k_scale0.5"] = dict(net_input=[H, W], **run(GMFSS_UNION(weights=wu, scale=0.5, device=dev), frames, 4))
        H, W = bench.net_size(1080, 1920, 1.0, 128)
        frames = bench.synth_clip(4, H, W, 5, dev)
        out["gmfss_union_1080p"] = dict(net_input=[H, W], **run(GMFSS_UNION(weights=wu, scale=1.0, device=dev), frames, 4))
    out["max_memory_GB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
