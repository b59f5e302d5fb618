#!/usr/bin/env python
"""GMFSS 1080p component timing on B200 (BASELINE.json configs[2], without GMFlow which is not built yet):
FeatureNet (2 frames), MetricNet, Model.inference (splats + GridNet) per interpolated frame; CUDA events,
median of N runs, smooth synthetic flows as the injected flow estimator."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    from drba_b200 import _lib
    from drba_b200.gmfss import GMFSS
    from drba_b200.weights import find_gmfss_weights, load_gmfss_state, synth_gmfss_state
    dev = torch.device("cuda", 0)
    w = find_gmfss_weights()
    state = load_gmfss_state(w) if w else synth_gmfss_state(0)
    if "flownet" not in state:
        from drba_b200.weights import find_gmflow_weights, load_gmflow_state
        wf = find_gmflow_weights()
        if wf:
            state["flownet"] = load_gmflow_state(wf)
    H, W = bench.net_size(1080, 1920)
    frames = bench.synth_clip(3, H, W, 7, dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    lo = 4.0 * torch.randn((1, 2, H // 32, W // 32), generator=g)
    flow = torch.nn.functional.interpolate(lo, size=(H // 2, W // 2), mode="bilinear", align_corners=False).to(dev)
    m = GMFSS(state=state, device=dev, flow_estimator=lambda a, b: flow)
    out = {"weights": "trained" if w else "synthetic", "net_input": [H, W]}
    if "flownet" in state:
        from drba_b200.gmflow import GMFlow
        from drba_b200.ops import resize_bilinear as _rb
        gm = GMFlow(state["flownet"], dev)
        a_h, b_h = _rb(frames[0], scale_factor=0.5), _rb(frames[1], scale_factor=0.5)
        out["gmflow_one_direction_ms"] = round(timeit(lambda: gm(a_h, b_h), n=5), 3)
        with _lib.LaunchProfiler() as prof:
            gm(a_h, b_h)
            fam = prof.summary()
        conv = [v for k, v in fam.items() if k.startswith("conv_tc")]
        out["gmflow_conv_breakdown_ms"] = {k: round(v["ms"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        out["gmflow_engine_TFLOPs"] = round(sum(v["flops"] for v in conv) / sum(v["ms"] for v in conv) / 1e9, 1)
        out["gmflow_engine_GFLOP"] = round(sum(v["flops"] for v in conv) / 1e9, 1)
    out["featurenet_2frames_ms"] = round(timeit(lambda: m.model.feat_ext([frames[0], frames[1]])), 3)
    r = m.model.reuse(frames[1], frames[0], 1.0)
    from drba_b200.ops import resize_bilinear
    i0h, i1h = resize_bilinear(frames[0], scale_factor=0.5), resize_bilinear(frames[1], scale_factor=0.5)
    out["metricnet_ms"] = round(timeit(lambda: m.model.metricnet(i1h, i0h, r[0], r[1])), 3)
    out["reuse_without_gmflow_ms"] = round(timeit(lambda: m.model.reuse(frames[1], frames[0], 1.0)), 3)
    t0 = torch.full((1, 1, H // 2, W // 2), 0.4, device=dev)
    t1 = torch.full((1, 1, H // 2, W // 2), 0.6, device=dev)
    out["inference_per_frame_ms"] = round(timeit(lambda: m.model.inference(frames[1], frames[0], r, t0, t1)), 3)
    with _lib.LaunchProfiler() as prof:
        m.model.inference(frames[1], frames[0], r, t0, t1)
        fam = prof.summary()
    out["inference_breakdown_ms"] = {k: round(v["ms"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    conv = [v for k, v in fam.items() if k.startswith("conv_tc")]
    fl, ms = sum(v["flops"] for v in conv), sum(v["ms"] for v in conv)
    out["gridnet_TFLOPs"] = round(fl / ms / 1e9, 1)
    out["gridnet_GFLOP"] = round(fl / 1e9, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
