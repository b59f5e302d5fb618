#!/usr/bin/env python
"""Generate golden vectors by running the reference itself (CPU, fp32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  The reference has no tests or
golden vectors of its own, so these outputs of the
reference's code are what pins the oracle.  The reference's CuPy splat cannot be
imported here (no cupy); its CPU twin models/softsplat/softsplat_torch.py is
executed on fp32 inputs, which is the CuPy path's arithmetic (softsplat.py:251
upcasts everything to fp32 before the kernel).
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DRBA_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def gen_ops():
    from models.softsplat.softsplat_torch import softsplat as ref_splat
    from models import drm as ref_drm
    from models.rife_426_heavy.warplayer import warp as ref_backwarp
    import torch.nn.functional as F

    torch.set_grad_enabled(False)
    rng = np.random.default_rng(1234)
    out = {}
    cases = []
    # (name, C, H, W, flow sigma)
    for name, C, H, W, sig in [("a", 1, 24, 40, 3.0), ("b", 3, 17, 23, 6.0), ("c", 2, 32, 32, 0.7), ("d", 5, 9, 50, 12.0)]:
        x = rng.standard_normal((1, C, H, W)).astype(np.float32)
        flow = (rng.standard_normal((1, 2, H, W)) * sig).astype(np.float32)
        if name == "b":  # non-finite targets and exact-integer targets
            flow[0, 0, 3, 4] = np.inf
            flow[0, 1, 5, 6] = np.nan
            flow[0, :, 7, 7] = 2.0
            flow[0, :, 0, 0] = -1.0
        metric = rng.standard_normal((1, 1, H, W)).astype(np.float32)
        out[f"splat_{name}_in"], out[f"splat_{name}_flow"], out[f"splat_{name}_metric"] = x, flow, metric
        for mode in ["sum", "avg", "linear", "soft", "avg-zeroeps", "soft-clipeps", "linear-addeps"]:
            m = None if mode.split("-")[0] in ("sum", "avg") else _t(metric)
            with torch.inference_mode():
                y = ref_splat(_t(x), _t(flow), m, mode)
            out[f"splat_{name}_{mode}"] = y.numpy()
        cases.append(name)
    out["splat_cases"] = np.array(cases)

    # DRM: smooth-ish flows so that holes and filled regions both occur
    drm_cases = []
    for name, H, W, sig in [("p", 32, 56, 4.0), ("q", 20, 28, 1.0)]:
        f10 = (rng.standard_normal((1, 2, H, W)) * sig).astype(np.float32)
        f12 = (rng.standard_normal((1, 2, H, W)) * sig).astype(np.float32)
        # a block of large displacement creates real holes
        f10[:, :, : H // 3, : W // 3] += 9.0
        f12[:, :, H // 2:, W // 2:] -= 7.0
        m10 = rng.standard_normal((1, 1, H, W)).astype(np.float32)
        m12 = rng.standard_normal((1, 1, H, W)).astype(np.float32)
        out[f"drm_{name}_f10"], out[f"drm_{name}_f12"] = f10, f12
        out[f"drm_{name}_m10"], out[f"drm_{name}_m12"] = m10, m12
        for t in [0.2, 0.4, 0.5]:
            for linear in [True, False]:
                tag = f"drm_{name}_t{t}_{'lin' if linear else 'nl'}"
                with torch.inference_mode():
                    r = ref_drm.calc_drm_rife(t, _t(f10), _t(f12), linear)
                    g = ref_drm.calc_drm_gmfss(t, _t(f10), _t(f12), _t(m10), _t(m12), linear)
                    a = ref_drm.calc_drm_rife_auxiliary(t, _t(f10), _t(f12), _t(m10), _t(m12), linear)
                    g0 = ref_drm.calc_drm_gmfss(t, _t(f10), _t(f12), None, None, linear)
                for k, v in r.items():
                    out[f"{tag}_rife_{k}"] = v.numpy()
                for k, v in g.items():
                    out[f"{tag}_gmfss_{k}"] = v.numpy()
                for k, v in g0.items():
                    out[f"{tag}_gmfssavg_{k}"] = v.numpy()
                for k, v in a.items():
                    out[f"{tag}_aux_{k}"] = v.numpy()
        drm_cases.append(name)
    out["drm_cases"] = np.array(drm_cases)
    # the 0/0 = NaN quirk of calc_drm_gmfss
    f10 = (rng.standard_normal((1, 2, 12, 16)) * 2).astype(np.float32)
    f12 = (rng.standard_normal((1, 2, 12, 16)) * 2).astype(np.float32)
    f10[:, :, 4:6, 5:8] = 0.0
    f12[:, :, 4:6, 5:8] = 0.0
    m = rng.standard_normal((1, 1, 12, 16)).astype(np.float32)
    with torch.inference_mode():
        g = ref_drm.calc_drm_gmfss(0.4, _t(f10), _t(f12), _t(m), _t(m), True)
    out["drm_nan_f10"], out["drm_nan_f12"], out["drm_nan_m"] = f10, f12, m
    for k, v in g.items():
        out[f"drm_nan_{k}"] = v.numpy()

    # get_drm_t
    d = rng.uniform(0.01, 0.99, (1, 1, 8, 12)).astype(np.float32)
    out["gdt_in"] = d
    for t in [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.8, 0.95]:
        out[f"gdt_t{t}"] = ref_drm.get_drm_t(_t(d), t).numpy()

    # backward warp (border) and resize
    x = rng.standard_normal((1, 4, 20, 30)).astype(np.float32)
    fl = (rng.standard_normal((1, 2, 20, 30)) * 5).astype(np.float32)
    out["bw_in"], out["bw_flow"] = x, fl
    out["bw_border"] = ref_backwarp(_t(x), _t(fl)).numpy()
    # zeros padding variant: models/model_gmfss/MetricNet.py:10-20
    from models.model_gmfss.MetricNet import backwarp as ref_bw_zeros
    out["bw_zeros"] = ref_bw_zeros(_t(x), _t(fl)).numpy()
    x = rng.standard_normal((1, 3, 32, 48)).astype(np.float32)
    out["rs_in"] = x
    for sf in [0.0625, 0.125, 0.25, 0.5, 2.0, 4.0]:
        out[f"rs_sf{sf}"] = F.interpolate(_t(x), scale_factor=sf, mode="bilinear", align_corners=False).numpy()
    out["rs_size_27_41"] = F.interpolate(_t(x), size=(27, 41), mode="bilinear", align_corners=False).numpy()
    out["rs_ac_27_41"] = F.interpolate(_t(x), size=(27, 41), mode="bilinear", align_corners=True).numpy()
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)
    print("ops_golden.npz:", len(out), "arrays")


def gen_rife():
    """RIFE wrapper golden: reference classes with (a) seeded synthetic weights that
    drba_b200.weights.synth_ifnet_state reproduces anywhere, (b) the real weights."""
    import models.rife as ref_rife
    from models.rife_426_heavy.IFNet_HDv3 import IFNet
    from drba_b200.weights import synth_ifnet_state, load_ifnet_state

    torch.set_grad_enabled(False)
    # importing the reference's torch splat flips this to "medium" process-wide
    # (softsplat_torch.py:13), which lets oneDNN run fp32 convs in bf16; undo it
    torch.set_float32_matmul_precision("highest")
    rng = np.random.default_rng(77)
    H, W = 64, 128

    def smooth(shape, k=9):
        a = rng.uniform(0, 1, shape).astype(np.float32)
        t = torch.from_numpy(a)
        t = torch.nn.functional.avg_pool2d(t, k, 1, k // 2, count_include_pad=False)
        return t

    base = smooth((1, 3, H + 16, W + 16))
    I0 = base[:, :, 8:8 + H, 8:8 + W].contiguous()
    I1 = base[:, :, 6:6 + H, 11:11 + W].contiguous()
    I2 = base[:, :, 5:5 + H, 15:15 + W].contiguous()
    out = {"I0": I0.numpy(), "I1": I1.numpy(), "I2": I2.numpy()}
    for tag, state in [("synth", synth_ifnet_state(0)),
                       ("real", load_ifnet_state(os.path.join(REF, "weights/train_log_rife_426_heavy")))]:
        m = ref_rife.RIFE.__new__(ref_rife.RIFE)
        m.ifnet = IFNet().eval()
        m.ifnet.load_state_dict(state, strict=True)
        m.scale = 1.0
        m.scale_list = [16, 8, 4, 2, 1]
        m.pad_size = 64
        # fp32 oracle mode: call the undecorated bodies (no autocast), fp32 splat
        with torch.inference_mode():
            fa, fb, f0, f1 = m.calc_flow(I1, I0)
            out[f"{tag}_flow10"], out[f"{tag}_flow01"] = fa.numpy(), fb.numpy()
            out[f"{tag}_f1"] = f0.numpy()
            y = m.ifnet(torch.cat((I0, I1), 1), timestep=0.4, scale_list=m.scale_list)[0]
            out[f"{tag}_ts0.4"] = y.numpy()
            body = ref_rife.RIFE.inference_ts_drba.__wrapped__.__wrapped__
            ts = np.array([0.6, 1.0, 1.4])
            o1, reuse = body(m, I0, I1, I2, ts, None, True)
            out[f"{tag}_w0_0.6"], out[f"{tag}_w0_1.4"] = o1[0].numpy(), o1[2].numpy()
            ts = np.array([0.8, 1.2])
            o2, reuse2 = body(m, I1, I2, I0, ts, reuse, True)
            out[f"{tag}_w1_0.8"], out[f"{tag}_w1_1.2"] = o2[0].numpy(), o2[1].numpy()
            o3, _ = body(m, I0, I1, I2, np.array([0.7]), None, False)
            out[f"{tag}_w0_0.7_nl"] = o3[0].numpy()
    np.savez_compressed(os.path.join(HERE, "rife_golden.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    print("rife_golden.npz:", len(out), "arrays")


def gen_gmfss():
    """GMFSS (models/gmfss.py, models/model_gmfss/*) in fp32 on CPU: per-net outputs and a DRBA window, with the
    optical flows of the reference's GMFlow stored alongside (the native GMFlow is not built yet; the GPU tests
    inject these flows).  Two weight sets: seeded synthetic FeatureNet / MetricNet / GridNet weights (reproducible
    on the GPU box without checkpoints) and the trained ones; GMFlow always runs with its trained weights."""
    import torch.nn.functional as F
    from models import gmfss as ref_gmfss
    from models.model_gmfss.GMFSS import Model
    from drba_b200.weights import load_gmfss_state, synth_gmfss_state

    torch.set_grad_enabled(False)
    torch.set_float32_matmul_precision("highest")
    g = torch.Generator().manual_seed(11)
    H, W = 128, 192

    def smooth(shape):
        lo = torch.rand((shape[0], shape[1], shape[2] // 8 + 2, shape[3] // 8 + 2), generator=g)
        hi = F.interpolate(lo, scale_factor=8, mode="bicubic", align_corners=False)[:, :, 4:4 + shape[2], 4:4 + shape[3]]
        return (hi + 0.05 * torch.rand(shape, generator=g)).clamp(0, 1)

    base = smooth((1, 3, H + 16, W + 16))
    I0 = base[:, :, 8:8 + H, 8:8 + W].contiguous()
    I1 = base[:, :, 7:7 + H, 10:10 + W].contiguous()
    I2 = base[:, :, 5:5 + H, 13:13 + W].contiguous()
    out = {"I0": I0.numpy(), "I1": I1.numpy(), "I2": I2.numpy()}
    wdir = os.path.join(REF, "weights/train_log_gmfss")
    for tag, state in [("synth", synth_gmfss_state(0)), ("real", load_gmfss_state(wdir))]:
        m = ref_gmfss.GMFSS.__new__(ref_gmfss.GMFSS)
        m.model = Model()
        m.model.load_model(wdir, -1)
        m.model.feat_ext.load_state_dict(state["feat"], strict=True)
        m.model.metricnet.load_state_dict(state["metric"], strict=True)
        m.model.fusionnet.load_state_dict(state["fusionnet"], strict=True)
        m.model.eval()
        m.scale = 1.0
        m.pad_size = 64
        with torch.inference_mode():
            r10 = m.model.reuse(I1, I0, 1.0)      # (flow10, flow01, metric10, metric01, feats(I1), feats(I0))
            r12 = m.model.reuse(I1, I2, 1.0)
            if tag == "synth":                     # flows depend on GMFlow only
                out["flow10"], out["flow01"] = r10[0].numpy(), r10[1].numpy()
                out["flow12"], out["flow21"] = r12[0].numpy(), r12[1].numpy()
            out[f"{tag}_metric10"], out[f"{tag}_metric01"] = r10[2].numpy(), r10[3].numpy()
            f1, f2, f3 = r10[4]
            out[f"{tag}_feat1_sub"] = f1[:, ::4].numpy().astype(np.float16)
            out[f"{tag}_feat2_sub"] = f2[:, ::8].numpy().astype(np.float16)
            out[f"{tag}_feat3"] = f3.numpy().astype(np.float16)
            y = m.model.inference(I1, I0, r10, timestep0=0.3, timestep1=0.7)
            out[f"{tag}_inf_scalar"] = y.numpy().astype(np.float16)
            body = ref_gmfss.GMFSS.inference_ts_drba.__wrapped__.__wrapped__
            o1, reuse = body(m, I0, I1, I2, np.array([0.6, 1.0, 1.4]), None, True)
            out[f"{tag}_w0_0.6"], out[f"{tag}_w0_1.4"] = o1[0].numpy().astype(np.float16), o1[2].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "gmfss_golden.npz"), **out)
    print("gmfss_golden.npz:", len(out), "arrays")


def gen_union():
    """GMFSS_UNION (models/gmfss_union.py, models/model_gmfss_union/*) in fp32 on CPU: one DRBA window per weight set
    (seeded synthetic GMFSS nets + synthetic RIFE, and the trained checkpoints), with the GMFlow flows stored."""
    import torch.nn.functional as F
    from models import gmfss_union as ref_union
    from models.model_gmfss_union.GMFSS import Model
    from models.rife_426_heavy.IFNet_HDv3 import IFNet
    from drba_b200.gmfss_union import load_union_rife_state
    from drba_b200.weights import load_gmfss_state, synth_gmfss_state, synth_ifnet_state

    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(21)
    H, W = 128, 256

    def smooth(shape):
        lo = torch.rand((shape[0], shape[1], shape[2] // 8 + 2, shape[3] // 8 + 2), generator=g)
        hi = F.interpolate(lo, scale_factor=8, mode="bicubic", align_corners=False)[:, :, 4:4 + shape[2], 4:4 + shape[3]]
        return (hi + 0.05 * torch.rand(shape, generator=g)).clamp(0, 1)

    base = smooth((1, 3, H + 16, W + 16))
    I0 = base[:, :, 8:8 + H, 8:8 + W].contiguous()
    I1 = base[:, :, 7:7 + H, 10:10 + W].contiguous()
    I2 = base[:, :, 5:5 + H, 13:13 + W].contiguous()
    out = {"I0": I0.numpy(), "I1": I1.numpy(), "I2": I2.numpy()}
    wdir = os.path.join(REF, "weights/train_log_gmfss_union")
    for tag, state, rife_state in [("synth", synth_gmfss_state(0, union=True), synth_ifnet_state(0)),
                                   ("real", load_gmfss_state(wdir), load_union_rife_state(wdir))]:
        m = ref_union.GMFSS_UNION.__new__(ref_union.GMFSS_UNION)
        m.model = Model()
        _load = torch.load      # the union loader has no map_location (model_gmfss_union/GMFSS.py:50-53)
        torch.load = lambda f, *a, **k: _load(f, *a, **{**k, "map_location": "cpu"})
        try:
            m.model.load_model(wdir, -1)
        finally:
            torch.load = _load
        m.model.feat_ext.load_state_dict(state["feat"], strict=True)
        m.model.metricnet.load_state_dict(state["metric"], strict=True)
        m.model.fusionnet.load_state_dict(state["fusionnet"], strict=True)
        m.model.eval()
        m.ifnet = IFNet().eval()
        m.ifnet.load_state_dict(rife_state, strict=False)
        m.scale = 1.0
        m.scale_list = [16, 8, 4, 2, 1]
        m.pad_size = 128
        with torch.inference_mode():
            if tag == "synth":
                r10 = m.model.reuse(I1, I0, 1.0)
                r12 = m.model.reuse(I1, I2, 1.0)
                out["flow10"], out["flow01"] = r10[0].numpy(), r10[1].numpy()
                out["flow12"], out["flow21"] = r12[0].numpy(), r12[1].numpy()
            body = ref_union.GMFSS_UNION.inference_ts_drba.__wrapped__.__wrapped__
            o1, reuse = body(m, I0, I1, I2, np.array([0.6, 1.0, 1.4]), None, True)
            out[f"{tag}_w0_0.6"], out[f"{tag}_w0_1.4"] = o1[0].numpy().astype(np.float16), o1[2].numpy().astype(np.float16)
            out[f"{tag}_metric12"] = reuse[3].numpy()      # reuse = (flow21, flow12, metric21, metric12, ...)
            o2 = ref_union.GMFSS_UNION.inference_ts.__wrapped__.__wrapped__(m, I0, I1, [0.5])
            out[f"{tag}_ts0.5"] = o2[0].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "union_golden.npz"), **out)
    print("union_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "rife", "gmfss", "union"]
    if "ops" in which:
        gen_ops()
    if "rife" in which:
        gen_rife()
    if "gmfss" in which:
        gen_gmfss()
    if "union" in which:
        gen_union()
