"""Parity at the size bench.py measures (net input 1088 x 1920, BASELINE.json configs[1]) and of the host logic
that only shows with the real model: multi-tile-per-CTA TMEM double buffering, super tiles, packed / stride-2 halo
boxes at scale, two-image programs, CUDA-graph `reuse` hand-over, frame-window shards of the real RIFE.

Every measured figure is also appended to gpurun_out/fullsize_parity.json (one JSON object per line) so that the
numbers behind the assertions can be quoted in DESIGN.md."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 1088, 1920


def _record(**kw):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "fullsize_parity.json"), "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _state():
    from drba_b200.weights import find_rife_weights, load_ifnet_state, synth_ifnet_state
    wdir = find_rife_weights()
    if wdir is not None:
        return load_ifnet_state(wdir), "real"
    return synth_ifnet_state(0), "synth"


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else float(10 * np.log10(1.0 / mse))


@pytest.fixture(scope="module")
def tf32_off():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _torch_weight(layer):
    """[cout_pad, cin, 3, 3] fp32 weight and [cout_pad] bias of a packed 3x3 layer (w[1][9][cout_pad][cin])."""
    w = layer.w[0].float()[:9]                             # [9][cout_pad][cin] (a tenth tap is the ResConv's identity tap)
    return w.permute(1, 2, 0).reshape(w.shape[1], w.shape[2], 3, 3).contiguous(), layer.b[0].float()


# (name, layer key, input h, w): every 3x3 conv shape of the fine IFNet levels of a 1088 x 1920 window + Head
_LAYERS_1080P = [
    ("block4.conv0a", 1088, 1920), ("block4.conv0b", 544, 960), ("block4.res3", 272, 480),
    ("block3.conv0a", 544, 960), ("block3.conv0b", 272, 480), ("block3.res0", 136, 240),
    ("block2.conv0a", 272, 480), ("block2.conv0b", 136, 240), ("block2.res7", 68, 120),
    ("block1.res1", 34, 60), ("block0.res5", 17, 30), ("encode.cnn1", 544, 960),
]


@pytest.mark.parametrize("key,h,w", _LAYERS_1080P)
def test_conv_tc_1080p_layer_vs_torch(tf32_off, key, h, w):
    """Each conv-program layer shape of the benchmarked window (full-size tile grids: > 148 tiles per layer) against
    torch's fp32 convolution (cuDNN, TF32 off) of the same fp16-rounded operands on the GPU.
    Tolerance: fp32 accumulation + one fp16 rounding: |err| <= 2e-3 + 1.5e-3 |y|."""
    from drba_b200.ifnet import IFNetEngine
    state, _ = _state()
    eng = IFNetEngine(state, "cuda", "fp16")
    layer = eng.tc[key]
    g = torch.Generator(device="cuda").manual_seed(h * 3 + w)
    x = torch.randn((h, w, layer.cin), device="cuda", generator=g).half()
    s = layer.stride
    oh, ow = (h - 1) // s + 1, (w - 1) // s + 1
    res = x if (s == 1 and layer.cin == layer.cout_pad and "res" in key) else None
    out = torch.full((oh, ow, layer.cout_pad), float("nan"), dtype=torch.float16, device="cuda")
    # (layers with an identity tap add the residual on the tensor core: no `res` pointer)
    eng._conv_tc(layer, x, h, w, out, oh, ow, layer.cout_pad, res=None if getattr(layer, "res_tap", False) else res)
    wt, b = _torch_weight(layer)
    ref = F.conv2d(x.float().permute(2, 0, 1)[None], wt, b, s, 1)
    if res is not None:
        ref = ref + x.float().permute(2, 0, 1)[None]
    ref = F.leaky_relu(ref, 0.2)[0].permute(1, 2, 0)
    got = out.float()
    torch.cuda.synchronize()
    assert torch.isfinite(got[:, :, :layer.cout]).all()
    err = (got - ref)[:, :, :layer.cout].abs()
    tol = 2e-3 + 1.5e-3 * ref[:, :, :layer.cout].abs()
    _record(test="conv_layer_1080p", layer=key, h=h, w=w, max_err=float(err.max()), ref_max=float(ref.abs().max()))
    assert (err <= tol).all(), f"{key}: max err {float(err.max()):.4g}, {int((err > tol).sum())} elements out of tolerance"


@pytest.mark.parametrize("bi,h,w", [(4, 1088, 1920), (3, 544, 960), (2, 272, 480)])
def test_block_program_1080p_two_images_vs_torch(tf32_off, bi, h, w):
    """The whole persistent program of an IFBlock (conv0a, conv0b, 8 x ResConv, lastconv; grid barriers; two images
    sharing the launch) at the benchmarked size against a layer-by-layer torch chain that rounds the activations to
    fp16 between layers like the kernel does.  IFNet_HDv3.py:84-96."""
    from drba_b200.ifnet import IFNetEngine, _BLOCKS
    state, _ = _state()
    eng = IFNetEngine(state, "cuda", "fp16")
    name, _, c = _BLOCKS[bi]
    g = torch.Generator(device="cuda").manual_seed(bi)
    cin_pad = eng.tc[f"{name}.conv0a"].cin
    xs = [(0.5 * torch.randn((h, w, cin_pad), device="cuda", generator=g)).half() for _ in range(2)]
    h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
    f16 = torch.float16
    a = [torch.empty((h2, w2, c // 2), dtype=f16, device="cuda") for _ in range(2)]
    p0 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(2)]
    p1 = [torch.empty((h4, w4, c), dtype=f16, device="cuda") for _ in range(2)]
    tmp = [torch.full((h, w, 16), float("nan"), dtype=torch.float32, device="cuda") for _ in range(2)]
    steps = [(eng.tc[f"{name}.conv0a"], h, w, xs, a, h2, w2, c // 2, None),
             (eng.tc[f"{name}.conv0b"], h2, w2, a, p0, h4, w4, c, None)]
    cur, nxt = p0, p1
    for i in range(8):
        steps.append((eng.tc[f"{name}.res{i}"], h4, w4, cur, nxt, h4, w4, c, None if getattr(eng.tc[f"{name}.res{i}"], "res_tap", False) else cur))
        cur, nxt = nxt, cur
    steps.append((eng.tc[f"{name}.last"], h4, w4, cur, tmp, h4, w4, 16, None))
    eng._conv_program(steps)
    torch.cuda.synchronize()
    sd = {k: v.float().cuda() for k, v in state.items()}
    for k in range(2):
        y = xs[k].float().permute(2, 0, 1)[None]
        for key in (f"{name}.conv0a", f"{name}.conv0b"):
            wt, b = _torch_weight(eng.tc[key])
            y = F.leaky_relu(F.conv2d(y, wt, b, 2, 1), 0.2)[:, :eng.tc[key].cout].half().float()
        for i in range(8):
            wt, b = _torch_weight(eng.tc[f"{name}.res{i}"])
            y = F.leaky_relu(F.conv2d(y, wt, b, 1, 1) + y, 0.2).half().float()
        wl = sd[f"{name}.lastconv.0.weight"].half().float()
        ref = F.pixel_shuffle(F.conv_transpose2d(y, wl, sd[f"{name}.lastconv.0.bias"], 2, 1), 2)[0].permute(1, 2, 0)
        got = tmp[k][:, :, :13]
        assert torch.isfinite(got).all()
        err = (got - ref).abs()
        scale = float(ref.abs().max())
        # ten layers of fp16 re-rounding: a rounding flip early in the chain moves later activations by one fp16 ulp
        bad = err > (2e-2 * scale + 2e-2 * ref.abs())
        _record(test="block_program_1080p", block=name, image=k, max_err=float(err.max()), ref_max=scale,
                mean_err=float(err.mean()), frac_bad=float(bad.float().mean()))
        assert float(err.mean()) <= 2e-3 * scale, f"{name} image {k}: mean err {float(err.mean()):.4g} (range {scale:.3g})"
        assert float(bad.float().mean()) <= 1e-5, f"{name} image {k}: max err {float(err.max()):.4g} (range {scale:.3g})"


def _bench_clip(n):
    import bench
    return bench.synth_clip(n, H, W, 1000, "cpu")


def _noise_clip(n):
    g = torch.Generator(device="cpu").manual_seed(5)
    return [torch.rand((1, 3, H, W), generator=g) for _ in range(n)]


# precision -> (min PSNR dB, max tolerated fraction of pixels off by more than 2/255) per clip kind
# measured on B200 (gpurun_out/fullsize_parity.json, trained weights): fp32 engine 113-123 dB (bench clip) / 82-88 dB
# (noise); fp16 engine 73.8-88.3 dB (bench clip) / 28.9-37.5 dB (noise)
_WINDOW_BARS = {("fp32", "bench"): (90.0, 1e-5), ("fp32", "noise"): (70.0, 1e-4),
                ("fp16", "bench"): (60.0, 5e-4), ("fp16", "noise"): (25.0, 0.2)}


@pytest.mark.parametrize("clip", ["bench", "noise"])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_rife_drba_windows_1080p_vs_oracle(precision, clip):
    """Two consecutive DRBA windows (cold start, then `reuse`) of RIFE at 1088 x 1920 -- the benchmarked
    configuration, with bench.py's synthetic clip and with a U(0,1) worst-case clip -- against the CPU
    oracle (oracle/ifnet.py, fp32).  models/rife.py:77-109.
    Bars (PSNR on [0,1] frames; fraction of pixels off by > 2/255): see _WINDOW_BARS.  The noise clip has no
    recoverable motion: flow decisions sit on knife edges, so it bounds robustness, not accuracy."""
    from drba_b200.rife import RIFE
    from oracle.ifnet import RIFEOracle
    torch.set_grad_enabled(False)
    state, wtag = _state()
    frames = _bench_clip(4) if clip == "bench" else _noise_clip(4)
    ora = RIFEOracle(state)
    m = RIFE(state=state, device="cuda", precision=precision)
    dev_frames = [f.cuda() for f in frames]
    ts_list = [np.array([0.6, 1.0, 1.4]), np.array([0.8, 1.2])]
    reuse_o = reuse_g = None
    min_psnr, max_bad = _WINDOW_BARS[(precision, clip)]
    for j, ts in enumerate(ts_list):
        want, reuse_o = ora.inference_ts_drba(frames[j], frames[j + 1], frames[j + 2], ts, reuse_o, True)
        got, reuse_g = m.inference_ts_drba(dev_frames[j], dev_frames[j + 1], dev_frames[j + 2], ts, reuse_g, True)
        torch.cuda.synchronize()
        for t, a, b in zip(ts, got, want):
            if t == 1.0:
                assert a is dev_frames[j + 1]
                continue
            a = a.float().cpu()
            p = _psnr(a, b)
            err = (a - b).abs()
            bad = float((err > 2.0 / 255.0).float().mean())
            _record(test="rife_window_1080p", precision=precision, clip=clip, weights=wtag, window=j, t=float(t),
                    psnr=p, max_abs=float(err.max()), mean_abs=float(err.mean()), frac_gt_2_255=bad)
            assert p >= min_psnr, f"{precision}/{clip} window {j} t={t}: PSNR {p:.1f} dB < {min_psnr}"
            assert bad <= max_bad, f"{precision}/{clip} window {j} t={t}: {bad:.2e} of the pixels off by > 2/255"
    # the flows handed to the next window: compare away from hole flips (the `< 0.999` decision, rife.py:66-70)
    for got, want in zip(reuse_g[:2], reuse_o[:2]):
        got, want = got.float().cpu(), want
        big = 2.0 * float(max(H, W))          # holes: max(H, W) * 2 (rife.py:69-73)
        flips = (got == big) != (want == big)
        frac = float(flips.float().mean())
        d = (got - want).abs()[~flips]
        _record(test="rife_window_1080p_flow", precision=precision, clip=clip, flip_frac=frac, flow_mean_err=float(d.mean()),
                flow_max_err=float(d.max()))
        assert frac < (2e-3 if precision == "fp32" else 3e-2)


@pytest.mark.parametrize("h,w,scale", [(768, 1280, 1.0), (576, 1024, 1.0), (320, 576, 1.0), (1152, 1920, 0.5), (2176, 3840, 0.5)])
def test_rife_drba_window_other_sizes_vs_oracle(h, w, scale):
    """The conv engine picks tiles, N splits, resident / streamed weights and halo modes per layer SIZE (narrow resident N
    splits for the coarse levels, packed halo boxes, grid-vs-split checks): one DRBA window of the fp16 engine at sizes
    other than the benchmarked one -- 720p (768 rows), 1024 x 576, a small frame, 1080p and 4K at scale 0.5 (GMFSS_union's RIFE
    configuration) -- against the CPU oracle.  models/rife.py:77-109.  Bar: PSNR >= 60 dB, <= 5e-4 of the pixels off by
    more than 2/255 (the bars of the 1080p test)."""
    import bench
    from drba_b200.rife import RIFE
    from oracle.ifnet import RIFEOracle
    torch.set_grad_enabled(False)
    state, wtag = _state()
    frames = bench.synth_clip(3, h, w, 1000, "cpu")
    ts = np.array([0.6, 1.0, 1.4])
    want, _ = RIFEOracle(state, scale=scale).inference_ts_drba(frames[0], frames[1], frames[2], ts, None, True)
    m = RIFE(state=state, scale=scale, device="cuda", precision="fp16")
    dev = [f.cuda() for f in frames]
    got, _ = m.inference_ts_drba(dev[0], dev[1], dev[2], ts, None, True)
    torch.cuda.synchronize()
    for t, a, b in zip(ts, got, want):
        if t == 1.0:
            continue
        a = a.float().cpu()
        p = _psnr(a, b)
        bad = float(((a - b).abs() > 2.0 / 255.0).float().mean())
        _record(test="rife_window_other_sizes", size=[h, w], scale=scale, weights=wtag, t=float(t), psnr=p, frac_gt_2_255=bad)
        assert p >= 60.0, f"{h}x{w} scale {scale} t={t}: PSNR {p:.1f} dB"
        assert bad <= 5e-4, f"{h}x{w} scale {scale} t={t}: {bad:.2e} of the pixels off by > 2/255"


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_graph_replay_same_key_windows_match_eager(precision):
    """Three and more consecutive windows with IDENTICAL timestamp lists (-t N, or 30 -> 60 fps): the graph of that
    shape receives its own outputs back as `reuse` (f1 of window k is the static f0-input of the same graph); the
    hand-over must not read a buffer it has already overwritten.  Also: a second frame size on the same model (the
    workspace grows) must not invalidate the first size's graphs."""
    from drba_b200.rife import RIFE
    from drba_b200.weights import synth_ifnet_state
    torch.set_grad_enabled(False)
    state = synth_ifnet_state(0)
    g = torch.Generator(device="cpu").manual_seed(3)

    def clip(h, w, n):
        base = F.interpolate(torch.rand((1, 3, h // 8 + 4, w // 8 + 8), generator=g), scale_factor=8, mode="bilinear")
        return [base[:, :, 2 * i:2 * i + h, 3 * i:3 * i + w].contiguous().cuda() for i in range(n)]

    small, large = clip(64, 128, 7), clip(128, 256, 5)
    ts = np.array([0.75, 1.25])
    outs = {}
    for graphs in (True, False):
        m = RIFE(state=state, device="cuda", precision=precision, graphs=graphs)
        res = []
        reuse = None
        for j in range(3):
            o, reuse = m.inference_ts_drba(small[j], small[j + 1], small[j + 2], ts, reuse, True)
            res += [x.clone() for x in o]
        reuse_l = None
        for j in range(3):       # larger frames on the same model: new graphs, larger workspace
            o, reuse_l = m.inference_ts_drba(large[j], large[j + 1], large[j + 2], ts, reuse_l, True)
            res += [x.clone() for x in o]
        for j in range(3, 5):    # back to the first size: its graphs (and their baked workspace addresses) still valid
            o, reuse = m.inference_ts_drba(small[j], small[j + 1], small[j + 2], ts, reuse, True)
            res += [x.clone() for x in o]
        torch.cuda.synchronize()
        outs[graphs] = res
    assert len(outs[True]) == len(outs[False]) == 16
    for k, (a, b) in enumerate(zip(outs[True], outs[False])):
        # identical kernels on identical inputs; only the fp32 atomics of the flow inversion / DRM reorder sums
        d = float((a - b).abs().max())
        assert d <= 2e-3, f"output {k}: graphs vs eager differ by {d:.3g}"
        assert _psnr(a.cpu(), b.cpu()) >= 70.0


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_address_keyed_graphs_match_eager(precision):
    """A driver loop hands the same buffers round and round (frame ring; `reuse` = the previous graph's outputs): after
    the second sighting of a (shape, timestamps, input addresses) combination the window is replayed from a graph
    captured directly on those addresses (no input copies).  Five passes over a ring of 6 frames with the 24 -> 60
    timestamp alternation, graphs on vs off: every output must agree; the ring contents CHANGE between passes (the
    graphs must read what the caller's tensors hold now)."""
    from drba_b200.rife import RIFE
    from drba_b200.weights import synth_ifnet_state
    torch.set_grad_enabled(False)
    state = synth_ifnet_state(0)
    g = torch.Generator(device="cpu").manual_seed(13)
    h, w, ring = 64, 128, 6
    base = F.interpolate(torch.rand((1, 3, h // 8 + 16, w // 8 + 24), generator=g), scale_factor=8, mode="bilinear")
    tss = [np.array([0.6, 1.0, 1.4]), np.array([0.8, 1.2])]
    outs = {}
    for graphs in (True, False):
        m = RIFE(state=state, device="cuda", precision=precision, graphs=graphs)
        bufs = [torch.empty((1, 3, h, w), device="cuda") for _ in range(ring)]
        res, reuse, j = [], None, 0
        for p in range(5):
            for k in range(ring):      # new content in the same buffers every pass
                bufs[k].copy_(base[:, :, 2 * k + 3 * p:2 * k + 3 * p + h, 3 * k + 5 * p:3 * k + 5 * p + w])
            for k in range(ring):
                o, reuse = m.inference_ts_drba(bufs[k], bufs[(k + 1) % ring], bufs[(k + 2) % ring], tss[j % 2], reuse, True)
                res += [x.clone() for x in o]
                j += 1
        torch.cuda.synchronize()
        outs[graphs] = res
        if graphs:
            assert len(m._agraphs) >= ring, "address-keyed graphs were never captured"
    assert len(outs[True]) == len(outs[False])
    for k, (a, b) in enumerate(zip(outs[True], outs[False])):
        d = float((a - b).abs().max())
        assert d <= 2e-3, f"output {k}: graphs vs eager differ by {d:.3g}"


def test_address_graph_table_is_bounded_and_second_caller_gets_graphs(monkeypatch):
    """The address-keyed table must stay small and keep serving NEW callers: (a) two loops over two different frame
    rings on one model (bench.py's device-resident loop, then its end-to-end loop) must both end up replaying address
    graphs -- the first round-2 version keyed on the `reuse` addresses as well, never closed its cycle, filled the table
    with transients and left the second caller on the generic graph; (b) with a table smaller than the number of
    distinct windows the least recently used graphs are evicted and results still equal the eager path."""
    from drba_b200.rife import RIFE
    from drba_b200.weights import synth_ifnet_state
    torch.set_grad_enabled(False)
    state = synth_ifnet_state(0)
    g = torch.Generator(device="cpu").manual_seed(17)
    h, w = 64, 128
    base = F.interpolate(torch.rand((1, 3, h // 8 + 16, w // 8 + 24), generator=g), scale_factor=8, mode="bilinear")
    tss = [np.array([0.6, 1.0, 1.4]), np.array([0.8, 1.2])]

    def run(m, ring, passes, off):
        bufs = [torch.empty((1, 3, h, w), device="cuda") for _ in range(ring)]
        res, reuse, j = [], None, 0
        for p in range(passes):
            for k in range(ring):
                bufs[k].copy_(base[:, :, 2 * k + 3 * p + off:2 * k + 3 * p + off + h, 3 * k + 5 * p:3 * k + 5 * p + w])
            for k in range(ring):
                o, reuse = m.inference_ts_drba(bufs[k], bufs[(k + 1) % ring], bufs[(k + 2) % ring], tss[j % 2], reuse, True)
                res += [x.clone() for x in o]
                j += 1
        torch.cuda.synchronize()
        return res

    eager = RIFE(state=state, device="cuda", precision="fp16", graphs=False)
    m = RIFE(state=state, device="cuda", precision="fp16", graphs=True)
    a = run(m, 8, 4, 0)
    n_first = len(m._agraphs)
    assert 8 <= n_first <= 16, f"first caller: {n_first} address graphs for a ring of 8 (the cycle must close)"
    c0 = m.captures
    b = run(m, 4, 6, 1)                      # a second caller with its own ring on the same model
    assert len(m._agraphs) <= n_first + 8
    assert m.captures - c0 <= 8, "second caller keeps capturing: its address cycle does not close"
    for got, want in ((a, run(eager, 8, 4, 0)), (b, run(eager, 4, 6, 1))):
        assert len(got) == len(want)
        for k, (x, y) in enumerate(zip(got, want)):
            assert float((x - y).abs().max()) <= 2e-3, f"output {k}"
    # (b) a table of 3 graphs for 8 distinct windows: eviction on every capture, same results
    monkeypatch.setattr(RIFE, "MAX_ADDRESS_GRAPHS", 3)
    m2 = RIFE(state=state, device="cuda", precision="fp16", graphs=True)
    c = run(m2, 8, 3, 0)
    assert len(m2._agraphs) <= 3
    for k, (x, y) in enumerate(zip(c, run(eager, 8, 3, 0))):
        assert float((x - y).abs().max()) <= 2e-3, f"evicting table, output {k}"


@pytest.mark.parametrize("size", [(64, 128), (1088, 1920)])
def test_lazy_flow_terms_match_materialised_flow(size):
    """Blocks 1 and 2 evaluate the flow as a sum of up-sampled lastconv outputs at their own sample positions and the
    first full-resolution flow state is written in one three-term pass before block 3 (drba_ifnet_assemble_terms,
    drba_ifnet_flow_sum) -- against the four accumulate passes of IFNet_HDv3.py:157 (engine.flow_terms = False).
    Same sums in the same order; only FMA contraction inside the sampling kernels differs."""
    from drba_b200.ifnet import IFNetEngine
    torch.set_grad_enabled(False)
    state, _ = _state()
    h, w = size
    g = torch.Generator(device="cpu").manual_seed(17)
    base = F.interpolate(torch.rand((1, 3, h // 8 + 8, w // 8 + 8), generator=g), scale_factor=8, mode="bicubic").clamp(0, 1)
    I0 = base[:, :, 0:h, 0:w].contiguous().cuda()
    I1 = base[:, :, 3:3 + h, 5:5 + w].contiguous().cuda()
    tmap = (0.3 + 0.4 * torch.rand((1, 1, h, w), generator=g)).cuda()
    eng = IFNetEngine(state, "cuda", "fp16")
    outs = []
    for lazy in (True, False):
        eng.flow_terms = lazy
        outs.append([y.clone() for y in eng.forward_multi([(I0, I1, tmap, None, None), (I1, I0, 0.5, None, None)], [16, 8, 4, 2, 1])])
    torch.cuda.synchronize()
    for a, b in zip(*outs):
        p = _psnr(a.cpu(), b.cpu())
        d = float((a - b).abs().max())
        _record(test="lazy_flow_terms", size=list(size), psnr=p, max_abs=d)
        assert p >= 70.0 and d <= 2e-2, (p, d)


@pytest.mark.parametrize("with_cut", [False, True])
def test_real_rife_shards_equal_sequential(with_cut):
    """driver.interpolate_shard x 2 shards of a 12-frame clip against driver.interpolate_sequence with the REAL model
    (CUDA graphs on, `reuse` aliasing graph buffers): the mid-stream `reuse` rebuild (one calc_flow)
    must reproduce the sequential state.  infer.py:112-156."""
    from drba_b200 import driver
    from drba_b200.rife import RIFE
    torch.set_grad_enabled(False)
    state, _ = _state()
    g = torch.Generator(device="cpu").manual_seed(9)
    h, w, n = 128, 256, 12
    base = F.interpolate(torch.rand((1, 3, h // 8 + 8, w // 8 + 8), generator=g), scale_factor=8, mode="bicubic").clamp(0, 1)
    frames = [base[:, :, 2 * i:2 * i + h, 3 * i:3 * i + w].contiguous().cuda() for i in range(n)]
    if with_cut:
        other = F.interpolate(torch.rand((1, 3, h // 8 + 8, w // 8 + 8), generator=g), scale_factor=8, mode="bicubic").clamp(0, 1)
        for i in range(6, n):
            frames[i] = other[:, :, i:i + h, 2 * i:2 * i + w].contiguous().cuda()
    index = {f.data_ptr(): i for i, f in enumerate(frames)}

    def scene(a, b):
        return with_cut and index[a.data_ptr()] == 5 and index[b.data_ptr()] == 6

    m = RIFE(state=state, device="cuda")
    seq = [x.clone() for x in driver.interpolate_sequence(m, frames, 24.0, 60.0, check_scene=scene)]
    parts = []
    for a, b in driver.shard_ranges(driver.num_iterations(n), 2):
        m2 = RIFE(state=state, device="cuda")          # a shard is its own replica
        parts += [x.clone() for x in driver.interpolate_shard(m2, frames, 24.0, 60.0, check_scene=scene, a=a, b=b)]
    torch.cuda.synchronize()
    assert len(parts) == len(seq)
    worst = 0.0
    for k, (x, y) in enumerate(zip(parts, seq)):
        d = float((x.float() - y.float()).abs().max())
        worst = max(worst, d)
        # same kernels, same inputs: only the atomics of the scatter kernels reorder fp32 sums
        assert d <= 2e-3, f"output {k}: shard vs sequential differ by {d:.3g}"
    _record(test="real_rife_shards", with_cut=with_cut, outputs=len(seq), worst_max_abs=worst)


def test_drm_and_invert_flow_1080p_vs_oracle():
    """calc_drm_rife and the calc_flow inversion at 1088 x 1920 against the C oracle on a smooth flow with
    occlusion folds (drm.py:65-107, rife.py:59-73)."""
    from drba_b200.drm import calc_drm_rife
    from drba_b200.ops import rife_invert_flow
    from oracle import cport
    rng = np.random.default_rng(2)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    f10 = np.stack([6 * np.sin(yy / 90) + 3, 4 * np.cos(xx / 70)], 0)[None].astype(np.float32)
    f12 = (-f10 * 0.8 + rng.standard_normal(f10.shape).astype(np.float32) * 0.3).astype(np.float32)
    f10[:, :, 300:500, 600:900] += 25.0            # a fold: many-to-one targets
    want = cport.calc_drm_rife(0.4, f10, f12, True)
    got = calc_drm_rife(0.4, torch.from_numpy(f10).cuda(), torch.from_numpy(f12).cuda(), True)
    for k in want:
        a, b = got[k].cpu().numpy(), want[k]
        flips = np.abs(a - b) > 1e-4                # hole decisions on the 0.999 knife edge
        _record(test="drm_rife_1080p", key=k, flip_frac=float(flips.mean()), max_err=float(np.abs(a - b)[~flips].max()))
        assert flips.mean() < 1e-4
    wi = cport.rife_invert_flow(f10)
    gi = rife_invert_flow(torch.from_numpy(f10).cuda()).cpu().numpy()
    flips = (gi == 2.0 * W) != (wi == 2.0 * W)
    assert flips.mean() < 1e-4
    np.testing.assert_allclose(gi[~flips], wi[~flips], rtol=1e-4, atol=1e-3)
