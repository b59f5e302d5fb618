"""tcgen05 implicit-GEMM conv (csrc/conv_tc.cu) against a plain PyTorch fp32 reference of the
same op on the same fp16-rounded operands.  Tolerance: the kernel accumulates in fp32 and
rounds the result to fp16 once, so |err| <= 2^-11 * |y| + accumulation-order noise."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _engine():
    from drba_b200.ifnet import IFNetEngine
    from drba_b200.weights import synth_ifnet_state
    return IFNetEngine(synth_ifnet_state(0), "cuda", "fp32")   # only used for its _conv_tc plumbing


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


@pytest.mark.parametrize("cin,cout,h,w,stride,res", [
    (64, 64, 16, 32, 1, False),      # Kc = 64 (128B swizzle), exactly 4 tiles
    (32, 32, 17, 30, 1, True),       # Kc = 32 (64B swizzle), ragged tiles, residual
    (16, 32, 24, 40, 1, False),      # Kc = 16 (32B swizzle)
    (48, 96, 34, 60, 1, False),      # Kc = 16, 3 chunks
    (96, 96, 17, 30, 1, True),       # Kc = 32, 3 chunks
    (192, 192, 17, 30, 1, True),     # N split in two tiles of 96
    (128, 128, 34, 60, 1, True),
    (64, 16, 40, 56, 2, False),      # stride 2 (rank-5 parity view)
    (48, 96, 32, 64, 2, False),
    (16, 32, 68, 120, 2, False),
    (32, 64, 33, 50, 1, False),      # packed halo mode (2 pixels per 128-byte line), ragged in both directions
    (16, 16, 37, 52, 1, True),       # packed halo mode (4 pixels per line), residual
    (32, 32, 21, 31, 1, True),       # odd width: falls back to the unpacked halo mode
    (32, 64, 70, 50, 2, False),      # stride-2 halo mode (four parity-plane boxes per stage), Kc = 32, ragged tiles
    (64, 32, 34, 18, 2, False),      # stride-2 halo mode, Kc = 64, tile wider than the output
])
def test_conv3x3_tc_vs_torch(cin, cout, h, w, stride, res):
    from drba_b200.ifnet import _tc_conv3x3
    eng = _engine()
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + h)
    x = torch.randn((1, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, 3, 3), generator=g) * (1.0 / (cin * 9)) ** 0.5
    b = torch.randn((cout,), generator=g) * 0.1
    assert not (res and (stride != 1 or cin != cout))
    layer = _tc_conv3x3(wt, b, stride, 1, "cuda")
    xh = x.half()
    x_nhwc = xh[0].permute(1, 2, 0).contiguous().cuda()
    oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    out = torch.full((oh, ow, layer.cout_pad), float("nan"), dtype=torch.float16, device="cuda")
    eng._conv_tc(layer, x_nhwc, h, w, out, oh, ow, layer.cout_pad, res=x_nhwc if res else None)
    torch.cuda.synchronize()
    ref = F.conv2d(xh.float(), wt.half().float(), b, stride, 1)
    if res:
        ref = ref + xh.float()
    ref = _lrelu(ref)[0].permute(1, 2, 0)
    got = out[:, :, :cout].float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    tol = 2e-3 + 1.5e-3 * ref.abs()
    assert (err <= tol).all(), f"max err {err.max().item():.4g} at |ref| max {ref.abs().max().item():.3g}"


@pytest.mark.parametrize("cin,h,w,form", [(32, 16, 32, "phases"), (64, 17, 30, "phases"), (192, 9, 15, "phases"), (96, 34, 60, "phases"),
                                          (32, 34, 60, "3x3"), (64, 17, 30, "3x3"), (32, 272, 480, "3x3")])
def test_lastconv_tc_vs_torch(cin, h, w, form):
    """ConvTranspose2d(cin, 52, 4, 2, 1) + PixelShuffle(2) (IFNet_HDv3.py:79-82): as four phase convs of four taps, and as
    one zero-padded 3x3 conv with 4 x 64 columns (halo mode, one phase resident per CTA, stores through the staging tile)."""
    from drba_b200.ifnet import _tc_lastconv, _tc_lastconv3x3
    eng = _engine()
    g = torch.Generator(device="cpu").manual_seed(cin + h)
    x = torch.randn((1, cin, h, w), generator=g)
    wt = torch.randn((cin, 52, 4, 4), generator=g) * (1.0 / (cin * 4)) ** 0.5
    b = torch.randn((52,), generator=g) * 0.1
    layer = (_tc_lastconv if form == "phases" else _tc_lastconv3x3)(wt, b, "cuda")
    xh = x.half()
    x_nhwc = xh[0].permute(1, 2, 0).contiguous().cuda()
    out = torch.full((4 * h, 4 * w, 16), float("nan"), dtype=torch.float32, device="cuda")
    eng._conv_tc(layer, x_nhwc, h, w, out, h, w, 16)
    torch.cuda.synchronize()
    ref = F.pixel_shuffle(F.conv_transpose2d(xh.float(), wt.half().float(), b, 2, 1), 2)[0].permute(1, 2, 0)
    got = out[:, :, :13].cpu()
    assert torch.isfinite(out).all()
    err = (got - ref).abs()
    assert (err <= 1e-3 + 1e-3 * ref.abs()).all(), f"max err {err.max().item():.4g}"
    assert (out[:, :, 13:] == 0).all()
    # 8 floats per pixel: the form the last IFBlock writes for the blend (flow + mask live in channels 0..4)
    out8 = torch.full((4 * h, 4 * w, 8), float("nan"), dtype=torch.float32, device="cuda")
    eng._conv_tc(layer, x_nhwc, h, w, out8, h, w, 8)
    torch.cuda.synchronize()
    assert torch.equal(out8, out[:, :, :8])


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def test_rife_fp16_engine_vs_golden(golden_rife):
    """End-to-end RIFE on the tensor-core engine against the reference-generated fp32 golden
    frames.  The reference's own GPU path runs these convs in fp16 (torch.autocast); tolerance:
    PSNR >= 40 dB and max-abs <= 0.05 on [0,1] frames for the trained checkpoint, looser for the
    synthetic weights (random nets amplify rounding)."""
    from drba_b200.rife import RIFE
    from drba_b200.weights import synth_ifnet_state, find_rife_weights, load_ifnet_state
    g = golden_rife
    I0, I1, I2 = (torch.from_numpy(g[k]).cuda() for k in ("I0", "I1", "I2"))
    cases = [("synth", synth_ifnet_state(0), 30.0)]
    wdir = find_rife_weights()
    if wdir is not None:
        cases.append(("real", load_ifnet_state(wdir), 40.0))
    for tag, state, min_psnr in cases:
        m = RIFE(state=state, device="cuda", precision="fp16")
        y = m.inference_ts(I0, I1, [0.4])[0]
        p = _psnr(y.cpu(), torch.from_numpy(g[f"{tag}_ts0.4"]))
        o1, reuse = m.inference_ts_drba(I0, I1, I2, np.array([0.6, 1.0, 1.4]), None, True)
        p1 = _psnr(o1[0].cpu(), torch.from_numpy(g[f"{tag}_w0_0.6"]))
        p2 = _psnr(o1[2].cpu(), torch.from_numpy(g[f"{tag}_w0_1.4"]))
        print(f"[{tag}] fp16 engine PSNR vs fp32 reference: ts0.4 {p:.1f} dB, drba 0.6 {p1:.1f} dB, 1.4 {p2:.1f} dB")
        assert min(p, p1, p2) >= min_psnr, (tag, p, p1, p2)


@pytest.mark.parametrize("c,h,w,nimg", [(32, 24, 40, 1), (64, 17, 30, 2), (192, 17, 30, 2), (16, 68, 120, 2)])
def test_conv_program_chain_vs_torch(c, h, w, nimg):
    """A persistent multi-layer program (stride-2 conv -> 3 x ResConv-style conv with residual, grid
    barrier between layers, 1-2 images sharing the launch) against a layer-by-layer torch reference
    that rounds the activations to fp16 between layers like the kernel does."""
    from drba_b200.ifnet import _tc_conv3x3
    eng = _engine()
    g = torch.Generator(device="cpu").manual_seed(c + h + nimg)
    cin0 = 32
    w0 = torch.randn((c, cin0, 3, 3), generator=g) * (1.0 / (cin0 * 9)) ** 0.5
    b0 = torch.randn((c,), generator=g) * 0.1
    ws = [torch.randn((c, c, 3, 3), generator=g) * (0.5 / (c * 9)) ** 0.5 for _ in range(3)]
    bs = [torch.randn((c,), generator=g) * 0.1 for _ in range(3)]
    l0 = _tc_conv3x3(w0, b0, 2, 1, "cuda")
    ls = [_tc_conv3x3(wi, bi, 1, 1, "cuda") for wi, bi in zip(ws, bs)]
    xs = [torch.randn((1, cin0, 2 * h, 2 * w), generator=g).half() for _ in range(nimg)]
    x_dev = [x[0].permute(1, 2, 0).contiguous().cuda() for x in xs]
    cp = l0.cout_pad
    p0 = [torch.zeros((h, w, cp), dtype=torch.float16, device="cuda") for _ in range(nimg)]
    p1 = [torch.zeros((h, w, cp), dtype=torch.float16, device="cuda") for _ in range(nimg)]
    steps = [(l0, 2 * h, 2 * w, x_dev, p0, h, w, cp, None)]
    cur, nxt = p0, p1
    for layer in ls:
        steps.append((layer, h, w, cur, nxt, h, w, cp, cur))
        cur, nxt = nxt, cur
    for rep in range(2):       # second run checks that the barrier words were left zeroed
        eng._conv_program(steps)
    torch.cuda.synchronize()
    from drba_b200.convnet import _sync
    assert int(_sync(torch.device('cuda', torch.cuda.current_device())).abs().sum().item()) == 0
    for k in range(nimg):
        y = _lrelu(F.conv2d(xs[k].float(), w0.half().float(), b0, 2, 1)).half()
        for wi, bi in zip(ws, bs):
            y = _lrelu(F.conv2d(y.float(), wi.half().float(), bi, 1, 1) + y.float()).half()
        ref = y[0].permute(1, 2, 0).float()
        got = cur[k][:, :, :c].float().cpu()
        assert torch.isfinite(got).all()
        err = (got - ref).abs()
        tol = 6e-3 + 4e-3 * ref.abs()      # 4 layers of fp16 re-rounding
        assert (err <= tol).all(), f"image {k}: max err {err.max().item():.4g}"


@pytest.mark.parametrize("h,w", [(64, 96), (70, 50), (34, 66)])
def test_head_cnn0_fast_path_matches_generic_kernel(monkeypatch, h, w):
    """Head.cnn0 (IFNet_HDv3.py:31) has its own kernel in the tensor-core engine; it must give exactly what the
    generic direct convolution gives, and both must match torch."""
    eng = _engine()
    g = torch.Generator(device="cpu").manual_seed(h * 7 + w)
    img = torch.rand((1, 3, h, w), generator=g).cuda()
    layer = eng.direct["encode.cnn0"]
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    outs = []
    for generic in ("1", "0"):
        monkeypatch.setenv("DRBA_DIRECT_GENERIC", generic)
        out = torch.full((h2, w2, 16), float("nan"), dtype=torch.float16, device="cuda")
        eng._conv_direct(layer, img.data_ptr(), h, w, (3 * h * w, h * w, w, 1), out, h2, w2, (h2 * w2 * 16, 1, w2 * 16, 16))
        torch.cuda.synchronize()
        outs.append(out.view(torch.int16).cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    got = torch.from_numpy(outs[1]).view(torch.float16).float()
    assert torch.isfinite(got).all() and got.abs().max() > 0
