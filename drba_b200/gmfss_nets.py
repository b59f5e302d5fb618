"""GMFSS networks on the tensor-core conv engine: FeatureNet, MetricNet, GridNet.

Mirrors models/model_gmfss/FeatureNet.py:6-33, MetricNet.py:23-65 and FusionNet.py:6-145 for batch 1.
All three are pre-activation nets (PReLU with ONE learned slope, then conv); a tensor is typically consumed
raw (as a residual) AND through one or two PReLUs with different slopes, so the conv epilogue writes up to
three versions of its result (include/drba_b200.h, drba_conv_layer.out/out1/out2) and no stand-alone
activation pass exists.  Activations are NHWC fp16, accumulation fp32 -- the reference's own GPU precision
under torch.autocast (models/gmfss.py:18-19).  Each net is a chain of persistent conv programs
(convnet.run_program): FeatureNet 6 layers (both frames side by side), MetricNet 5, GridNet 45.
"""
import ctypes

import torch

from . import _lib
from ._torch_util import ptr, require_cuda, stream_ptr
from .convnet import ACT_NONE, ACT_PRELU, Step, run_program
from .ifnet import _TcLayer, _pad16, _taps3x3, _tc_conv3x3, _tc_convT


def _tc_pixelshuffle_conv(weight, bias, device):
    """Conv2d(cin, 4*c, 3, 1, 1) + PixelShuffle(2) (FusionNet.py:44-47) as four phase convs (phase = sub-pixel
    (i, j), output channel c of phase i*2+j is conv channel c*4 + i*2 + j) written at (2y+i, 2x+j)."""
    c4, cin = weight.shape[0], weight.shape[1]
    c = c4 // 4
    wp = torch.zeros((4, 9, _pad16(c), _pad16(cin)))
    bp = torch.zeros((4, _pad16(c)))
    dys, dxs = [], []
    dy, dx = _taps3x3()
    for g in range(4):
        wg = weight[g::4].float()                 # [c, cin, 3, 3]
        wp[g, :, :c, :cin] = wg.permute(2, 3, 0, 1).reshape(9, c, cin)
        bp[g, :c] = bias[g::4].float()
        dys += dy
        dxs += dx
    return _TcLayer(wp, bp, dys, dxs, 1, 0, c, 0, device, cin_real=cin, out_os=2)


def _f(t):
    return float(t.reshape(-1)[0])


class _Bufs:
    def __init__(self, device):
        self.device, self.b = device, {}

    def get(self, key, shape, dtype=torch.float16):
        t = self.b.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.b[key] = t
        return t


def pack_planes(planes, out, prelu=None, scales=None):
    """planes: list of [H,W] fp32 CUDA tensors (views are fine if contiguous) -> out [H][W][>=16] fp16."""
    n = len(planes)
    H, W = planes[0].shape[-2:]
    arr = (ctypes.c_void_p * n)(*[p.data_ptr() for p in planes])
    sc = (ctypes.c_float * n)(*(scales if scales is not None else [1.0] * n))
    with _lib.launch("pack_planes", 1, nbytes=float(H * W * (4 * n + 2 * out.shape[-1]))):
        rc = _lib.lib().drba_pack_planes_nhwc_f16(ctypes.addressof(arr), ctypes.addressof(sc), n, H, W,
                                                  0 if prelu is None else 1, 0.0 if prelu is None else float(prelu),
                                                  ptr(out), out.shape[-1], stream_ptr(out.device))
    _lib.check(rc, "drba_pack_planes_nhwc_f16")
    return out


def unpack_planes(x, C, clamp=None, tanh10=False):
    """x [H][W][cstride] fp16 -> [1,C,H,W] fp32 (optionally clamped, or tanh(x) * 10)."""
    H, W, cs = x.shape
    out = torch.empty((1, C, H, W), dtype=torch.float32, device=x.device)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    with _lib.launch("unpack_planes", 1, nbytes=float(H * W * (2 * cs + 4 * C))):
        rc = _lib.lib().drba_unpack_nhwc_f16(ptr(x), cs, ptr(out), C, H, W, 2 if tanh10 else (0 if clamp is None else 1), float(lo), float(hi),
                                             stream_ptr(x.device))
    _lib.check(rc, "drba_unpack_nhwc_f16")
    return out


class FeatureNet:
    """models/model_gmfss/FeatureNet.py:6-33.  __call__(imgs) runs 1-2 frames side by side and returns per frame
    (feat1 [H/2][W/2][64], feat2 [H/4][W/4][128], feat3 [H/8][W/8][192]) NHWC fp16."""

    def __init__(self, sd, device):
        self.device = torch.device(device)
        d = self.device
        self.a = [_f(sd["block1.0.weight"]), _f(sd["block1.2.weight"]), _f(sd["block2.0.weight"]),
                  _f(sd["block2.2.weight"]), _f(sd["block3.0.weight"]), _f(sd["block3.2.weight"])]
        self.L = [_tc_conv3x3(sd["block1.1.weight"], sd["block1.1.bias"], 2, 0, d),
                  _tc_conv3x3(sd["block1.3.weight"], sd["block1.3.bias"], 1, 0, d),
                  _tc_conv3x3(sd["block2.1.weight"], sd["block2.1.bias"], 2, 0, d),
                  _tc_conv3x3(sd["block2.3.weight"], sd["block2.3.bias"], 1, 0, d),
                  _tc_conv3x3(sd["block3.1.weight"], sd["block3.1.bias"], 2, 0, d),
                  _tc_conv3x3(sd["block3.3.weight"], sd["block3.3.bias"], 1, 0, d)]
        self.bufs = _Bufs(d)

    def __call__(self, imgs):
        require_cuda(*imgs)
        n = len(imgs)
        _, _, H, W = imgs[0].shape
        assert H % 8 == 0 and W % 8 == 0
        a, L, B = self.a, self.L, self.bufs
        h1, w1, h2, w2, h3, w3 = H // 2, W // 2, H // 4, W // 4, H // 8, W // 8
        with torch.cuda.device(self.device):
            x0 = [pack_planes(list(im.float().contiguous()[0]), B.get(("x0", k, H, W), (H, W, 16)), prelu=a[0]) for k, im in enumerate(imgs)]
            t1 = [B.get(("t1", k, H, W), (h1, w1, 64)) for k in range(n)]
            f1 = [torch.empty((h1, w1, 64), dtype=torch.float16, device=self.device) for _ in range(n)]
            f1a = [B.get(("f1a", k, H, W), (h1, w1, 64)) for k in range(n)]
            t2 = [B.get(("t2", k, H, W), (h2, w2, 128)) for k in range(n)]
            f2 = [torch.empty((h2, w2, 128), dtype=torch.float16, device=self.device) for _ in range(n)]
            f2a = [B.get(("f2a", k, H, W), (h2, w2, 128)) for k in range(n)]
            t3 = [B.get(("t3", k, H, W), (h3, w3, 192)) for k in range(n)]
            f3 = [torch.empty((h3, w3, 192), dtype=torch.float16, device=self.device) for _ in range(n)]
            steps = [Step(L[0], H, W, x0, t1, h1, w1, 64, act=ACT_PRELU, slope0=a[1]),
                     Step(L[1], h1, w1, t1, f1, h1, w1, 64, act=ACT_NONE, out1=f1a, act1=ACT_PRELU, slope1=a[2]),
                     Step(L[2], h1, w1, f1a, t2, h2, w2, 128, act=ACT_PRELU, slope0=a[3]),
                     Step(L[3], h2, w2, t2, f2, h2, w2, 128, act=ACT_NONE, out1=f2a, act1=ACT_PRELU, slope1=a[4]),
                     Step(L[4], h2, w2, f2a, t3, h3, w3, 192, act=ACT_PRELU, slope0=a[5]),
                     Step(L[5], h3, w3, t3, f3, h3, w3, 192, act=ACT_NONE)]
            run_program(steps, self.device, tag="featurenet")
        return [(f1[k], f2[k], f3[k]) for k in range(n)]


class MetricNet:
    """models/model_gmfss/MetricNet.py:23-65: (img0, img1, flow01, flow10) at half resolution -> metric0, metric1
    [1,1,h,w] fp32."""

    def __init__(self, sd, device, union=False):
        self.device = torch.device(device)
        self.union = union      # model_gmfss_union/MetricNet.py:41-42,63: Tanh() on the output, then * 10
        d = self.device
        self.p = [_f(sd["metric_net1.0.weight"]), _f(sd["metric_net2.0.weight"]), _f(sd["metric_net3.0.weight"]),
                  _f(sd["metric_out.0.weight"])]
        self.L = [_tc_conv3x3(sd["metric_in.weight"], sd["metric_in.bias"], 1, 0, d),
                  _tc_conv3x3(sd["metric_net1.1.weight"], sd["metric_net1.1.bias"], 1, 0, d),
                  _tc_conv3x3(sd["metric_net2.1.weight"], sd["metric_net2.1.bias"], 1, 0, d),
                  _tc_conv3x3(sd["metric_net3.1.weight"], sd["metric_net3.1.bias"], 1, 0, d),
                  _tc_conv3x3(sd["metric_out.1.weight"], sd["metric_out.1.bias"], 1, 0, d)]
        self.bufs = _Bufs(d)

    def __call__(self, img0, img1, flow01, flow10):
        require_cuda(img0, img1, flow01, flow10)
        img0, img1, flow01, flow10 = (t.float().contiguous() for t in (img0, img1, flow01, flow10))
        _, _, h, w = img0.shape
        p, L, B = self.p, self.L, self.bufs
        with torch.cuda.device(self.device):
            x = B.get(("x", h, w), (h, w, 16))
            with _lib.launch("gmfss_metric_prep", 1, nbytes=float(h * w * (40 + 32))):
                rc = _lib.lib().drba_gmfss_metric_prep(ptr(img0), ptr(img1), ptr(flow01), ptr(flow10), ptr(x), h, w,
                                                       stream_ptr(self.device))
            _lib.check(rc, "drba_gmfss_metric_prep")
            raw = [B.get(("raw", i, h, w), (h, w, 64)) for i in range(3)]
            act = [B.get(("act", i, h, w), (h, w, 64)) for i in range(2)]
            o16 = B.get(("o16", h, w), (h, w, 16))
            steps = [Step(L[0], h, w, [x], [raw[0]], h, w, 64, act=ACT_NONE, out1=[act[0]], act1=ACT_PRELU, slope1=p[0]),
                     Step(L[1], h, w, [act[0]], [raw[1]], h, w, 64, res=[raw[0]], act=ACT_NONE, out1=[act[1]], act1=ACT_PRELU, slope1=p[1]),
                     Step(L[2], h, w, [act[1]], [raw[2]], h, w, 64, res=[raw[1]], act=ACT_NONE, out1=[act[0]], act1=ACT_PRELU, slope1=p[2]),
                     Step(L[3], h, w, [act[0]], [act[1]], h, w, 64, res=[raw[2]], act=ACT_PRELU, slope0=p[3]),
                     Step(L[4], h, w, [act[1]], [o16], h, w, 16, act=ACT_NONE)]
            run_program(steps, self.device, tag="metricnet")
            m = unpack_planes(o16, 2, tanh10=self.union)
        return m[:, :1], m[:, 1:2]


class GridNet:
    """models/model_gmfss/FusionNet.py:55-145.  Inputs are the ALREADY pre-activated conv inputs of the four head
    blocks (their producers apply the head PReLU on write, see `head_slopes`):
        x  [h][w][16]  (12 real: img0, I1t, I2t, img1)      x1 [h][w][128]
        x2 [h/2][w/2][256]                                  x3 [h/4][w/4][384]
    Returns the frame [1,3,2h,2w] fp32 clamped to [0,1] (GMFSS.py:190)."""

    def __init__(self, sd, device):
        self.device = torch.device(device)
        d = self.device
        def block(prefix, stride=1, up=False):
            a0, a1 = _f(sd[prefix + ".0.weight"]), _f(sd[prefix + ".2.weight"])
            if up:
                c1 = _tc_convT(sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], d)
            else:
                c1 = _tc_conv3x3(sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], stride, 0, d)
            c2 = _tc_conv3x3(sd[prefix + ".3.weight"], sd[prefix + ".3.bias"], 1, 0, d)
            return a0, a1, c1, c2

        self.blk = {}
        for name in ("head", "head1", "head2", "head3", "01", "04", "05", "11", "14", "15", "21", "24", "25"):
            src = "head0" if (name == "head" and "residual_model_head0.0.weight" in sd) else name     # union names its image head `head0`
            self.blk[name] = block("residual_model_" + src)
        for name in ("10", "20", "11", "21"):
            self.blk["d" + name] = block("downsample_model_" + name, stride=2)
        for name in ("04", "14", "05", "15"):
            self.blk["u" + name] = block("upsample_model_" + name, up=True)
        t = "residual_model_tail."
        self.tail_before = _tc_conv3x3(sd[t + "conv_before_upsample.0.weight"], sd[t + "conv_before_upsample.0.bias"], 1, 0, d)
        self.tail_slope = _f(sd[t + "conv_before_upsample.1.weight"])
        self.tail_up = _tc_pixelshuffle_conv(sd[t + "upsample.0.weight"], sd[t + "upsample.0.bias"], d)
        self.tail_last = _tc_conv3x3(sd[t + "conv_last.weight"], sd[t + "conv_last.bias"], 1, 0, d)
        self.head_slopes = tuple(self.blk[k][0] for k in ("head", "head1", "head2", "head3"))
        self.bufs = _Bufs(d)

    def __call__(self, x, x1, x2, x3):
        h, w = x.shape[0], x.shape[1]
        assert h % 4 == 0 and w % 4 == 0
        hb, wb, hc, wc = h // 2, w // 2, h // 4, w // 4
        B, K = self.bufs, self.blk
        A = lambda n: B.get((n, h, w), (h, w, 64))          # noqa: E731
        Bb = lambda n: B.get((n, h, w), (hb, wb, 128))      # noqa: E731
        C = lambda n: B.get((n, h, w), (hc, wc, 192))       # noqa: E731
        S = []

        def conv(layer, src, dst, ih, iw, oh, ow, cs, **kw):
            for k in ("res", "out1", "out2"):
                if kw.get(k) is not None:
                    kw[k] = [kw[k]]
            S.append(Step(layer, ih, iw, [src], [dst], oh, ow, cs, **kw))

        def two(name, src, tmp, dst, dims, res=None, outs=()):
            """conv1 (+PReLU) -> conv2 (+res) with up to two extra activated outputs [(buffer, slope), ...]."""
            a0, a1, c1, c2 = K[name]
            ih, iw, oh, ow, cs = dims       # input grid, output grid of the block, channels of the block
            if c1.out_os == 2:              # ConvTranspose: the launch geometry is the INPUT grid
                conv(c1, src, tmp, ih, iw, ih, iw, cs, act=ACT_PRELU, slope0=a1)
            else:
                conv(c1, src, tmp, ih, iw, oh, ow, cs, act=ACT_PRELU, slope0=a1)
            kw = {"act": ACT_NONE, "res": res}
            raw_needed = dst is not None
            extra = list(outs)
            if not raw_needed:      # only an activated version is consumed: it becomes the primary output
                buf, sl = extra.pop(0)
                dst = buf
                kw.update(act=ACT_PRELU, slope0=sl)
            if len(extra) > 0:
                kw.update(out1=extra[0][0], act1=ACT_PRELU, slope1=extra[0][1])
            if len(extra) > 1:
                kw.update(out2=extra[1][0], act2=ACT_PRELU, slope2=extra[1][1])
            conv(c2, tmp, dst, oh, ow, oh, ow, cs, **kw)

        dA, dB, dC = (h, w, h, w, 64), (hb, wb, hb, wb, 128), (hc, wc, hc, wc, 192)
        s0 = lambda n: K[n][0]      # first PReLU slope of a block        # noqa: E731
        with torch.cuda.device(self.device):
            two("head", x, A("tA"), A("H0"), dA)
            two("head1", x1, A("tA"), A("X00"), dA, res=A("H0"), outs=[(A("X00a"), s0("01")), (A("X00b"), s0("d10"))])
            two("01", A("X00a"), A("tA"), A("X01"), dA, res=A("X00"), outs=[(A("X01a"), s0("04")), (A("X01b"), s0("d11"))])
            two("d10", A("X00b"), Bb("tB"), Bb("D10"), (h, w, hb, wb, 128))
            two("head2", x2, Bb("tB"), Bb("X10"), dB, res=Bb("D10"), outs=[(Bb("X10a"), s0("11")), (Bb("X10b"), s0("d20"))])
            two("d20", Bb("X10b"), C("tC"), C("D20"), (hb, wb, hc, wc, 192))
            two("head3", x3, C("tC"), C("X20"), dC, res=C("D20"), outs=[(C("X20a"), s0("21"))])
            two("11", Bb("X10a"), Bb("tB"), Bb("R11"), dB, res=Bb("X10"))
            two("d11", A("X01b"), Bb("tB"), Bb("X11"), (h, w, hb, wb, 128), res=Bb("R11"),
                outs=[(Bb("X11a"), s0("14")), (Bb("X11b"), s0("d21"))])
            two("21", C("X20a"), C("tC"), C("R21"), dC, res=C("X20"))
            two("d21", Bb("X11b"), C("tC"), C("X21"), (hb, wb, hc, wc, 192), res=C("R21"), outs=[(C("X21a"), s0("24"))])
            two("24", C("X21a"), C("tC"), C("X24"), dC, res=C("X21"), outs=[(C("X24a"), s0("25")), (C("X24b"), s0("u14"))])
            two("25", C("X24a"), C("tC"), None, dC, res=C("X24"), outs=[(C("X25a"), s0("u15"))])
            two("14", Bb("X11a"), Bb("tB"), Bb("R14"), dB, res=Bb("X11"))
            two("u14", C("X24b"), Bb("tB"), Bb("X14"), (hc, wc, hb, wb, 128), res=Bb("R14"),
                outs=[(Bb("X14a"), s0("u04")), (Bb("X14b"), s0("15"))])
            two("04", A("X01a"), A("tA"), A("R04"), dA, res=A("X01"))
            two("u04", Bb("X14a"), A("tA"), A("X04"), (hb, wb, h, w, 64), res=A("R04"), outs=[(A("X04a"), s0("05"))])
            two("15", Bb("X14b"), Bb("tB"), Bb("R15"), dB, res=Bb("X14"))
            two("u15", C("X25a"), Bb("tB"), None, (hc, wc, hb, wb, 128), res=Bb("R15"), outs=[(Bb("X15a"), s0("u05"))])
            two("05", A("X04a"), A("tA"), A("R05"), dA, res=A("X04"))
            two("u05", Bb("X15a"), A("tA"), A("X05"), (hb, wb, h, w, 64), res=A("R05"))
            # tail: conv + PReLU, conv + PixelShuffle(2), conv_last (FusionNet.py:36-52)
            conv(self.tail_before, A("X05"), A("tA"), h, w, h, w, 64, act=ACT_PRELU, slope0=self.tail_slope)
            T1 = B.get(("T1", h, w), (2 * h, 2 * w, 64))
            conv(self.tail_up, A("tA"), T1, h, w, h, w, 64, act=ACT_NONE)
            o16 = B.get(("o16", h, w), (2 * h, 2 * w, 16))
            conv(self.tail_last, T1, o16, 2 * h, 2 * w, 2 * h, 2 * w, 16, act=ACT_NONE)
            run_program(S, self.device, tag="gridnet")
            return unpack_planes(o16, 3, clamp=(0.0, 1.0))
