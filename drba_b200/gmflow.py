"""GMFlow on libdrba_b200.so -- host-side orchestration.

Mirrors models/gmflow/gmflow.py:92-185 with the reference's defaults (2 scales, attn_splits [2, 8],
corr_radius [-1, 4], prop_radius [-1, 1], upsample_factor 4, unidirectional).  Every dense contraction runs on
the tcgen05 engine (csrc/conv_tc.cu): backbone convs, q/k/v/merge/FFN linears as 1x1 convs, the window
attention products QK^T and PV, the global correlation and the propagation scores as batched GEMMs
(`bgemm` layers), the upsampler head.  csrc/gmflow.cu holds the stages in between.  No torch compute ops
run here; torch owns buffers and builds two constant tables at first use (sine position encoding).

Token tensors: [2][h][w][128] NHWC fp16, index 0 = feature0, 1 = feature1 (the reference concatenates both
directions in the batch dimension, transformer.py:294-305; "concat1" is the same buffer with halves swapped).
"""
import ctypes
import math

import torch

from . import _lib
from ._torch_util import ptr, require_cuda, stream_ptr
from .convnet import ACT_GELU, ACT_NONE, ACT_RELU, Step, run_program
from .ifnet import _TcLayer, _pad16, _tc_conv3x3
from .ops import resize_bilinear

C = 128
_LL4 = ctypes.c_longlong * 4


def _tc_linear(weight, bias, device, scale=1.0, stride=1):
    """nn.Linear / 1x1 conv as a one-tap layer: weight [out, in] (or [out, in, 1, 1])."""
    w = weight.float().reshape(weight.shape[0], -1) * scale
    cout, cin = w.shape
    wp = torch.zeros((1, 1, _pad16(cout), _pad16(cin)))
    wp[0, 0, :cout, :cin] = w
    bp = None
    if bias is not None:
        bp = torch.zeros((1, _pad16(cout)))
        bp[0, :cout] = bias.float() * scale
    layer = _TcLayer(wp, bp if bp is not None else torch.zeros((1, _pad16(cout))), [0], [0], stride, 0, cout, 0, device, cin_real=cin)
    if bp is None:
        layer.b = None
    return layer


def _nobias(layer):
    layer.b = None
    return layer


class _Operand:
    """The B matrix of a batched GEMM presented as a 'layer': rows [batch][n_pad][K] fp16."""

    def __init__(self, mat, n, n_pad, k):
        self.w, self.b, self.slope = mat, None, None
        self.cin = self.cin_real = k
        self.G = self.T = 1
        self.dy = (ctypes.c_int * 1)(0)
        self.dx = (ctypes.c_int * 1)(0)
        self.cout_pad, self.cout, self.stride, self.epilogue, self.out_os, self.act = n_pad, n, 1, 0, 1, 0


class GMFlow:
    def __init__(self, sd, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DrbaError("drba_b200.GMFlow needs a CUDA device (no CPU fallback)")
        d = self.device
        sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        self.L = _lib.lib()
        # backbone (backbone.py:46-117); conv1 is 7x7 stride 2 with 3 input channels: CUDA-core direct conv
        w = sd["backbone.conv1.weight"]
        self.conv1_w = w.permute(2, 3, 1, 0).reshape(49, 3, 64).contiguous().to(d)
        self.conv1_b = torch.zeros(64, device=d)
        self.conv1_dy = (ctypes.c_int * 49)(*[ky - 3 for ky in range(7) for kx in range(7)])
        self.conv1_dx = (ctypes.c_int * 49)(*[kx - 3 for ky in range(7) for kx in range(7)])
        self.blocks = []
        for name, stride in (("layer1.0", 1), ("layer1.1", 1), ("layer2.0", 2), ("layer2.1", 1), ("layer3.0", 1), ("layer3.1", 1)):
            p = "backbone." + name
            w1, w2 = sd[p + ".conv1.weight"], sd[p + ".conv2.weight"]
            blk = {"c1": _nobias(_tc_conv3x3(w1, torch.zeros(w1.shape[0]), stride, 0, d)),
                   "c2": _nobias(_tc_conv3x3(w2, torch.zeros(w2.shape[0]), 1, 0, d)), "stride": stride, "cout": w1.shape[0], "ds": None}
            if (p + ".downsample.0.weight") in sd:
                blk["ds"] = _tc_linear(sd[p + ".downsample.0.weight"], sd[p + ".downsample.0.bias"], d, stride=stride)
            self.blocks.append(blk)
        self.conv2 = _tc_linear(sd["backbone.conv2.weight"], sd["backbone.conv2.bias"], d)
        wt = sd["backbone.trident_conv.weight"]
        self.trident = [_nobias(_tc_conv3x3(wt, torch.zeros(C), 1, 0, d)), _nobias(_tc_conv3x3(wt, torch.zeros(C), 2, 0, d))]
        # transformer (transformer.py:107-258): the 1/sqrt(C) of the scores is folded into q_proj
        inv = 1.0 / math.sqrt(C)
        self.tf = []
        for i in range(6):
            layer = {}
            for part in ("self_attn", "cross_attn_ffn"):
                p = f"transformer.layers.{i}.{part}."
                e = {"q": _tc_linear(sd[p + "q_proj.weight"], None, d, scale=inv), "k": _tc_linear(sd[p + "k_proj.weight"], None, d),
                     "v": _tc_linear(sd[p + "v_proj.weight"], None, d), "merge": _tc_linear(sd[p + "merge.weight"], None, d),
                     "g1": sd[p + "norm1.weight"].to(d), "b1": sd[p + "norm1.bias"].to(d)}
                if part == "cross_attn_ffn":
                    e["mlp0"] = _tc_linear(sd[p + "mlp.0.weight"], None, d)
                    e["mlp2"] = _tc_linear(sd[p + "mlp.2.weight"], None, d)
                    e["g2"], e["b2"] = sd[p + "norm2.weight"].to(d), sd[p + "norm2.bias"].to(d)
                layer[part] = e
            self.tf.append(layer)
        self.prop_q = _tc_linear(sd["feature_flow_attn.q_proj.weight"], sd["feature_flow_attn.q_proj.bias"], d)
        self.prop_k = _tc_linear(sd["feature_flow_attn.k_proj.weight"], sd["feature_flow_attn.k_proj.bias"], d)
        self.up0 = _tc_conv3x3(sd["upsampler.0.weight"], sd["upsampler.0.bias"], 1, 0, d)
        self.up2 = _tc_linear(sd["upsampler.2.weight"], sd["upsampler.2.bias"], d)
        self._bufs = {}
        self._pos = {}
        self.debug = None        # set to a dict to capture intermediates (tests)

    # ---------------------------------------------------------------- plumbing
    def buf(self, key, shape, dtype=torch.float16, zero=False):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def st(self):
        return stream_ptr(self.device)

    def _chk(self, rc, what, n=1):
        """rc: the return code of an already issued call, or a zero-argument callable issuing it (then the call
        is bracketed by the launch profiler's events under the family name `gmflow.<what>`)."""
        if callable(rc):
            with _lib.launch("gmflow." + what, n):
                rc = rc()
        else:
            _lib.count(n)
        _lib.check(rc, what)

    def _pos_table(self, wh, ww):
        """Sine position encoding of one attention window (position.py:30-54): fp32 [wh][ww][C]; a constant table."""
        key = (wh, ww)
        t = self._pos.get(key)
        if t is None:
            npf = C // 2
            y = torch.arange(1, wh + 1, dtype=torch.float32).view(wh, 1).expand(wh, ww) / (wh + 1e-6) * (2 * math.pi)
            x = torch.arange(1, ww + 1, dtype=torch.float32).view(1, ww).expand(wh, ww) / (ww + 1e-6) * (2 * math.pi)
            i = torch.arange(npf, dtype=torch.float32)
            dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="floor") / npf)
            px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
            px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
            py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
            t = torch.cat((py, px), dim=2).contiguous().to(self.device)
            self._pos[key] = t
        return t

    def _inorm(self, xs, cch, h, w, outs, relu_x=1, skips=None, skip_stats=None, final_relu=0, tag=""):
        """InstanceNorm (+ReLU) per image, optionally out = relu(skip(+IN) + relu(IN(x)))."""
        stats = []
        for k, x in enumerate(xs):
            s = self.buf(("st", tag, k, cch), (cch, 2), torch.float64)
            s.zero_()
            self._chk(lambda: self.L.drba_gmflow_inorm_stats(ptr(x), cch, h, w, ptr(s), self.st()), "inorm_stats")
            stats.append(s)
        for k, x in enumerate(xs):
            self._chk(lambda: self.L.drba_gmflow_inorm_apply(ptr(x), ptr(stats[k]), relu_x, ptr(skips[k]) if skips else None,
                                                     ptr(skip_stats[k]) if skip_stats else None, final_relu, ptr(outs[k]), cch, h, w, self.st()),
                      "inorm_apply")
        return stats

    def _stats_only(self, xs, cch, h, w, tag):
        out = []
        for k, x in enumerate(xs):
            s = self.buf(("st", tag, k, cch), (cch, 2), torch.float64)
            s.zero_()
            self._chk(lambda: self.L.drba_gmflow_inorm_stats(ptr(x), cch, h, w, ptr(s), self.st()), "inorm_stats")
            out.append(s)
        return out

    # ---------------------------------------------------------------- backbone
    def backbone(self, imgs):
        """imgs: two [1,3,H,W] fp32 -> (features 1/8 [2][H/8][W/8][128], features 1/4 [2][H/4][W/4][128])."""
        _, _, H, W = imgs[0].shape
        h1, w1 = H // 2, W // 2
        x = []
        for k, im in enumerate(imgs):
            nb = self.buf(("norm", k, H, W), (1, 3, H, W), torch.float32)
            self._chk(lambda: self.L.drba_gmflow_normalize_img(ptr(im.float().contiguous()), ptr(nb), H, W, self.st()), "normalize")
            raw = self.buf(("c1raw", k, H, W), (h1, w1, 64))
            self._chk(lambda: self.L.drba_conv2d_direct_f32(ptr(nb), ptr(self.conv1_w), ptr(self.conv1_b), None, ptr(raw), 1, 1, 3, H, W,
                                                            _LL4(3 * H * W, H * W, W, 1), 64, h1, w1, _LL4(h1 * w1 * 64, 1, w1 * 64, 64),
                                                            2, 1, 0, 0, 49, self.conv1_dy, self.conv1_dx, 0, self.st()), "conv1_7x7")
            x.append(raw)
        cur = [self.buf(("x0", k, H, W), (h1, w1, 64)) for k in range(2)]
        self._inorm(x, 64, h1, w1, cur, relu_x=1, tag="c1")
        h, w, cin = h1, w1, 64
        for bi, blk in enumerate(self.blocks):
            s, co = blk["stride"], blk["cout"]
            oh, ow = h // s, w // s
            a_raw = [self.buf(("a_raw", bi, k, H, W), (oh, ow, co)) for k in range(2)]
            run_program([Step(blk["c1"], h, w, cur, a_raw, oh, ow, co, act=ACT_NONE)], self.device, tag="gmflow.backbone")
            a = [self.buf(("a", bi, k, H, W), (oh, ow, co)) for k in range(2)]
            self._inorm(a_raw, co, oh, ow, a, relu_x=1, tag="a")
            b_raw = [self.buf(("b_raw", bi, k, H, W), (oh, ow, co)) for k in range(2)]
            run_program([Step(blk["c2"], oh, ow, a, b_raw, oh, ow, co, act=ACT_NONE)], self.device, tag="gmflow.backbone")
            out = [self.buf(("blk", bi, k, H, W), (oh, ow, co)) for k in range(2)]
            if blk["ds"] is not None:
                sk_raw = [self.buf(("sk_raw", bi, k, H, W), (oh, ow, co)) for k in range(2)]
                run_program([Step(blk["ds"], h, w, cur, sk_raw, oh, ow, co, act=ACT_NONE)], self.device, tag="gmflow.backbone")
                sk_stats = self._stats_only(sk_raw, co, oh, ow, "sk")
                self._inorm(b_raw, co, oh, ow, out, relu_x=1, skips=sk_raw, skip_stats=sk_stats, final_relu=1, tag="b")
            else:
                self._inorm(b_raw, co, oh, ow, out, relu_x=1, skips=cur, final_relu=1, tag="b")
            cur, h, w, cin = out, oh, ow, co
        y = [self.buf(("conv2", k, H, W), (h, w, C)) for k in range(2)]
        f4 = self.buf(("f4", H, W), (2, h, w, C))
        f8 = self.buf(("f8", H, W), (2, h // 2, w // 2, C))
        run_program([Step(self.conv2, h, w, cur, y, h, w, C, act=ACT_NONE),
                     Step(self.trident[0], h, w, y, [f4[0], f4[1]], h, w, C, act=ACT_NONE)], self.device, tag="gmflow.backbone")
        run_program([Step(self.trident[1], h, w, y, [f8[0], f8[1]], h // 2, w // 2, C, act=ACT_NONE)], self.device, tag="gmflow.backbone")
        return f8, f4

    # ---------------------------------------------------------------- transformer
    def _attention_layer(self, e, x, swap_target, h, w, k, shifted, ffn, tag):
        """One TransformerLayer (transformer.py:146-188) applied in place to x [2][h][w][C]."""
        L = self.L
        tgt = [x[1], x[0]] if swap_target else [x[0], x[1]]
        src = [x[0], x[1]]
        q = self.buf(("q", h, w), (2, h, w, C))
        kk = self.buf(("k", h, w), (2, h, w, C))
        v = self.buf(("v", h, w), (2, h, w, C))
        run_program([Step(e["q"], h, w, src, [q[0], q[1]], h, w, C, act=ACT_NONE),
                     Step(e["k"], h, w, tgt, [kk[0], kk[1]], h, w, C, act=ACT_NONE),
                     Step(e["v"], h, w, tgt, [v[0], v[1]], h, w, C, act=ACT_NONE)], self.device, tag="gmflow.qkv")
        nw = 2 * k * k
        wh, ww = h // k, w // k
        Lw = wh * ww
        Lp = _pad16(Lw)
        qw = self.buf(("qw", h, w), (nw, Lw, C))
        kw = self.buf(("kw", h, w), (nw, Lp, C), zero=True)
        vt = self.buf(("vt", h, w), (nw, C, Lp), zero=True)
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(q), ptr(qw), 2, h, w, C, k, int(shifted), Lw, 0, self.st()), "window_pack")
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(kk), ptr(kw), 2, h, w, C, k, int(shifted), Lp, 0, self.st()), "window_pack")
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(v), ptr(vt), 2, h, w, C, k, int(shifted), Lp, 1, self.st()), "window_pack")
        S = self.buf(("S", h, w), (nw, Lw, Lp))
        run_program([Step(_Operand(kw, Lw, Lp, C), nw, Lw, [qw], [S], nw, Lw, Lp, act=ACT_NONE, bgemm=1)], self.device, tag="gmflow.qk")
        self._chk(lambda: L.drba_gmflow_softmax_rows(ptr(S), nw, Lw, Lp, int(shifted), k, h, w, self.st()), "softmax_rows")
        o = self.buf(("o", h, w), (nw, Lw, C))
        m = self.buf(("m", h, w), (nw, Lw, C))
        run_program([Step(_Operand(vt, C, C, Lp), nw, Lw, [S], [o], nw, Lw, C, act=ACT_NONE, bgemm=1),
                     Step(e["merge"], nw, Lw, [o], [m], nw, Lw, C, act=ACT_NONE)], self.device, tag="gmflow.pv")
        if not ffn:
            self._chk(lambda: L.drba_gmflow_ln_residual(ptr(x), ptr(m), ptr(e["g1"]), ptr(e["b1"]), ptr(x), 2, h, w, C, k, int(shifted), Lw, 0, self.st()), "ln_residual")
            return
        cat = self.buf(("cat", h, w), (2, h, w, 2 * C))
        self._chk(lambda: L.drba_gmflow_ln_residual(ptr(x), ptr(m), ptr(e["g1"]), ptr(e["b1"]), ptr(cat), 2, h, w, C, k, int(shifted), Lw, 1, self.st()), "ln_residual")
        hid = self.buf(("hid", h, w), (2, h, w, 8 * C))
        m2 = self.buf(("m2", h, w), (2, h, w, C))
        run_program([Step(e["mlp0"], h, w, [cat[0], cat[1]], [hid[0], hid[1]], h, w, 8 * C, act=ACT_GELU),
                     Step(e["mlp2"], h, w, [hid[0], hid[1]], [m2[0], m2[1]], h, w, C, act=ACT_NONE)], self.device, tag="gmflow.ffn")
        self._chk(lambda: L.drba_gmflow_ln_residual(ptr(x), ptr(m2), ptr(e["g2"]), ptr(e["b2"]), ptr(x), 2, h, w, C, 0, 0, 0, 0, self.st()), "ln_residual")

    # ---------------------------------------------------------------- forward
    def _match_and_propagate(self, src, tgt, flow, h, w, radius, prop_r, s):
        """correlation soft-argmax (matching.py) + flow propagation by self-attention on `src` (transformer.py:332-409).
        src / tgt: [h][w][C] token maps; flow: running flow or None.  Returns the propagated flow [1,2,h,w]."""
        L, dev, dbg = self.L, self.device, self.debug
        n = h * w
        pred = torch.empty((1, 2, h, w), dtype=torch.float32, device=dev)
        if radius < 0:      # global matching (matching.py:7-43)
            S = self.buf(("corr", h, w), (1, n, _pad16(n)))
            run_program([Step(_Operand(tgt.reshape(1, n, C), n, _pad16(n), C), 1, n, [src.reshape(1, n, C)], [S], 1, n, _pad16(n),
                              act=ACT_NONE, bgemm=1)], dev, tag="gmflow.corr")
            self._chk(lambda: L.drba_gmflow_soft_readout(ptr(S), n, n, _pad16(n), None, w, 1, 1.0 / math.sqrt(C), ptr(pred), self.st()), "soft_readout")
        else:
            self._chk(lambda: L.drba_gmflow_local_match(ptr(src), ptr(tgt), h, w, C, radius, ptr(pred), self.st()), "local_match")
        if flow is None:
            flow = pred
        else:
            tot = torch.empty_like(pred)
            self._chk(lambda: L.drba_axpby_f32(ptr(flow), 1.0, ptr(pred), 1.0, ptr(tot), pred.numel(), self.st()), "axpby")
            flow = tot
        if dbg is not None:
            dbg[f"match{s}"] = flow.clone()
        q = self.buf(("pq", h, w), (h, w, C))
        kq = self.buf(("pk", h, w), (h, w, C))
        out = torch.empty_like(flow)
        if prop_r < 0:      # note: k = k_proj(q_proj(x)) in the global branch (transformer.py:355-356)
            run_program([Step(self.prop_q, h, w, [src], [q], h, w, C, act=ACT_NONE),
                         Step(self.prop_k, h, w, [q], [kq], h, w, C, act=ACT_NONE)], dev, tag="gmflow.prop")
            S = self.buf(("corr", h, w), (1, n, _pad16(n)))
            run_program([Step(_Operand(kq.reshape(1, n, C), n, _pad16(n), C), 1, n, [q.reshape(1, n, C)], [S], 1, n, _pad16(n),
                              act=ACT_NONE, bgemm=1)], dev, tag="gmflow.corr")
            val = flow.reshape(2, n).t().contiguous()            # [n][2] value table (layout plumbing)
            self._chk(lambda: L.drba_gmflow_soft_readout(ptr(S), n, n, _pad16(n), ptr(val), w, 0, 1.0 / math.sqrt(C), ptr(out), self.st()), "soft_readout")
        else:
            run_program([Step(self.prop_q, h, w, [src], [q], h, w, C, act=ACT_NONE),
                         Step(self.prop_k, h, w, [src], [kq], h, w, C, act=ACT_NONE)], dev, tag="gmflow.prop")
            self._chk(lambda: L.drba_gmflow_local_propagate(ptr(q), ptr(kq), ptr(flow), h, w, C, ptr(out), self.st()), "local_propagate")
        if dbg is not None:
            dbg[f"prop{s}"] = out.clone()
        return out

    def _refine_and_upsample(self, f4, d, flow8):
        """Scale 1 (1/4 resolution) for direction d (0: img0 -> img1, 1: img1 -> img0) and the convex upsampling."""
        L, dev, dbg = self.L, self.device, self.debug
        h, w, k = f4.shape[1], f4.shape[2], 8
        x = self.buf(("x", 1, h, w), (2, h, w, C))
        x[0].copy_(f4[d])
        up = resize_bilinear(flow8, size=(h, w), align_corners=True)         # gmflow.py:117-119
        flow = torch.empty_like(up)
        self._chk(lambda: L.drba_axpby_f32(ptr(up), 2.0, None, 0.0, ptr(flow), up.numel(), self.st()), "axpby")
        self._chk(lambda: L.drba_gmflow_warp_feature(ptr(f4[1 - d]), ptr(flow), ptr(x[1]), h, w, C, self.st()), "warp_feature")
        pos = self._pos_table(h // k, w // k)
        self._chk(lambda: L.drba_gmflow_add_position(ptr(x), ptr(pos), 2, h, w, h // k, w // k, C, self.st()), "add_position")
        self.transformer_ref_order(x, h, w, k)
        if dbg is not None:
            dbg["tf1"] = x.clone()
        flow = self._match_and_propagate(x[0], x[1], flow, h, w, 4, 1, 1)
        # convex upsampling x4 (gmflow.py:68-90)
        xin = self.buf(("upin", h, w), (h, w, 144))
        self._chk(lambda: L.drba_gmflow_upsampler_input(ptr(flow), ptr(x[0]), ptr(xin), h, w, self.st()), "upsampler_input")
        hid = self.buf(("uphid", h, w), (h, w, 256))
        mask = self.buf(("upmask", h, w), (h, w, 144))
        run_program([Step(self.up0, h, w, [xin], [hid], h, w, 256, act=ACT_RELU),
                     Step(self.up2, h, w, [hid], [mask], h, w, 144, act=ACT_NONE)], dev, tag="gmflow.upsampler")
        out = torch.empty((1, 2, 4 * h, 4 * w), dtype=torch.float32, device=dev)
        self._chk(lambda: L.drba_gmflow_convex_upsample(ptr(mask), ptr(flow), ptr(out), h, w, self.st()), "convex_upsample")
        return out

    def _forward(self, img0, img1, both):
        require_cuda(img0, img1)
        _, _, H, W = img0.shape
        if H % 32 != 0 or W % 32 != 0:
            raise _lib.DrbaError("GMFlow input must be a multiple of 32 (two 1/8 windows, eight 1/4 windows)")
        L, dev, dbg = self.L, self.device, self.debug
        with torch.cuda.device(dev):
            f8, f4 = self.backbone([img0, img1])
            if dbg is not None:
                dbg["feat8"] = f8.clone()
                dbg["feat4"] = f4.clone()
            # scale 0 (1/8 resolution, 2 x 2 windows, global matching).  The transformer runs both directions side
            # by side (transformer.py:294-305), so its output serves flow(img0 -> img1) AND flow(img1 -> img0):
            # the reference's second call GMFlow(img1, img0) recomputes exactly these features with the halves swapped
            h, w, k = f8.shape[1], f8.shape[2], 2
            x = self.buf(("x", 0, h, w), (2, h, w, C))
            x.copy_(f8)
            pos = self._pos_table(h // k, w // k)
            self._chk(lambda: L.drba_gmflow_add_position(ptr(x), ptr(pos), 2, h, w, h // k, w // k, C, self.st()), "add_position")
            self.transformer_ref_order(x, h, w, k)
            if dbg is not None:
                dbg["tf0"] = x.clone()
            outs = []
            for d in ((0, 1) if both else (0,)):
                flow8 = self._match_and_propagate(x[d], x[1 - d], None, h, w, -1, -1, 0)
                outs.append(self._refine_and_upsample(f4, d, flow8))
        return outs

    def __call__(self, img0, img1):
        """flow img0 -> img1, [1,2,H,W] fp32 (models/gmflow/gmflow.py:92-185)."""
        return self._forward(img0, img1, False)[0]

    def pair(self, img0, img1):
        """(flow img0 -> img1, flow img1 -> img0): what Model.reuse obtains from two GMFlow calls
        (models/model_gmfss/GMFSS.py:73-74), sharing the backbone and the 1/8-scale transformer between them."""
        return tuple(self._forward(img0, img1, True))

    def transformer_ref_order(self, x, h, w, k):
        """The reference updates `concat1` (the cross-attention targets) only AFTER a whole block
        (transformer.py:300-305): block i's cross-attention for feature0 reads feature1 as block i-1 left it,
        not feature1 after block i's self-attention.  Both halves live in x, so the targets are snapshotted."""
        snap = self.buf(("snap", h, w), (2, h, w, C))
        for i, layer in enumerate(self.tf):
            shifted = k > 1 and i % 2 == 1
            snap.copy_(x)
            self._attention_layer(layer["self_attn"], x, False, h, w, k, shifted, False, "self")
            self._cross(layer["cross_attn_ffn"], x, snap, h, w, k, shifted)

    def _cross(self, e, x, snap, h, w, k, shifted):
        """cross_attn_ffn with targets taken from the snapshot (swapped halves)."""
        # temporarily present [x0, x1] as sources and [snap1, snap0] as targets
        L = self.L
        src = [x[0], x[1]]
        tgt = [snap[1], snap[0]]
        q = self.buf(("q", h, w), (2, h, w, C))
        kk = self.buf(("k", h, w), (2, h, w, C))
        v = self.buf(("v", h, w), (2, h, w, C))
        run_program([Step(e["q"], h, w, src, [q[0], q[1]], h, w, C, act=ACT_NONE),
                     Step(e["k"], h, w, tgt, [kk[0], kk[1]], h, w, C, act=ACT_NONE),
                     Step(e["v"], h, w, tgt, [v[0], v[1]], h, w, C, act=ACT_NONE)], self.device, tag="gmflow.qkv")
        nw = 2 * k * k
        wh, ww = h // k, w // k
        Lw = wh * ww
        Lp = _pad16(Lw)
        qw = self.buf(("qw", h, w), (nw, Lw, C))
        kw = self.buf(("kw", h, w), (nw, Lp, C), zero=True)
        vt = self.buf(("vt", h, w), (nw, C, Lp), zero=True)
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(q), ptr(qw), 2, h, w, C, k, int(shifted), Lw, 0, self.st()), "window_pack")
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(kk), ptr(kw), 2, h, w, C, k, int(shifted), Lp, 0, self.st()), "window_pack")
        self._chk(lambda: L.drba_gmflow_window_pack(ptr(v), ptr(vt), 2, h, w, C, k, int(shifted), Lp, 1, self.st()), "window_pack")
        S = self.buf(("S", h, w), (nw, Lw, Lp))
        run_program([Step(_Operand(kw, Lw, Lp, C), nw, Lw, [qw], [S], nw, Lw, Lp, act=ACT_NONE, bgemm=1)], self.device, tag="gmflow.qk")
        self._chk(lambda: L.drba_gmflow_softmax_rows(ptr(S), nw, Lw, Lp, int(shifted), k, h, w, self.st()), "softmax_rows")
        o = self.buf(("o", h, w), (nw, Lw, C))
        m = self.buf(("m", h, w), (nw, Lw, C))
        run_program([Step(_Operand(vt, C, C, Lp), nw, Lw, [S], [o], nw, Lw, C, act=ACT_NONE, bgemm=1),
                     Step(e["merge"], nw, Lw, [o], [m], nw, Lw, C, act=ACT_NONE)], self.device, tag="gmflow.pv")
        cat = self.buf(("cat", h, w), (2, h, w, 2 * C))
        self._chk(lambda: L.drba_gmflow_ln_residual(ptr(x), ptr(m), ptr(e["g1"]), ptr(e["b1"]), ptr(cat), 2, h, w, C, k, int(shifted), Lw, 1, self.st()), "ln_residual")
        hid = self.buf(("hid", h, w), (2, h, w, 8 * C))
        m2 = self.buf(("m2", h, w), (2, h, w, C))
        run_program([Step(e["mlp0"], h, w, [cat[0], cat[1]], [hid[0], hid[1]], h, w, 8 * C, act=ACT_GELU),
                     Step(e["mlp2"], h, w, [hid[0], hid[1]], [m2[0], m2[1]], h, w, C, act=ACT_NONE)], self.device, tag="gmflow.ffn")
        self._chk(lambda: L.drba_gmflow_ln_residual(ptr(x), ptr(m2), ptr(e["g2"]), ptr(e["b2"]), ptr(x), 2, h, w, C, 0, 0, 0, 0, self.st()), "ln_residual")
