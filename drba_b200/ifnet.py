"""IFNet 4.26-heavy engine: host-side orchestration of the CUDA kernels in libdrba_b200.so.

Mirrors models/rife_426_heavy/IFNet_HDv3.py (IFNet.forward :126-177, IFBlock :62-96, Head
:28-47, ResConv :50-59) for the inference branch the DRBA wrappers use (batch 1).  No torch
compute ops run here: torch only owns the device buffers and the stream.

Per refinement block:   assemble (warp+cat+resize, 1 kernel) -> conv0a -> conv0b -> 8 x ResConv
-> lastconv (ConvTranspose + PixelShuffle) -> upsample/accumulate (1 kernel); then one blend
kernel.  Two conv engines share this schedule:
  precision="fp32": csrc/conv_direct.cu (CUDA cores, exact reference arithmetic up to summation order)
  precision="fp16": csrc/conv_tc.cu (tcgen05 implicit GEMM, fp16 operands / fp32 accumulate -- the
                    reference's own GPU precision under torch.autocast, models/rife.py:26)
"""
import ctypes
import os

import torch

from . import _lib
from ._torch_util import ptr, require_cuda, stream_ptr

_BLOCKS = [("block0", 39, 192), ("block1", 52, 128), ("block2", 52, 96), ("block3", 52, 64), ("block4", 52, 32)]

_LL4 = ctypes.c_longlong * 4


def _taps3x3():
    dy = [ky - 1 for ky in range(3) for kx in range(3)]
    dx = [kx - 1 for ky in range(3) for kx in range(3)]
    return dy, dx


# ConvTranspose2d(k=4, s=2, p=1): output row 2y+py reads input rows y+dy with kernel row ky
_CT_PHASE = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}   # py -> [(ky, dy), ...]


class _DirectLayer:
    """One conv layer packed for drba_conv2d_direct_f32: w[T][Cin][Cout]."""

    def __init__(self, w, b, dy, dx, stride, act, device):
        self.w = w.contiguous().to(device)
        self.b = b.contiguous().to(device)
        self.T, self.cin, self.cout = self.w.shape
        self.dy = (ctypes.c_int * self.T)(*dy)
        self.dx = (ctypes.c_int * self.T)(*dx)
        self.stride = stride
        self.act = act


def _pack_conv3x3(weight, bias, beta=None):
    w = weight.float()
    b = bias.float()
    if beta is not None:   # ResConv: conv(x) * beta + x  (IFNet_HDv3.py:58-59) -> fold beta
        bt = beta.float().reshape(-1)
        w = w * bt[:, None, None, None]
        b = b * bt
    return w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]), b


def _pack_convT_phase(weight, py, px):
    # weight: [Cin, Cout, 4, 4]
    taps, dy, dx = [], [], []
    for ky, ddy in _CT_PHASE[py]:
        for kx, ddx in _CT_PHASE[px]:
            taps.append(weight[:, :, ky, kx].float())
            dy.append(ddy)
            dx.append(ddx)
    return torch.stack(taps, 0), dy, dx


class _TcLayer:
    """One conv layer packed for drba_conv_tc_f16: w[G][T][cout_pad][cin_pad] fp16, bias[G][cout_pad] fp32."""

    def __init__(self, w, b, dy, dx, stride, act, cout, epilogue, device, cin_real=None, out_os=1):
        self.out_os = out_os
        self.cin_real = cin_real if cin_real is not None else w.shape[3]
        self.w = w.to(torch.float16).contiguous().to(device)
        self.b = b.float().contiguous().to(device)
        self.G, self.T, self.cout_pad, self.cin = self.w.shape
        self.cout = cout
        self.dy = (ctypes.c_int * (self.G * self.T))(*dy)
        self.dx = (ctypes.c_int * (self.G * self.T))(*dx)
        self.stride = stride
        self.act = act
        self.epilogue = epilogue
        self.slope = None


def _pad16(n):
    return (n + 15) // 16 * 16


def _packed_input_channels(first):
    """Packed channel -> reference channel of an IFBlock's conv input as written by
    drba_ifnet_assemble (NHWC fp16): [f0 16 | f1 16 | img0 3, img1 3, (first: timestep), 0.. |
    timestep, mask, feat 8, flow 4, 0, 0]; -1 = zero padding.  See csrc/ifnet.cu ref_channel()."""
    m = [6 + k for k in range(16)] + [22 + k for k in range(16)]
    m += [0, 1, 2, 3, 4, 5] + ([38] if first else [-1]) + [-1] * 9
    if not first:
        m += [38 + k for k in range(14)] + [-1, -1]
    return m


def _tc_conv3x3(weight, bias, stride, act, device, beta=None, in_map=None, res_tap=False):
    """res_tap: append a tenth tap at offset (0, 0) whose weight tile is the identity -- the ResConv's `+ x`
    (IFNet_HDv3.py:59) is then accumulated by the tensor core, exactly (x * 1.0 in the fp32 accumulator), and the
    epilogue has no residual to load; the layer is run WITHOUT a `res` pointer."""
    w, b = _pack_conv3x3(weight, bias, beta)           # [9][Cin][Cout]
    if res_tap:
        assert stride == 1 and w.shape[1] == w.shape[2] and in_map is None
        w = torch.cat([w, torch.eye(w.shape[1])[None]], 0)
    T, cin, cout = w.shape
    if in_map is not None:                             # permute / pad the input channels
        wm = torch.zeros((T, len(in_map), cout))
        for pc, rc in enumerate(in_map):
            if rc >= 0:
                wm[:, pc, :] = w[:, rc, :]
        w, cin_real, cin = wm, cin, len(in_map)
    else:
        cin_real = cin
    wp = torch.zeros((1, T, _pad16(cout), _pad16(cin)))
    wp[0, :, :cout, :cin] = w.permute(0, 2, 1)
    bp = torch.zeros((1, _pad16(cout)))
    bp[0, :cout] = b
    dy, dx = _taps3x3()
    if res_tap:
        dy, dx = dy + [0], dx + [0]
    layer = _TcLayer(wp, bp, dy, dx, stride, act, cout, 0, device, cin_real=cin_real)
    layer.res_tap = bool(res_tap)
    layer.flop_taps = 9          # algorithmic FLOPs of the 3x3 conv (the identity tap is the residual add)
    return layer


def _tc_lastconv(weight, bias, device):
    cin, cout = weight.shape[0], weight.shape[1]       # [Cin, 52, 4, 4]
    wp = torch.zeros((4, 4, 64, cin))
    bp = torch.zeros((4, 64))
    dys, dxs = [], []
    for py in (0, 1):
        for px in (0, 1):
            w, dy, dx = _pack_convT_phase(weight, py, px)   # [4][Cin][Cout]
            wp[py * 2 + px, :, :cout, :] = w.permute(0, 2, 1)
            bp[py * 2 + px, :cout] = bias.float()
            dys += dy
            dxs += dx
    return _TcLayer(wp, bp, dys, dxs, 1, 0, cout, 1, device)


def _tc_lastconv3x3(weight, bias, device):
    """The same lastconv as ONE 3x3 conv with 4 x 64 output columns (phase-major): a phase's weights are zero on the
    five taps it does not use.  2.25x the MMAs, but the layer can then run in the conv engine's halo mode with one
    phase's weights resident per CTA: its input is read once per tile instead of once per (phase, tap)."""
    cin, cout = weight.shape[0], weight.shape[1]       # [Cin, 52, 4, 4]
    wp = torch.zeros((1, 9, 256, cin))
    bp = torch.zeros((1, 256))
    for py in (0, 1):
        for px in (0, 1):
            g = py * 2 + px
            for ky, ddy in _CT_PHASE[py]:
                for kx, ddx in _CT_PHASE[px]:
                    wp[0, (ddy + 1) * 3 + (ddx + 1), g * 64:g * 64 + cout, :] = weight[:, :, ky, kx].float().t()
            bp[0, g * 64:g * 64 + cout] = bias.float()
    dy, dx = _taps3x3()
    layer = _TcLayer(wp, bp, dy, dx, 1, 0, cout, 1, device)
    layer.flop_taps = 16        # algorithmic FLOPs: four phases x four taps, not the nine zero-padded taps x one group issued
    return layer


def _tc_convT(weight, bias, device):
    """ConvTranspose2d(cin, cout, 4, 2, 1) as four phase convs writing NHWC fp16 at (2y+py, 2x+px)."""
    cin, cout = weight.shape[0], weight.shape[1]
    cp = _pad16(cout)
    wp = torch.zeros((4, 4, cp, cin))
    bp = torch.zeros((4, cp))
    dys, dxs = [], []
    for py in (0, 1):
        for px in (0, 1):
            w, dy, dx = _pack_convT_phase(weight, py, px)
            wp[py * 2 + px, :, :cout, :] = w.permute(0, 2, 1)
            bp[py * 2 + px, :cout] = bias.float()
            dys += dy
            dxs += dx
    return _TcLayer(wp, bp, dys, dxs, 1, 0, cout, 0, device, out_os=2)


class IFNetEngine:
    def __init__(self, state, device, precision="fp32"):
        if precision not in ("fp32", "fp16"):
            raise ValueError("precision must be 'fp32' or 'fp16'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DrbaError("IFNetEngine needs a CUDA device (no CPU fallback)")
        self.precision = precision
        self.L = _lib.lib()
        self._bufs = {}
        dy9, dx9 = _taps3x3()
        sd = {k: v.detach().float().cpu() for k, v in state.items()}
        d = self.device
        self.direct = {}
        for name, cin, c in _BLOCKS:
            w, b = _pack_conv3x3(sd[f"{name}.conv0.0.0.weight"], sd[f"{name}.conv0.0.0.bias"])
            self.direct[f"{name}.conv0a"] = _DirectLayer(w, b, dy9, dx9, 2, 1, d)
            w, b = _pack_conv3x3(sd[f"{name}.conv0.1.0.weight"], sd[f"{name}.conv0.1.0.bias"])
            self.direct[f"{name}.conv0b"] = _DirectLayer(w, b, dy9, dx9, 2, 1, d)
            for i in range(8):
                p = f"{name}.convblock.{i}"
                w, b = _pack_conv3x3(sd[p + ".conv.weight"], sd[p + ".conv.bias"], sd[p + ".beta"])
                self.direct[f"{name}.res{i}"] = _DirectLayer(w, b, dy9, dx9, 1, 1, d)
            for py in (0, 1):
                for px in (0, 1):
                    w, dy, dx = _pack_convT_phase(sd[f"{name}.lastconv.0.weight"], py, px)
                    self.direct[f"{name}.last{py}{px}"] = _DirectLayer(w, sd[f"{name}.lastconv.0.bias"], dy, dx, 1, 0, d)
        for i in range(3):
            w, b = _pack_conv3x3(sd[f"encode.cnn{i}.weight"], sd[f"encode.cnn{i}.bias"])
            self.direct[f"encode.cnn{i}"] = _DirectLayer(w, b, dy9, dx9, 2 if i == 0 else 1, 1, d)
        for py in (0, 1):
            for px in (0, 1):
                w, dy, dx = _pack_convT_phase(sd["encode.cnn3.weight"], py, px)
                self.direct[f"encode.cnn3.{py}{px}"] = _DirectLayer(w, sd["encode.cnn3.bias"], dy, dx, 1, 0, d)
        self.tc = {}
        if precision == "fp16":
            for name, cin, c in _BLOCKS:
                self.tc[f"{name}.conv0a"] = _tc_conv3x3(sd[f"{name}.conv0.0.0.weight"], sd[f"{name}.conv0.0.0.bias"], 2, 1, d,
                                                        in_map=_packed_input_channels(name == "block0"))
                self.tc[f"{name}.conv0b"] = _tc_conv3x3(sd[f"{name}.conv0.1.0.weight"], sd[f"{name}.conv0.1.0.bias"], 2, 1, d)
                # blocks 3 / 4 (many tiles per SM, epilogue-bound): the residual as an identity tap
                res_tap = c <= 64 and os.environ.get("DRBA_RES_TAP", "1") != "0"
                for i in range(8):
                    p = f"{name}.convblock.{i}"
                    self.tc[f"{name}.res{i}"] = _tc_conv3x3(sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, 1, d, sd[p + ".beta"],
                                                            res_tap=res_tap)
                # blocks 3 / 4 (64 / 32 channels: one phase's 3x3 weights fit shared memory): the 3x3 form
                last3 = c <= 64 and os.environ.get("DRBA_LAST3X3", "1") != "0"
                self.tc[f"{name}.last"] = (_tc_lastconv3x3 if last3 else _tc_lastconv)(sd[f"{name}.lastconv.0.weight"], sd[f"{name}.lastconv.0.bias"], d)
            for i in (1, 2):
                self.tc[f"encode.cnn{i}"] = _tc_conv3x3(sd[f"encode.cnn{i}.weight"], sd[f"encode.cnn{i}.bias"], 1, 1, d)
            self.tc["encode.cnn3"] = _tc_convT(sd["encode.cnn3.weight"], sd["encode.cnn3.bias"], d)
        self.flow_terms = os.environ.get("DRBA_FLOW_TERMS", "1") != "0"
        # opt-in: measured SLOWER than the two separate kernels on B200 (block 4: 560 us vs 300 + 94 us; the 16 producer
        # warps that fit next to the operand stages run at 40 % issue utilisation)
        self.fused_conv0a = os.environ.get("DRBA_FUSED_CONV0A", "0") != "0"
        self.launches = 0   # kernels launched through this engine (bench.py reports it)
        self._sync = None   # grid-barrier words of the persistent conv programs

    # ------------------------------------------------------------------ helpers
    def _buf(self, key, shape, dtype=torch.float32):
        t = self._bufs.get(key)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def _check(self, rc, what):
        self.launches += 1
        _lib.check(rc, what)

    @staticmethod
    def _launch(name, flops=0.0, nbytes=0.0):
        return _lib.launch(name, 1, flops, nbytes)

    def _conv_direct(self, layer, x_ptr, H, W, in_strides, out_t, OH, OW, out_strides, OS=1, PY=0, PX=0, res_ptr=None):
        out_dtype = 1 if out_t.dtype == torch.float16 else 0
        with self._launch("conv_direct_f32", flops=2.0 * layer.T * layer.cin * layer.cout * OH * OW):
            rc = self.L.drba_conv2d_direct_f32(x_ptr, ptr(layer.w), ptr(layer.b), res_ptr, ptr(out_t), out_dtype,
                                               1, layer.cin, H, W, _LL4(*in_strides),
                                               layer.cout, OH, OW, _LL4(*out_strides),
                                               layer.stride, OS, PY, PX, layer.T, layer.dy, layer.dx,
                                               layer.act, stream_ptr(self.device))
        self._check(rc, "drba_conv2d_direct_f32")

    def _conv_tc(self, layer, x, H, W, out, OH, OW, out_cstride, res=None, tag=""):
        # algorithmic FLOPs: real channels only (2 * Cin * Cout * taps * Hout * Wout)
        with self._launch("conv_tc_f16" + (("/" + tag) if tag else ""), flops=2.0 * getattr(layer, "flop_taps", layer.G * layer.T) * layer.cin_real * layer.cout * OH * OW):
            rc = self.L.drba_conv_tc_f16(ptr(x), H, W, layer.cin, ptr(layer.w), ptr(layer.b), layer.G, layer.T,
                                         layer.dy, layer.dx, layer.cout_pad, layer.cout, layer.stride, OH, OW,
                                         layer.epilogue, layer.act, ptr(res), ptr(out), out_cstride, layer.out_os, stream_ptr(self.device))
        self._check(rc, "drba_conv_tc_f16")

    def _conv_program(self, steps, tag=""):
        """One persistent launch for a chain of tensor-core conv layers (convnet.run_program).
        steps: [(layer, H, W, ins, outs, OH, OW, out_cstride, ress)] with ins/outs/ress lists of one tensor
        per image (ress may be None); layer i+1 may read what layer i wrote."""
        from .convnet import run_program
        run_program(steps, self.device, tag)
        self.launches += 1

    @staticmethod
    def _nchw(c, h, w):
        return (c * h * w, h * w, w, 1)

    # ------------------------------------------------------------------ Head (IFNet_HDv3.py:28-47)
    def encode(self, img):
        """img [1,3,H,W] fp32 -> feature map [H][W][16] (NHWC; fp32 in the exact engine, fp16 in the
        tensor-core engine)."""
        require_cuda(img)
        img = img.float().contiguous()
        _, _, H, W = img.shape
        assert H % 2 == 0 and W % 2 == 0
        h2, w2 = H // 2, W // 2
        if self.precision == "fp16":
            # first conv (Cin = 3) on CUDA cores straight into NHWC fp16, the rest on the tensor cores
            f16 = torch.float16
            a = self._buf(("enc_ah", H, W), (h2, w2, 16), f16)
            b = self._buf(("enc_bh", H, W), (h2, w2, 16), f16)
            feat = torch.empty((H, W, 16), dtype=f16, device=self.device)
            with torch.cuda.device(self.device):
                self._conv_direct(self.direct["encode.cnn0"], ptr(img), H, W, self._nchw(3, H, W), a, h2, w2,
                                  (h2 * w2 * 16, 1, w2 * 16, 16))
                self._conv_program([(self.tc["encode.cnn1"], h2, w2, [a], [b], h2, w2, 16, None),
                                    (self.tc["encode.cnn2"], h2, w2, [b], [a], h2, w2, 16, None),
                                    (self.tc["encode.cnn3"], h2, w2, [a], [feat], h2, w2, 16, None)], tag="encode")
            return feat
        a = self._buf(("enc_a", H, W), (16, h2, w2))
        b = self._buf(("enc_b", H, W), (16, h2, w2))
        feat = torch.empty((H, W, 16), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._conv_direct(self.direct["encode.cnn0"], ptr(img), H, W, self._nchw(3, H, W), a, h2, w2, self._nchw(16, h2, w2))
            self._conv_direct(self.direct["encode.cnn1"], ptr(a), h2, w2, self._nchw(16, h2, w2), b, h2, w2, self._nchw(16, h2, w2))
            self._conv_direct(self.direct["encode.cnn2"], ptr(b), h2, w2, self._nchw(16, h2, w2), a, h2, w2, self._nchw(16, h2, w2))
            nhwc = (H * W * 16, 1, W * 16, 16)
            for py in (0, 1):
                for px in (0, 1):
                    self._conv_direct(self.direct[f"encode.cnn3.{py}{px}"], ptr(a), h2, w2, self._nchw(16, h2, w2),
                                      feat, h2, w2, nhwc, OS=2, PY=py, PX=px)
        return feat

    # ------------------------------------------------------------------ IFBlock (IFNet_HDv3.py:84-96)
    def _assemble(self, out, out_dtype, cstride, img0, img1, f0, f1, timestep, ts_scalar, flow, prev, H, W, s, terms=None):
        tmp_prev, layout_prev, s_prev = prev if prev is not None else (None, 0, 1)
        if terms:
            # flow = sum of up-sampled lastconv outputs, evaluated at this block's sample positions (coarse blocks)
            (t0, _, s0), (t1, _, s1) = terms[0], (terms[1] if len(terms) > 1 else (None, 0, 0))
            with self._launch("ifnet_assemble"):
                rc = self.L.drba_ifnet_assemble_terms(ptr(img0), ptr(img1), ptr(f0), ptr(f1), ptr(timestep), float(ts_scalar),
                                                      ptr(t0), s0, ptr(t1), s1, ptr(tmp_prev), s_prev, ptr(out), H, W, s,
                                                      stream_ptr(self.device))
            self._check(rc, "drba_ifnet_assemble_terms")
            return
        with self._launch("ifnet_assemble"):
            rc = self.L.drba_ifnet_assemble(ptr(img0), ptr(img1), ptr(f0), ptr(f1), 1 if f0.dtype == torch.float16 else 0,
                                            ptr(timestep), float(ts_scalar),
                                            None if prev is None else ptr(flow), ptr(tmp_prev), layout_prev, s_prev,
                                            ptr(out), out_dtype, cstride, H, W, s, stream_ptr(self.device))
        self._check(rc, "drba_ifnet_assemble")

    def _block_conv0a(self, name, jobs, outs, H, W, s):
        """Block input + conv0a in one kernel (drba_ifnet_block_conv0a_f16) for all jobs of the window."""
        layer = self.tc[f"{name}.conv0a"]
        arr = (_lib.BlockInput * len(jobs))()
        for b, j, o in zip(arr, jobs, outs):
            tmp_prev, _, s_prev = j["prev"]
            b.img0, b.img1, b.f0, b.f1 = ptr(j["img0"]), ptr(j["img1"]), ptr(j["f0"]), ptr(j["f1"])
            b.timestep, b.timestep_scalar = ptr(j["ts_t"]), float(j["ts_s"])
            b.flow, b.tmp_prev, b.s_prev, b.out = ptr(j["flow"]), ptr(tmp_prev), s_prev, ptr(o)
        oh, ow = H // s // 2, W // s // 2
        with _lib.launch("ifnet_block_conv0a/" + name, 1, flops=len(jobs) * 2.0 * 9 * layer.cin_real * layer.cout * oh * ow):
            rc = self.L.drba_ifnet_block_conv0a_f16(ctypes.addressof(arr), len(jobs), ptr(layer.w), ptr(layer.b), layer.cout,
                                                    H, W, s, stream_ptr(self.device))
        self._check(rc, "drba_ifnet_block_conv0a_f16")

    def _block(self, bi, jobs, H, W, s):
        """Runs block `bi` for every job (one job = one image pair: dict with img0, img1, f0, f1, ts_t, ts_s,
        flow, prev); returns per job (tmp, layout, s): the block's lastconv output (13 ch at 1/s).
        Tensor-core engine: one assemble per job, then ONE persistent conv program for all jobs."""
        name, cin, c = _BLOCKS[bi]
        h, w = H // s, W // s
        assert H % (4 * s) == 0 and W % (4 * s) == 0, "frame size must be a multiple of 4 * scale"
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        if self.precision == "fp16":
            f16 = torch.float16
            cin_pad = 48 if bi == 0 else 64
            nj = len(jobs)
            a = [self._buf(("ah", bi, H, W, k), (h2, w2, c // 2), f16) for k in range(nj)]
            # blocks 3 / 4 (scale 2 / 1, materialised flow): the block input is assembled inside the kernel that runs
            # conv0a (csrc/ifnet_fused.cu) -- the 64-channel input is never written to memory
            fused = (self.fused_conv0a and s in (1, 2) and c // 2 in (16, 32) and h % 2 == 0 and w % 2 == 0
                     and all(j["prev"] is not None and j["flow"] is not None and not j.get("terms") and j["prev"][1] == 1 for j in jobs))
            if fused:
                self._block_conv0a(name, jobs, a, H, W, s)
            else:
                xs = [self._buf(("xh", bi, H, W, k), (h, w, cin_pad), f16) for k in range(nj)]
                for k, j in enumerate(jobs):
                    self._assemble(xs[k], 1, cin_pad, j["img0"], j["img1"], j["f0"], j["f1"], j["ts_t"], j["ts_s"], j["flow"], j["prev"], H, W, s,
                                   terms=j.get("terms"))
            p0 = [self._buf(("p0h", bi, H, W, k), (h4, w4, c), f16) for k in range(nj)]
            p1 = [self._buf(("p1h", bi, H, W, k), (h4, w4, c), f16) for k in range(nj)]
            # the last block's output is read by the blend only (flow + mask): 8 floats per pixel instead of 16
            tch = 8 if bi == len(_BLOCKS) - 1 else 16
            tmp = [self._buf(("tmp13", bi, H, W, k), (h, w, tch), torch.float32) for k in range(nj)]
            steps = [] if fused else [(self.tc[f"{name}.conv0a"], h, w, xs, a, h2, w2, c // 2, None)]
            steps.append((self.tc[f"{name}.conv0b"], h2, w2, a, p0, h4, w4, c, None))
            cur, nxt = p0, p1
            for i in range(8):
                layer = self.tc[f"{name}.res{i}"]
                steps.append((layer, h4, w4, cur, nxt, h4, w4, c, None if layer.res_tap else cur))
                cur, nxt = nxt, cur
            steps.append((self.tc[f"{name}.last"], h4, w4, cur, tmp, h4, w4, tch, None))
            self._conv_program(steps, tag=name)
            return [(t, 2 if tch == 8 else 1, s) for t in tmp]
        # exact engine: NCHW fp32 activations
        outs = []
        for k, j in enumerate(jobs):
            x = self._buf(("x", bi, H, W, k), (cin, h, w))
            self._assemble(x, 0, 0, j["img0"], j["img1"], j["f0"], j["f1"], j["ts_t"], j["ts_s"], j["flow"], j["prev"], H, W, s)
            a = self._buf(("a", bi, H, W), (c // 2, h2, w2))
            self._conv_direct(self.direct[f"{name}.conv0a"], ptr(x), h, w, self._nchw(cin, h, w), a, h2, w2, self._nchw(c // 2, h2, w2))
            p0 = self._buf(("p0", bi, H, W), (c, h4, w4))
            p1 = self._buf(("p1", bi, H, W), (c, h4, w4))
            st4 = self._nchw(c, h4, w4)
            self._conv_direct(self.direct[f"{name}.conv0b"], ptr(a), h2, w2, self._nchw(c // 2, h2, w2), p0, h4, w4, st4)
            cur, nxt = p0, p1
            for i in range(8):
                self._conv_direct(self.direct[f"{name}.res{i}"], ptr(cur), h4, w4, st4, nxt, h4, w4, st4, res_ptr=ptr(cur))
                cur, nxt = nxt, cur
            ct = self._buf(("ct", bi, H, W, k), (52, h2, w2))
            for py in (0, 1):
                for px in (0, 1):
                    self._conv_direct(self.direct[f"{name}.last{py}{px}"], ptr(cur), h4, w4, st4, ct, h4, w4,
                                      self._nchw(52, h2, w2), OS=2, PY=py, PX=px)
            outs.append((ct, 0, s))
        return outs

    def _flow_sum(self, tmps, flow, H, W):
        (t0, _, s0), (t1, _, s1), (t2, _, s2) = tmps[0], tmps[1], tmps[2]
        with self._launch("ifnet_flow_accum", nbytes=float(H * W * 16)):
            rc = self.L.drba_ifnet_flow_sum(ptr(t0), s0, ptr(t1), s1, ptr(t2), s2, 3, ptr(flow), H, W, stream_ptr(self.device))
        self._check(rc, "drba_ifnet_flow_sum")

    def _flow_accum(self, prev, flow, planar, accumulate, H, W):
        tmp, layout, s = prev
        with self._launch("ifnet_flow_accum", nbytes=float(H * W * (32 if accumulate else 16) + (16 * H * W if planar is not None else 0))):
            rc = self.L.drba_ifnet_flow_accum(ptr(tmp), layout, s, ptr(flow), ptr(planar), 1 if accumulate else 0,
                                              H, W, stream_ptr(self.device))
        self._check(rc, "drba_ifnet_flow_accum")

    @staticmethod
    def _int_scale(s):
        si = int(round(s))
        if si < 1 or abs(si - s) > 1e-9 or (si & (si - 1)) != 0:
            raise _lib.DrbaError(f"unsupported block scale {s}: the fused IFNet path needs power-of-two integer "
                                 f"scales (scale <= 1.0 in RIFE(...))")
        return si

    # ------------------------------------------------------------------ IFNet.forward (IFNet_HDv3.py:126-177)
    MAX_BATCH = 2   # images per persistent conv program (include/drba_b200.h: drba_conv_layer)

    def forward_multi(self, reqs, scale_list):
        """IFNet.forward for several independent requests of the same frame size, run side by side
        (they share every conv launch and the weights).  reqs: [(img0, img1, timestep, f0, f1)];
        timestep is a float or a [1,1,H,W] tensor; f0/f1 from encode() or None.  Returns the frames."""
        outs = []
        for i in range(0, len(reqs), self.MAX_BATCH):
            outs += self._forward_batch(reqs[i:i + self.MAX_BATCH], scale_list)
        return outs

    def _forward_batch(self, reqs, scale_list):
        jobs = []
        H = W = None
        for k, (img0, img1, timestep, f0, f1) in enumerate(reqs):
            require_cuda(img0, img1)
            img0, img1 = img0.float().contiguous(), img1.float().contiguous()
            if H is None:
                _, _, H, W = img0.shape
            elif tuple(img0.shape[2:]) != (H, W):
                raise _lib.DrbaError("forward_multi needs frames of one size")
            ts_t, ts_s = (timestep.float().contiguous(), 0.0) if torch.is_tensor(timestep) else (None, float(timestep))
            jobs.append({"img0": img0, "img1": img1, "ts_t": ts_t, "ts_s": ts_s, "f0": f0, "f1": f1, "prev": None, "flow": None})
        with torch.cuda.device(self.device):
            for k, j in enumerate(jobs):
                j["f0"] = self.encode(j["img0"]) if j["f0"] is None else j["f0"]
                j["f1"] = self.encode(j["img1"]) if j["f1"] is None else j["f1"]
                j["flow"] = self._buf(("flow", H, W, k), (H, W, 4))
            # tensor-core engine: blocks 1 and 2 sample 1/16 and 1/4 of the pixels, so they evaluate the flow (a sum of
            # up-sampled lastconv outputs) at their own sample positions; the full-resolution flow state is first
            # written -- in one pass, three terms -- before block 3 (same sums, same order: IFNet_HDv3.py:157)
            lazy = self.flow_terms and self.precision == "fp16"
            for bi in range(5):
                if bi > 0:     # flow (+)= s * up(previous lastconv[0:4])
                    for j in jobs:
                        if not lazy:
                            self._flow_accum(j["prev"], j["flow"], None, bi > 1, H, W)
                        elif bi <= 2:
                            j["terms"] = list(j["tmps"])
                        elif bi == 3:
                            j["terms"] = None
                            self._flow_sum(j["tmps"], j["flow"], H, W)
                        else:
                            self._flow_accum(j["prev"], j["flow"], None, True, H, W)
                prevs = self._block(bi, jobs, H, W, self._int_scale(scale_list[bi]))
                for j, p in zip(jobs, prevs):
                    j["prev"] = p
                    j.setdefault("tmps", []).append(p)
            outs = []
            for j in jobs:
                out = torch.empty((1, 3, H, W), dtype=torch.float32, device=self.device)
                tmp, layout, s = j["prev"]
                with self._launch("ifnet_blend", nbytes=float(H * W * (16 + 24 + 12))):
                    rc = self.L.drba_ifnet_blend(ptr(j["img0"]), ptr(j["img1"]), ptr(j["flow"]), ptr(tmp), layout, s, ptr(out), H, W,
                                                 stream_ptr(self.device))
                self._check(rc, "drba_ifnet_blend")
                outs.append(out)
        return outs

    def forward(self, img0, img1, timestep, scale_list, f0=None, f1=None):
        """img0, img1 [1,3,H,W] fp32; timestep float or [1,1,H,W] tensor; f0/f1 from encode().
        Returns the interpolated frame [1,3,H,W] fp32 (merged[4] of the reference)."""
        return self._forward_batch([(img0, img1, timestep, f0, f1)], scale_list)[0]

    def block0_flow(self, img0, img1, f0, f1, timestep, scale):
        """ifnet.block0(cat(a, b, f0, f1, timestep), None, scale)[0] (models/rife.py:45-46): [1,4,H,W]."""
        require_cuda(img0, img1)
        img0, img1 = img0.float().contiguous(), img1.float().contiguous()
        _, _, H, W = img0.shape
        with torch.cuda.device(self.device):
            job = {"img0": img0, "img1": img1, "ts_t": None, "ts_s": float(timestep), "f0": f0, "f1": f1, "prev": None, "flow": None}
            prev = self._block(0, [job], H, W, self._int_scale(scale))[0]
            flow = torch.empty((1, 4, H, W), dtype=torch.float32, device=self.device)
            self._flow_accum(prev, None, flow, False, H, W)
        return flow
