"""Host-side plumbing of the persistent tensor-core conv programs (csrc/conv_tc.cu).

A program is a list of steps; every step is one conv layer applied to 1-2 independent images.  Steps are
chained inside ONE launch (grid barrier between layers), at most CONV_MAX_LAYERS per launch; longer
chains are cut into several launches.  Used by the IFNet engine (ifnet.py) and the GMFSS networks
(gmfss_nets.py)."""
import ctypes

import torch

from . import _lib
from ._torch_util import ptr, stream_ptr

ACT_NONE, ACT_LRELU, ACT_PRELU_VEC, ACT_RELU, ACT_PRELU, ACT_GELU = 0, 1, 2, 3, 4, 5


class Step:
    """One layer of a program.  ins/outs/res/...: lists with one tensor per image (or None)."""

    __slots__ = ("layer", "H", "W", "ins", "outs", "OH", "OW", "cstride", "res", "res2", "out1", "out2",
                 "act", "act1", "act2", "slope0", "slope1", "slope2", "bgemm")

    def __init__(self, layer, H, W, ins, outs, OH, OW, cstride, res=None, res2=None, out1=None, out2=None,
                 act=None, act1=0, act2=0, slope0=0.0, slope1=0.0, slope2=0.0, bgemm=0):
        self.layer, self.H, self.W, self.ins, self.outs, self.OH, self.OW, self.cstride = layer, H, W, ins, outs, OH, OW, cstride
        self.res, self.res2, self.out1, self.out2 = res, res2, out1, out2
        self.act = layer.act if act is None else act
        self.act1, self.act2, self.slope0, self.slope1, self.slope2 = act1, act2, slope0, slope1, slope2
        self.bgemm = bgemm


_sync_words = {}


def _sync(device):
    """Grid-barrier words of the persistent programs: one pair per (device, stream) -- programs on one stream run
    one after the other and may share them, programs on different streams must not."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    t = _sync_words.get(key)
    if t is None:
        t = torch.zeros(2, dtype=torch.int32, device=device)
        _sync_words[key] = t
    return t


def run_program(steps, device, tag=""):
    """Launch `steps` (Step objects or the legacy 9-tuples of ifnet.py) as persistent conv programs."""
    steps = [s if isinstance(s, Step) else Step(*s[:8], res=s[8]) for s in steps]
    L = _lib.lib()
    for i0 in range(0, len(steps), _lib.CONV_MAX_LAYERS):
        chunk = steps[i0:i0 + _lib.CONV_MAX_LAYERS]
        nimg = len(chunk[0].ins)
        arr = (_lib.ConvLayer * len(chunk))()
        flops = 0.0
        for c, s in zip(arr, chunk):
            layer = s.layer
            for k in range(nimg):
                c.in_[k] = ptr(s.ins[k])
                c.out[k] = ptr(s.outs[k])
                c.res[k] = ptr(s.res[k]) if s.res is not None else None
                c.res2[k] = ptr(s.res2[k]) if s.res2 is not None else None
                c.out1[k] = ptr(s.out1[k]) if s.out1 is not None else None
                c.out2[k] = ptr(s.out2[k]) if s.out2 is not None else None
            c.w, c.bias, c.slope = ptr(layer.w), ptr(layer.b), ptr(layer.slope)
            c.H, c.W, c.Cin, c.G, c.T = s.H, s.W, layer.cin, layer.G, layer.T
            n = layer.G * layer.T
            c.dy[:n] = layer.dy[:n]
            c.dx[:n] = layer.dx[:n]
            c.cout_pad, c.cout, c.S, c.OH, c.OW = layer.cout_pad, layer.cout, layer.stride, s.OH, s.OW
            c.epilogue, c.act, c.out_cstride, c.out_os = layer.epilogue, s.act, s.cstride, layer.out_os
            c.act1, c.act2, c.slope0, c.slope1, c.slope2 = s.act1, s.act2, s.slope0, s.slope1, s.slope2
            c.bgemm = s.bgemm
            flops += nimg * 2.0 * getattr(layer, "flop_taps", layer.G * layer.T) * layer.cin_real * layer.cout * s.OH * s.OW
        with _lib.launch("conv_tc_f16" + (("/" + tag) if tag else ""), 1, flops=flops):
            rc = L.drba_conv_tc_program_f16(ctypes.addressof(arr), len(chunk), nimg, ptr(_sync(device)), stream_ptr(device))
        _lib.check(rc, "drba_conv_tc_program_f16")
