"""RIFE wrapper -- drop-in mirror of models/rife.py (class RIFE, :15-109).

Same constructor arguments, public attributes (``scale``, ``scale_list``, ``pad_size``) and
methods (``inference_ts``, ``calc_flow``, ``inference_ts_drba``) with the reference's argument
meaning and return structure; the work runs in libdrba_b200.so through ``IFNetEngine``.

Differences a caller can observe (DESIGN.md "boundary"):
* the per-frame feature maps handed around in ``reuse`` / returned by ``calc_flow`` are the
  engine's [H, W, 16] channels-last tensors (the reference returns [1, 16, H, W]); they are
  opaque to infer.py, which only passes them back in;
* ``precision='fp32'`` (default here) computes everything in fp32; ``precision='fp16'`` runs the
  convolutions on the tensor cores with fp16 operands like the reference under torch.autocast
  (rife.py:26, :78); flows, DRM maps, warps and splats are fp32 in both;
* only the DRM map the caller consumes is computed (rife.py:99 / :105 use one of the two).
"""
import os

import torch

from . import _lib
from .drm import calc_drm_rife
from .ifnet import IFNetEngine
from .ops import rife_invert_flow
from .weights import load_ifnet_state


class RIFE:
    def __init__(self, weights='weights/train_log_rife_426_heavy', scale=1.0,
                 device=None, precision="fp32", state=None):
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DrbaError("drba_b200.RIFE runs on a CUDA device only; there is no CPU fallback "
                                 "(use the reference implementation on CPU)")
        if state is None:
            if not os.path.isfile(os.path.join(weights, 'flownet.pkl')):
                raise FileNotFoundError(os.path.join(weights, 'flownet.pkl'))
            state = load_ifnet_state(weights)
        self.device = device
        self.ifnet = IFNetEngine(state, device, precision)
        self.scale = scale
        self.scale_list = [16 / scale, 8 / scale, 4 / scale, 2 / scale, 1 / scale]
        self.pad_size = 64

    @torch.inference_mode()
    def inference_ts(self, I0, I1, ts):
        """models/rife.py:25-39."""
        output = []
        for t in ts:
            if t == 0:
                output.append(I0)
            elif t == 1:
                output.append(I1)
            else:
                output.append(self.ifnet.forward(I0, I1, float(t), self.scale_list))
        return output

    @torch.inference_mode()
    def calc_flow(self, a, b, f0=None, f1=None):
        """models/rife.py:41-75: block0-only bidirectional flow at 1/16 scale, inverted by
        forward-warping it onto itself, holes <- max(H, W)."""
        a3, b3 = a[:, :3], b[:, :3]
        f0 = self.ifnet.encode(a3) if f0 is None else f0
        f1 = self.ifnet.encode(b3) if f1 is None else f1
        flow = self.ifnet.block0_flow(a3, b3, f0, f1, 0.5, self.scale_list[0])
        flow01 = rife_invert_flow(flow[:, :2])
        flow10 = rife_invert_flow(flow[:, 2:])
        return flow01, flow10, f0, f1

    @torch.inference_mode()
    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        """models/rife.py:77-109."""
        flow10, flow01, f1, f0 = self.calc_flow(I1, I0) if not reuse else reuse
        if reuse is None:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2)
        else:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2, f0=reuse[2])

        output = list()
        for t in ts:
            if t == 0:
                output.append(I0)
            elif t == 1:
                output.append(I1)
            elif t == 2:
                output.append(I2)
            elif 0 < t < 1:
                t = 1 - t
                drm = calc_drm_rife(t, flow10, flow12, linear, only='drm_t1_t01')
                out = self.ifnet.forward(I1, I0, drm['drm_t1_t01'], self.scale_list, f0=f1, f1=f0)
                output.append(out)
            elif 1 < t < 2:
                t = t - 1
                drm = calc_drm_rife(t, flow10, flow12, linear, only='drm_t1_t12')
                out = self.ifnet.forward(I1, I2, drm['drm_t1_t12'], self.scale_list, f0=f1, f1=f2)
                output.append(out)

        # next flow10, flow01 = reverse(current flow12, flow21)
        return output, (flow21, flow12, f2, f1)
