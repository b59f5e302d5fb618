"""GMFSS_UNION wrapper -- drop-in mirror of models/gmfss_union.py (:10-100) and models/model_gmfss_union/GMFSS.py
(Model.reuse :55-79, Model.inference :81-155).

Union = GMFSS whose GridNet sees (I1t, rife, I2t) where `rife` is a RIFE-4.26-heavy frame computed at half
resolution with the auxiliary DRM map as per-pixel timestep, plus the timestep-alignment masks of
model_gmfss_union/GMFSS.py:118-152 and a MetricNet with a tanh * 10 head.  Everything runs in libdrba_b200.so:
GMFlow (gmflow.py), FeatureNet / MetricNet / GridNet (gmfss_nets.py), the list-based splats, IFNet (ifnet.py),
calc_drm_gmfss / calc_drm_rife_auxiliary (drm.py).
"""
import os

import torch

from . import _lib
from .drm import calc_drm_gmfss, calc_drm_rife_auxiliary
from ._graphs import WindowGraphs
from .gmfss import Model
from .ifnet import IFNetEngine
from .ops import resize_bilinear
from .tools import resize
from .weights import ifnet_param_shapes, load_gmfss_state


class UnionModel(Model):
    def __init__(self, state, device, flow_estimator=None):
        super().__init__(state, device, flow_estimator, union=True)

    # models/model_gmfss_union/GMFSS.py:81
    def inference(self, img0, img1, reuse_things, timestep0, timestep1, rife, enable_mask=True):
        align = bool(torch.is_tensor(timestep0) and enable_mask)
        return self._inference(img0, img1, reuse_things, timestep0, timestep1, rife, align)


def load_union_rife_state(weights_dir):
    """rife.pkl as models/gmfss_union.py:18-20 loads it (convert(): 'module.' prefix, strict=False)."""
    raw = torch.load(os.path.join(weights_dir, "rife.pkl"), map_location="cpu")
    want = dict(ifnet_param_shapes())
    state = {k.replace("module.", ""): v.detach().float().contiguous() for k, v in raw.items() if "module." in k}
    return {k: v for k, v in state.items() if k in want}


class GMFSS_UNION:
    def __init__(self, weights='weights/train_log_gmfss_union', scale=1.0, device=None, state=None, rife_state=None,
                 flow_estimator=None, precision="fp16", graphs=None, output_dtype=torch.float16):
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DrbaError("drba_b200.GMFSS_UNION runs on a CUDA device only; there is no CPU fallback")
        if state is None:
            if not os.path.isfile(os.path.join(weights, 'fusionnet.pkl')):
                raise FileNotFoundError(os.path.join(weights, 'fusionnet.pkl'))
            state = load_gmfss_state(weights)
            rife_state = load_union_rife_state(weights)
        self.model = UnionModel(state, device, flow_estimator)
        self.output_dtype = output_dtype      # the reference returns the autocast dtype (fp16 on CUDA)
        self.ifnet = IFNetEngine(rife_state, device, precision)
        self.scale = scale
        self.scale_list = [16 / self.scale, 8 / self.scale, 4 / self.scale, 2 / self.scale, 1 / self.scale]
        self.pad_size = 128
        # graphs: every distinct window shape is captured into a CUDA graph once and replayed (_graphs.py).  Default: on
        # with the native GMFlow; off with an injected flow_estimator (arbitrary Python, may synchronise)
        if graphs is None:
            graphs = flow_estimator is None
        self._windows = WindowGraphs(self._drba_eager, device) if graphs else None

    @torch.inference_mode()
    def inference_ts(self, I0, I1, ts):
        """models/gmfss_union.py:25-43."""
        reuse = self.model.reuse(I0, I1, self.scale)
        output = []
        for t in ts:
            if t == 0:
                output.append(I0)
            elif t == 1:
                output.append(I1)
            else:
                I0s = resize_bilinear(I0, scale_factor=0.5)
                I1s = resize_bilinear(I1, scale_factor=0.5)
                rife = self.ifnet.forward(I0s, I1s, float(t), self.scale_list)
                output.append(self.model.inference(I0, I1, reuse, timestep0=t, timestep1=1 - t, rife=rife).to(self.output_dtype))
        return output

    @torch.inference_mode()
    def shard_reuse(self, Ia, Ib):
        """`reuse` as the sequential loop leaves it after the window that ends on (Ia, Ib): Model.reuse(Ia, Ib) with
        its pairs swapped (last line of inference_ts_drba) -- lets a frame-window shard start mid-stream (driver.py)."""
        r = self.model.reuse(Ia, Ib, self.scale)
        return [value for pair in zip(r[1::2], r[0::2]) for value in pair]

    @torch.inference_mode()
    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        """models/gmfss_union.py:45-100."""
        if self._windows is not None:
            return self._windows(I0, I1, I2, ts, reuse, linear)
        return self._drba_eager(I0, I1, I2, ts, reuse, linear)

    def _drba_eager(self, I0, I1, I2, ts, reuse=None, linear=False):
        reuseI1I0 = self.model.reuse(I1, I0, self.scale) if reuse is None else reuse
        reuseI1I2 = self.model.reuse(I1, I2, self.scale)
        flow10, metric10 = reuseI1I0[0], reuseI1I0[2]
        flow12, metric12 = reuseI1I2[0], reuseI1I2[2]
        I0s, I1s, I2s = [resize_bilinear(x, scale_factor=0.5) for x in [I0, I1, I2]]
        output = []
        for t in ts:
            if t == 0:
                output.append(I0)
            elif t == 1:
                output.append(I1)
            elif t == 2:
                output.append(I2)
            elif 0 < t < 1:
                t = 1 - t
                drm_gmfss = calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear)
                drm_rife = calc_drm_rife_auxiliary(t, flow10, flow12, metric10, metric12, linear, only='drm_t1_t01')
                tmap = resize(drm_rife['drm_t1_t01'], I0s.shape[2:])
                rife = self.ifnet.forward(I1s, I0s, tmap, self.scale_list)
                output.append(self.model.inference(I1, I0, reuseI1I0, timestep0=drm_gmfss['drm1t_t01'],
                                                   timestep1=drm_gmfss['drm0t_t01'], rife=rife).to(self.output_dtype))
            elif 1 < t < 2:
                t = t - 1
                drm_gmfss = calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear)
                drm_rife = calc_drm_rife_auxiliary(t, flow10, flow12, metric10, metric12, linear, only='drm_t1_t12')
                tmap = resize(drm_rife['drm_t1_t12'], I0s.shape[2:])
                rife = self.ifnet.forward(I1s, I2s, tmap, self.scale_list)
                output.append(self.model.inference(I1, I2, reuseI1I2, timestep0=drm_gmfss['drm1t_t12'],
                                                   timestep1=drm_gmfss['drm2t_t12'], rife=rife).to(self.output_dtype))
        # next reuseI1I0 = reverse(current reuseI1I2)
        reuse = [value for pair in zip(reuseI1I2[1::2], reuseI1I2[0::2]) for value in pair]
        return output, reuse
