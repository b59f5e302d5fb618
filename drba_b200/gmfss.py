"""GMFSS wrapper -- drop-in mirror of models/gmfss.py (class GMFSS, :8-73) and models/model_gmfss/GMFSS.py
(class Model: reuse :58-83, inference :85-190).

Built natively here: FeatureNet, MetricNet (with its warp / consistency-check
front end), the per-scale flow and metric scaling, all forward splats (one list build per (flow, metric,
scale), applied to the image and to the NHWC fp16 feature map, written straight into GridNet's concat
buffers with the head PReLU fused), GridNet, calc_drm_gmfss.

GMFlow (row a-9) is `drba_b200.gmflow.GMFlow` (tcgen05 conv / batched-GEMM programs + csrc/gmflow.cu); it is built
from `flownet.pkl` (state["flownet"]).  A different `flow_estimator(img_a, img_b) -> flow_ab [1,2,h,w]` can be
injected (the component parity tests inject the reference GMFlow's flows from tests/golden); with neither,
`Model.reuse` raises DrbaError -- there is no silent fallback.

Differences a caller can observe: the entries of `reuse` holding features are tuples of NHWC fp16 tensors (opaque to infer.py).
"""
import os

import torch

from . import _lib
from ._graphs import WindowGraphs
from ._torch_util import Workspace, ptr, require_cuda, stream_ptr
from .drm import calc_drm_gmfss
from .gmfss_nets import FeatureNet, GridNet, MetricNet, pack_planes
from .ops import resize_bilinear
from .weights import load_gmfss_state

SOFT = 3


def _splat_lists_build(flow, metric, h, w, device):
    L = _lib.lib()
    need = L.drba_splat_lists_workspace_bytes(h, w)
    ws = Workspace.get(need, device)
    with _lib.launch("splat_lists_build", 6, nbytes=float(h * w * (12 + 100))):
        rc = L.drba_splat_lists_build(ptr(flow), ptr(metric), SOFT, h, w, ws.data_ptr(), need, stream_ptr(device))
    _lib.check(rc, "drba_splat_lists_build")
    return ws


class Model:
    def __init__(self, state, device, flow_estimator=None, union=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DrbaError("drba_b200 GMFSS runs on a CUDA device only; there is no CPU fallback")
        self.feat_ext = FeatureNet(state["feat"], self.device)
        self.metricnet = MetricNet(state["metric"], self.device, union=union)
        self.fusionnet = GridNet(state["fusionnet"], self.device)
        if flow_estimator is None and "flownet" in state:
            from .gmflow import GMFlow
            flow_estimator = GMFlow(state["flownet"], self.device)      # models/model_gmfss/GMFSS.py:21 self.flownet
        self.flow_estimator = flow_estimator
        self.version = 3.9

    def _scaled(self, x, alpha):
        out = torch.empty_like(x)
        with _lib.launch("axpby", 1, nbytes=8.0 * x.numel()):
            rc = _lib.lib().drba_axpby_f32(ptr(x), float(alpha), None, 0.0, ptr(out), x.numel(), stream_ptr(self.device))
        _lib.check(rc, "drba_axpby_f32")
        return out

    # models/model_gmfss/GMFSS.py:58-83
    def reuse(self, img0, img1, scale):
        require_cuda(img0, img1)
        if self.flow_estimator is None:
            raise _lib.DrbaError("no GMFlow weights (state['flownet'] / flownet.pkl) and no flow_estimator=callable(img_a, img_b) "
                                 "-> flow_ab were given; there is no fallback")
        feat_ext0, feat_ext1 = self.feat_ext([img0, img1])
        img0h = resize_bilinear(img0, scale_factor=0.5)
        img1h = resize_bilinear(img1, scale_factor=0.5)
        if scale != 1.0:
            imgf0 = resize_bilinear(img0h, scale_factor=scale)
            imgf1 = resize_bilinear(img1h, scale_factor=scale)
        else:
            imgf0, imgf1 = img0h, img1h
        if hasattr(self.flow_estimator, "pair"):      # native GMFlow: both directions share backbone + 1/8-scale transformer
            flow01, flow10 = self.flow_estimator.pair(imgf0, imgf1)
        else:
            flow01 = self.flow_estimator(imgf0, imgf1)
            flow10 = self.flow_estimator(imgf1, imgf0)
        if scale != 1.0:
            flow01 = self._scaled(resize_bilinear(flow01, scale_factor=1. / scale), 1. / scale)
            flow10 = self._scaled(resize_bilinear(flow10, scale_factor=1. / scale), 1. / scale)
        metric0, metric1 = self.metricnet(img0h, img1h, flow01, flow10)
        return flow01, flow10, metric0, metric1, feat_ext0, feat_ext1

    # models/model_gmfss/GMFSS.py:85-190
    def inference(self, img0, img1, reuse_things, timestep0, timestep1, swap_thresh=1):
        return self._inference(img0, img1, reuse_things, timestep0, timestep1, None, False)

    def _inference(self, img0, img1, reuse_things, timestep0, timestep1, rife, align_timesteps):
        """Shared body of Model.inference (gmfss: GMFSS.py:85-190; union: model_gmfss_union/GMFSS.py:81-155).
        rife: None (gmfss: GridNet sees img0, I1t, I2t, img1) or the RIFE frame [1,3,h,w] (union: I1t, rife, I2t).
        align_timesteps: union's timestep warping / hole filling / ratio > 25 swaps (:118-152)."""
        require_cuda(img0, img1)
        flow01, flow10, metric0, metric1, feats0, feats1 = reuse_things
        dev = self.device
        L = _lib.lib()
        img0h = resize_bilinear(img0.float(), scale_factor=0.5)
        img1h = resize_bilinear(img1.float(), scale_factor=0.5)
        _, _, h, w = img0h.shape
        assert h % 4 == 0 and w % 4 == 0
        B = self.fusionnet.bufs
        sl = self.fusionnet.head_slopes
        x1 = B.get(("in1", h, w), (h, w, 128))
        x2 = B.get(("in2", h, w), (h // 2, w // 2, 256))
        x3 = B.get(("in3", h, w), (h // 4, w // 4, 384))
        warped, tw, gaps = [], [], []
        with torch.cuda.device(dev):
            ones = torch.ones((1, 1, h, w), dtype=torch.float32, device=dev) if align_timesteps else None
            for side, (img_h, flow, metric, feats, t) in enumerate(((img0h, flow01, metric0, feats0, timestep0),
                                                                     (img1h, flow10, metric1, feats1, timestep1))):
                flow, metric = flow.float().contiguous(), metric.float().contiguous()
                tmap, tscalar = (t.float().contiguous(), 0.0) if torch.is_tensor(t) else (None, float(t))
                for si, (s, feat, xbuf, slope) in enumerate(((1, feats[0], x1, sl[1]), (2, feats[1], x2, sl[2]), (4, feats[2], x3, sl[3]))):
                    hs, ws_ = h // s, w // s
                    F = torch.empty((1, 2, hs, ws_), dtype=torch.float32, device=dev)
                    Z = torch.empty((1, 1, hs, ws_), dtype=torch.float32, device=dev)
                    with _lib.launch("gmfss_scale_flow", 1, nbytes=float(h * w * 16 / (s * s) * 4 + hs * ws_ * 12)):
                        rc = L.drba_gmfss_scale_flow(ptr(flow), ptr(metric), ptr(tmap), tscalar, h, w, s, ptr(F), ptr(Z), stream_ptr(dev))
                    _lib.check(rc, "drba_gmfss_scale_flow")
                    lists = _splat_lists_build(F, Z, hs, ws_, dev)
                    C = feat.shape[2]
                    if s == 1:      # I1t / I2t: the half-resolution image (GMFSS.py:96-98)
                        It = torch.empty_like(img_h)
                        with _lib.launch("splat_apply_nchw", 1, nbytes=float(hs * ws_ * (24 + 40))):
                            rc = L.drba_splat_lists_apply_nchw_f32(lists.data_ptr(), ptr(img_h), ptr(It), 3, hs, ws_, 1, 0, stream_ptr(dev))
                        _lib.check(rc, "drba_splat_lists_apply_nchw_f32")
                        warped.append(It)
                        if align_timesteps:     # union :118-124: the timestep map and a ones map ride the same lists
                            tws, g = torch.empty_like(tmap), torch.empty_like(tmap)
                            with _lib.launch("splat_apply_nchw", 2, nbytes=float(hs * ws_ * 2 * (8 + 40))):
                                rc = L.drba_splat_lists_apply_nchw_f32(lists.data_ptr(), ptr(tmap), ptr(tws), 1, hs, ws_, 1, 0, stream_ptr(dev))
                                rc = rc or L.drba_splat_lists_apply_nchw_f32(lists.data_ptr(), ptr(ones), ptr(g), 1, hs, ws_, 1, 0, stream_ptr(dev))
                            _lib.check(rc, "drba_splat_lists_apply_nchw_f32")
                            tw.append(tws)
                            gaps.append(g)
                    with _lib.launch("splat_apply_nhwc", 1, nbytes=float(hs * ws_ * (4 * C + 40))):
                        rc = L.drba_splat_lists_apply_nhwc_f16(lists.data_ptr(), ptr(feat), C, C, ptr(xbuf), 2 * C, side * C, hs, ws_,
                                                               1, 0, 1, float(slope), stream_ptr(dev))
                    _lib.check(rc, "drba_splat_lists_apply_nhwc_f16")
                    _lib.check(L.drba_splat_lists_release(lists.data_ptr(), hs, ws_, stream_ptr(dev)), "drba_splat_lists_release")
            if align_timesteps:
                t0w, t1w = tw
                with _lib.launch("gmfss_union_masks", 5):
                    rc = L.drba_gmfss_union_fix_timesteps(ptr(t0w), ptr(t1w), ptr(gaps[0]), ptr(gaps[1]), t0w.numel(), stream_ptr(dev))
                    rc = rc or L.drba_gmfss_union_swap_nchw_f32(ptr(warped[0]), ptr(warped[1]), 3, ptr(t0w), ptr(t1w), h, w, stream_ptr(dev))
                    rc = rc or L.drba_gmfss_union_swap_nhwc_f16(ptr(x1), 64, ptr(t0w), ptr(t1w), h, w, stream_ptr(dev))
                _lib.check(rc, "drba_gmfss_union_*")
                for sc, xbuf, C in ((0.5, x2, 128), (0.25, x3, 192)):
                    a, b = resize_bilinear(t0w, scale_factor=sc), resize_bilinear(t1w, scale_factor=sc)
                    with _lib.launch("gmfss_union_masks", 1):
                        rc = L.drba_gmfss_union_swap_nhwc_f16(ptr(xbuf), C, ptr(a), ptr(b), a.shape[2], a.shape[3], stream_ptr(dev))
                    _lib.check(rc, "drba_gmfss_union_swap_nhwc_f16")
            if rife is None:
                planes = list(img0h[0]) + list(warped[0][0]) + list(warped[1][0]) + list(img1h[0])
            else:
                planes = list(warped[0][0]) + list(rife.float().contiguous()[0]) + list(warped[1][0])
            x = pack_planes(planes, B.get(("in0", h, w), (h, w, 16)), prelu=sl[0])
            return self.fusionnet(x, x1, x2, x3)


class GMFSS:
    def __init__(self, weights=r'weights/train_log_gmfss', scale=1.0, device=None, state=None, flow_estimator=None, graphs=None,
                 output_dtype=torch.float16):
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DrbaError("drba_b200.GMFSS runs on a CUDA device only; there is no CPU fallback")
        if state is None:
            if not os.path.isfile(os.path.join(weights, 'fusionnet.pkl')):
                raise FileNotFoundError(os.path.join(weights, 'fusionnet.pkl'))
            state = load_gmfss_state(weights)
        self.model = Model(state, device, flow_estimator)
        # the reference returns frames in the autocast dtype (fp16 on CUDA); torch.float32 keeps GridNet's
        # unpacked fp32 output as it is
        self.output_dtype = output_dtype
        self.scale = scale
        self.pad_size = 64
        # graphs: every distinct window shape is captured into a CUDA graph once and replayed (_graphs.py).  Default: on
        # with the native GMFlow; off with an injected flow_estimator (arbitrary Python, may synchronise)
        if graphs is None:
            graphs = flow_estimator is None
        self._windows = WindowGraphs(self._drba_eager, device) if graphs else None

    @torch.inference_mode()
    def inference_ts(self, I0, I1, ts):
        """models/gmfss.py:17-33."""
        reuse = self.model.reuse(I0, I1, self.scale)
        output = []
        for t in ts:
            if t == 0:
                output.append(I0)
            elif t == 1:
                output.append(I1)
            else:
                output.append(self.model.inference(I0, I1, reuse, timestep0=t, timestep1=1 - t).to(self.output_dtype))
        return output

    @torch.inference_mode()
    def shard_reuse(self, Ia, Ib):
        """`reuse` as the sequential loop leaves it after the window that ends on (Ia, Ib): Model.reuse(Ia, Ib) with
        its pairs swapped (last line of inference_ts_drba) -- lets a frame-window shard start mid-stream (driver.py)."""
        r = self.model.reuse(Ia, Ib, self.scale)
        return [value for pair in zip(r[1::2], r[0::2]) for value in pair]

    @torch.inference_mode()
    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        """models/gmfss.py:35-73."""
        if self._windows is not None:
            return self._windows(I0, I1, I2, ts, reuse, linear)
        return self._drba_eager(I0, I1, I2, ts, reuse, linear)

    def _drba_eager(self, I0, I1, I2, ts, reuse=None, linear=False):
        reuseI1I0 = self.model.reuse(I1, I0, self.scale) if reuse is None else reuse
        reuseI1I2 = self.model.reuse(I1, I2, self.scale)
        flow10, metric10 = reuseI1I0[0], reuseI1I0[2]
        flow12, metric12 = reuseI1I2[0], reuseI1I2[2]
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
                drm = calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear)
                output.append(self.model.inference(I1, I0, reuseI1I0, timestep0=drm['drm1t_t01'], timestep1=drm['drm0t_t01']).to(self.output_dtype))
            elif 1 < t < 2:
                t = t - 1
                drm = calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear)
                output.append(self.model.inference(I1, I2, reuseI1I2, timestep0=drm['drm1t_t12'], timestep1=drm['drm2t_t12']).to(self.output_dtype))
        # next reuseI1I0 = reverse(current reuseI1I2)
        reuse = [value for pair in zip(reuseI1I2[1::2], reuseI1I2[0::2]) for value in pair]
        return output, reuse
