"""ctypes binding of oracle/drba_oracle.c (test infrastructure, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdrba_oracle.so")

MODES = {"sum": 0, "avg": 1, "linear": 2, "soft": 3}
EPS = {None: 0, "addeps": 0, "zeroeps": 1, "clipeps": 2}
EPS_UNKNOWN = 3   # any other suffix: no branch of softsplat.py:273-290 fires, raw denominator


def build(force=False):
    src = os.path.join(_HERE, "drba_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ci = ctypes.c_int
        L.orc_splat_sum.argtypes = [fp, fp, fp, ci, ci, ci, ci]
        L.orc_splat_sum.restype = None
        L.orc_softsplat.argtypes = [fp, fp, fp, fp, ci, ci, ci, ci, ci, ci, fp]
        L.orc_softsplat.restype = ci
        L.orc_get_drm_t.argtypes = [fp, ctypes.c_double, ctypes.c_double, fp, ctypes.c_size_t]
        L.orc_get_drm_t.restype = None
        L.orc_drm_rife.argtypes = [ctypes.c_double, fp, fp, fp, fp, ci, fp, fp, ci, ci, ci]
        L.orc_drm_rife.restype = ci
        L.orc_drm_gmfss.argtypes = [ctypes.c_double, fp, fp, fp, fp, ci, fp, fp, fp, fp, ci, ci, ci]
        L.orc_drm_gmfss.restype = ci
        L.orc_backwarp.argtypes = [fp, fp, fp, ci, ci, ci, ci, ci]
        L.orc_backwarp.restype = None
        L.orc_resize_bilinear.argtypes = [fp, fp, ci, ci, ci, ci, ci, ci, ci, ctypes.c_float, ctypes.c_float]
        L.orc_resize_bilinear.restype = None
        L.orc_rife_invert_flow.argtypes = [fp, fp, ci, ci, ci]
        L.orc_rife_invert_flow.restype = ci
        _lib = L
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def softsplat(ten_in, flow, metric, mode):
    """numpy NCHW float32 restatement of softsplat(tenIn, tenFlow, tenMetric, strMode)."""
    parts = mode.split("-")
    m, e = MODES[parts[0]], EPS.get(parts[1] if len(parts) > 1 else None, EPS_UNKNOWN)
    x, f, mt = _f32(ten_in), _f32(flow), _f32(metric)
    if m == 0:
        assert mt is None
    if m in (2, 3):
        assert mt is not None
    if m == 1:
        mt = None
    n, c, h, w = x.shape
    out = np.empty_like(x)
    rc = lib().orc_softsplat(_p(x), _p(f), _p(mt), _p(out), n, c, h, w, m, e, None)
    assert rc == 0, rc
    return out


def get_drm_t(drm, t, precision=1e-3):
    d = _f32(drm)
    out = np.empty_like(d)
    lib().orc_get_drm_t(_p(d), float(t), float(precision), _p(out), d.size)
    return out


def calc_drm_rife(t, flow10, flow12, linear=False, metric10=None, metric12=None):
    f10, f12, m10, m12 = _f32(flow10), _f32(flow12), _f32(metric10), _f32(metric12)
    n, _, h, w = f10.shape
    a = np.empty((n, 1, h, w), np.float32)
    b = np.empty((n, 1, h, w), np.float32)
    rc = lib().orc_drm_rife(float(t), _p(f10), _p(f12), _p(m10), _p(m12), int(bool(linear)), _p(a), _p(b), n, h, w)
    assert rc == 0, rc
    return {"drm_t1_t01": a, "drm_t1_t12": b}


def calc_drm_rife_auxiliary(t, flow10, flow12, metric10, metric12, linear=False):
    return calc_drm_rife(t, flow10, flow12, linear, metric10, metric12)


def calc_drm_gmfss(t, flow10, flow12, metric10, metric12, linear=False):
    f10, f12, m10, m12 = _f32(flow10), _f32(flow12), _f32(metric10), _f32(metric12)
    n, _, h, w = f10.shape
    outs = [np.empty((n, 1, h, w), np.float32) for _ in range(4)]
    rc = lib().orc_drm_gmfss(float(t), _p(f10), _p(f12), _p(m10), _p(m12), int(bool(linear)),
                             *[_p(o) for o in outs], n, h, w)
    assert rc == 0, rc
    return dict(zip(["drm0t_t01", "drm1t_t01", "drm1t_t12", "drm2t_t12"], outs))


def backwarp(ten_in, flow, padding="border"):
    x, f = _f32(ten_in), _f32(flow)
    n, c, h, w = x.shape
    out = np.empty_like(x)
    pad = 0 if padding == "border" else 1
    nthr = min(c, os.cpu_count() or 1)
    if n != 1 or nthr < 2 or h * w < 65536:
        lib().orc_backwarp(_p(x), _p(f), _p(out), n, c, h, w, pad)
        return out
    # channels are independent: run channel slices on host threads (ctypes drops the GIL);
    # the arithmetic per element is unchanged
    from concurrent.futures import ThreadPoolExecutor
    L = lib()
    bounds = np.linspace(0, c, nthr + 1).astype(int)

    def run(i):
        a, b = int(bounds[i]), int(bounds[i + 1])
        if b > a:
            L.orc_backwarp(_p(x[:, a:b]), _p(f), _p(out[:, a:b]), 1, b - a, h, w, pad)
    with ThreadPoolExecutor(nthr) as ex:
        list(ex.map(run, range(nthr)))
    return out


def resize_bilinear(x, size=None, scale_factor=None, align_corners=False):
    x = _f32(x)
    n, c, h, w = x.shape
    if size is not None:
        oh, ow = size
        rh, rw = h / oh, w / ow
    else:
        oh, ow = int(np.floor(h * scale_factor)), int(np.floor(w * scale_factor))
        rh = rw = 1.0 / scale_factor
    out = np.empty((n, c, oh, ow), np.float32)
    lib().orc_resize_bilinear(_p(x), _p(out), n, c, h, w, oh, ow, int(align_corners), rh, rw)
    return out


def rife_invert_flow(flow_t0):
    f = _f32(flow_t0)
    n, _, h, w = f.shape
    out = np.empty_like(f)
    rc = lib().orc_rife_invert_flow(_p(f), _p(out), n, h, w)
    assert rc == 0, rc
    return out
