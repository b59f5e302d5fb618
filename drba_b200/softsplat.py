"""softsplat(tenIn, tenFlow, tenMetric, strMode) -- host-side mirror of the reference operator.

Same name, argument meaning and assertions as models/softsplat/softsplat.py:248-293 (CuPy
path) and its twin models/softsplat/softsplat_torch.py:20-70; the work is done by
``drba_softsplat_f32`` in libdrba_b200.so (csrc/splat.cu).  Plug-in seam: the reference's
consumers pick their ``warp`` at import time (models/drm.py:3-7, models/rife.py:8-12, ...);
INTEGRATION.md shows the third branch that selects this module.

Behavioural notes:
* inputs are upcast to fp32 and the result is cast back to ``tenIn.dtype`` (softsplat.py:251,:293);
* ``avg`` accepts and ignores a metric (the CuPy path's assert is commented out, :256);
* ``avg-zeroeps`` etc. behave as in the torch twin (ones channel appended for every ``avg-*``),
  not as in the CuPy wrapper, whose ``strMode == 'avg'`` test drops the ones channel for
  suffixed modes (:260) -- no call site in the reference uses those spellings.
"""
import torch

from . import _lib
from ._torch_util import Workspace, f32c, ptr, require_cuda, stream_ptr

MODES = {"sum": 0, "avg": 1, "linear": 2, "soft": 3}
EPS = {None: 0, "addeps": 0, "zeroeps": 1, "clipeps": 2, "?": 3}


def softsplat(tenIn, tenFlow, tenMetric, strMode: str, _variant=0):
    parts = strMode.split("-")
    assert parts[0] in ["sum", "avg", "linear", "soft"]
    if strMode == "sum":
        assert tenMetric is None
    if parts[0] == "linear":
        assert tenMetric is not None
    if parts[0] == "soft":
        assert tenMetric is not None
    sub = parts[1] if len(parts) > 1 else None
    if sub not in EPS:      # an unknown suffix matches no branch of softsplat.py:273-290: raw denominator
        sub = "?"
    mode, eps = MODES[parts[0]], EPS[sub]
    if parts[0] in ("sum", "avg"):
        tenMetric = None

    require_cuda(tenIn, tenFlow, tenMetric)
    output_dtype = tenIn.dtype
    x, flow, metric = f32c(tenIn), f32c(tenFlow), f32c(tenMetric)
    n, c, h, w = x.shape
    assert flow.shape == (n, 2, h, w), "tenFlow must be [N,2,H,W]"
    if metric is not None:
        assert metric.shape == (n, 1, h, w), "tenMetric must be [N,1,H,W]"
    out = torch.empty_like(x)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        need = L.drba_softsplat_workspace_bytes(n, c if _variant not in (3, 4) else max(c, 16), h, w, mode)
        ws = Workspace.get(need, x.device)
        group_bytes = n * h * w * 16
        has_w = 0 if mode == 0 else 1
        cc_max = max(1, (need // group_bytes) * 4 - has_w) if group_bytes else 1
        chunks = max(1, -(-c // cc_max))
        with _lib.launch("softsplat", 2 * chunks,
                         nbytes=4.0 * n * h * w * ((c + 2 + (1 if metric is not None else 0)) + c)):
            rc = L.drba_softsplat_f32_variant(ptr(x), ptr(flow), ptr(metric), ptr(out), n, c, h, w, mode, eps,
                                              ws.data_ptr(), need, int(_variant), stream_ptr(x.device))
    _lib.check(rc, "drba_softsplat_f32")
    return out.to(output_dtype)


warp = softsplat  # the name every reference consumer imports it under
