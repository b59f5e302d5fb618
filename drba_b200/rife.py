"""RIFE wrapper -- drop-in mirror of models/rife.py (class RIFE, :15-109).

Same constructor arguments, public attributes (``scale``, ``scale_list``, ``pad_size``) and
methods (``inference_ts``, ``calc_flow``, ``inference_ts_drba``) with the reference's argument
meaning and return structure; the work runs in libdrba_b200.so through ``IFNetEngine``.

Differences a caller can observe (DESIGN.md "boundary"):
* the per-frame feature maps handed around in ``reuse`` / returned by ``calc_flow`` have the
  reference's shape [1, 16, H, W] but are channels-last VIEWS of the engine's [H, W, 16] buffers
  (fp16 with the default engine, like the reference's features under autocast); feature maps
  passed back in may have either layout (a contiguous NCHW tensor is repacked);
* ``precision='fp16'`` (default: what the reference does on CUDA under torch.autocast, rife.py:26,
  :78) runs the convolutions on the tensor cores with fp16 operands; ``precision='fp32'`` is the
  exact engine (CUDA cores, fp32 everywhere); flows, DRM maps, warps and splats are fp32 in both;
* only the DRM map the caller consumes is computed (rife.py:99 / :105 use one of the two);
* ``graphs=True`` (default): each distinct window shape (frame size, timestamp list, with/without
  ``reuse``) is captured once into a CUDA graph and replayed, which removes the ~150 kernel-launch
  gaps per window.  Replays read static input buffers (inputs are copied in) and the returned
  frames are copies of the graph's output buffers; the returned ``reuse`` tuple aliases graph
  buffers that are valid for the NEXT call (infer.py consumes it in the very next window; graphs
  share one memory pool).  Window shapes whose input ADDRESSES recur (a frame ring, and ``reuse``
  coming straight from the previous graph) are replayed from graphs captured on those addresses:
  no input copies at all.
"""
import os

import torch

from . import _lib
from .drm import calc_drm_rife
from .ifnet import IFNetEngine
from .ops import rife_invert_flow
from .weights import load_ifnet_state


def _nhwc(f):
    """Feature map in the engine's layout [H, W, 16]: a [1, 16, H, W] tensor (reference layout) is viewed -- or, if
    it is not channels-last in memory, repacked -- as [H, W, 16]."""
    if f.dim() == 4:
        f = f[0].permute(1, 2, 0)
    return f if f.is_contiguous() else f.contiguous()


def _nchw_view(f):
    """[H, W, 16] engine buffer -> the reference's [1, 16, H, W] shape (a view, no copy)."""
    return f.permute(2, 0, 1).unsqueeze(0) if f.dim() == 3 else f


class RIFE:
    def __init__(self, weights='weights/train_log_rife_426_heavy', scale=1.0,
                 device=None, precision="fp16", state=None, graphs=True):
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
        self.graphs = bool(graphs)
        self.address_graphs = os.environ.get("DRBA_ADDRESS_GRAPHS", "1") != "0"
        self._graphs = {}
        self._agraphs, self._aseen, self._pool = {}, {}, None
        self._ause, self._atick = {}, 0       # last replay tick per address graph (least-recently-used eviction)
        self.captures = 0          # graphs captured so far (a driver can warm up until this stops growing)
        self._capture_stream = None

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
        f0 = self.ifnet.encode(a3) if f0 is None else _nhwc(f0)
        f1 = self.ifnet.encode(b3) if f1 is None else _nhwc(f1)
        flow = self.ifnet.block0_flow(a3, b3, f0, f1, 0.5, self.scale_list[0])
        # both directions in one call: the planar [1, 4, H, W] flow is a batch of two [2, H, W] fields
        inv = rife_invert_flow(flow.view(2, 2, flow.shape[2], flow.shape[3]))
        return inv[0:1], inv[1:2], _nchw_view(f0), _nchw_view(f1)

    @torch.inference_mode()
    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        """models/rife.py:77-109."""
        if self.graphs:
            return self._drba_graphed(I0, I1, I2, ts, reuse, linear)
        return self._drba_eager(I0, I1, I2, ts, reuse, linear)

    # ---- address-keyed graphs ----------------------------------------------------------------------------------
    # The generic graph of a window shape reads static input buffers: every window copies three frames (75 MB at
    # 1080p) and the `reuse` tuple (167 MB) into them.  A driver loop hands the same few FRAME buffers round and round
    # (a frame ring), so a window whose (shape, timestamps, frame ADDRESSES) combination has been seen before gets a
    # graph captured directly on those frame tensors and on the `reuse` tensors of that moment (the previous graph's
    # own outputs): a replay reads whatever the frames hold now, and `reuse` is copied only when the caller's tensors
    # are not the captured ones.  In a cyclic driver every graph but one per cycle follows the graph it followed at
    # capture time, so one window per cycle pays the 167 MB copy and the others none.
    # (Round 2, first version: the key also held the `reuse` addresses.  Every new graph owns new output buffers, so
    # the next window's key was new again: the table never closed, filled up with transient graphs, and a second
    # caller on the same model -- bench.py's end-to-end loop -- ran on the generic graph: 2.22 instead of 2.08 ms.)
    MAX_ADDRESS_GRAPHS = 64

    def _address_key(self, key, frames, reuse):
        for f in frames:
            if f.dtype != torch.float32 or not f.is_contiguous():
                return None
        k = key + tuple(f.data_ptr() for f in frames)
        if reuse:
            for r in reuse:
                if not torch.is_tensor(r):
                    return None
                k += (r.dtype, tuple(r.shape), tuple(r.stride()))
        return k

    @staticmethod
    def _copy_reuse(s_reuse, reuse):
        """reuse -> the buffers a graph reads.  `reuse` may alias those buffers (the f1 of the previous window IS one of
        them and belongs in another now): copies that read a static buffer go first, from a snapshot when more than one
        of them could chain."""
        static = {d.data_ptr() for d in s_reuse}
        pending = [(d, s) for d, s in zip(s_reuse, reuse) if d.data_ptr() != s.data_ptr()]
        first = [(d, s) for d, s in pending if s.data_ptr() in static]
        if len(first) > 1:
            first = [(d, s.clone()) for d, s in first]
        for d, s in first:
            d.copy_(s)
        for d, s in pending:
            if s.data_ptr() not in static:
                d.copy_(s)

    def _drba_graphed(self, I0, I1, I2, ts, reuse, linear):
        ts_key = tuple(float(t) for t in ts)
        frames = (I0, I1, I2)
        for f in frames:
            if not f.is_cuda:
                raise _lib.DrbaError("drba_b200.RIFE runs on CUDA tensors only (no CPU fallback)")
        key = (tuple(I0.shape), ts_key, bool(reuse), bool(linear))
        akey = self._address_key(key, frames, reuse) if self.address_graphs else None
        if akey is not None:
            hit = self._agraphs.get(akey)
            if hit is None:
                seen = self._aseen.get(akey, 0) + 1
                self._aseen[akey] = seen
                if seen >= 2:
                    if len(self._agraphs) >= self.MAX_ADDRESS_GRAPHS:
                        victim = min(self._ause, key=self._ause.get)      # least recently replayed graph out
                        self._ause.pop(victim)
                        self._agraphs.pop(victim)
                    hit = self._capture_on_addresses(akey, frames, ts, reuse, linear)
                elif len(self._aseen) > 4096:
                    self._aseen.clear()
            if hit is not None:
                graph, outs, new_reuse, passthrough, n_kernels, cap_reuse = hit
                self._atick += 1
                self._ause[akey] = self._atick
                if reuse:
                    self._copy_reuse(cap_reuse, reuse)
                graph.replay()
                _lib.count(n_kernels)
                _lib.check_async("RIFE window graph")
                return [frames[pt] if pt >= 0 else o.clone() for o, pt in zip(outs, passthrough)], new_reuse
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key, frames, ts, reuse, linear)
        s_in, s_reuse, graph, outs, new_reuse, passthrough, n_kernels = entry
        for dst, src in zip(s_in, frames):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src.float())
        if reuse:
            self._copy_reuse(s_reuse, reuse)
        graph.replay()
        _lib.count(n_kernels)      # kernels inside the replayed graph (recorded at capture)
        _lib.check_async("RIFE window graph")
        output = []
        for o, pt in zip(outs, passthrough):
            output.append(frames[pt] if pt >= 0 else o.clone())   # t in {0,1,2}: the input tensor itself (rife.py:89-94)
        return output, new_reuse

    def _capture_on_addresses(self, akey, frames, ts, reuse, linear):
        """Capture this window on the caller's own tensors (see above)."""
        dev = self.device
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=dev)
        cs = self._capture_stream
        r = tuple(reuse) if reuse else None
        cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):      # warm-up: allocates engine buffers / workspaces outside the capture
            self._drba_eager(*frames, ts, r, linear)
        torch.cuda.current_stream(dev).wait_stream(cs)
        torch.cuda.synchronize(dev)
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        k0 = _lib.KERNEL_LAUNCHES
        with torch.cuda.graph(graph, stream=cs, pool=self._pool):
            outs, new_reuse = self._drba_eager(*frames, ts, r, linear)
        n_kernels = _lib.KERNEL_LAUNCHES - k0
        _lib.count(-n_kernels)     # capture records, it does not execute
        passthrough = [next((k for k, f in enumerate(frames) if o is f), -1) for o in outs]
        hit = (graph, outs, new_reuse, passthrough, n_kernels, r)      # r: the `reuse` tensors the graph reads (kept alive)
        self._agraphs[akey] = hit
        self._atick += 1
        self._ause[akey] = self._atick
        self.captures += 1
        return hit

    def _capture(self, key, frames, ts, reuse, linear):
        dev = self.device
        s_in = [torch.empty_like(f, dtype=torch.float32, memory_format=torch.contiguous_format) for f in frames]
        s_reuse = [torch.empty_like(r) for r in reuse] if reuse else None
        for dst, src in zip(s_in, frames):
            dst.copy_(src)
        if reuse:
            for dst, src in zip(s_reuse, reuse):
                dst.copy_(src)
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=dev)
        cs = self._capture_stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):      # warm-up: allocates engine buffers / workspaces outside the capture
            self._drba_eager(*s_in, ts, tuple(s_reuse) if reuse else None, linear)
        torch.cuda.current_stream(dev).wait_stream(cs)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        k0 = _lib.KERNEL_LAUNCHES
        with torch.cuda.graph(graph, stream=cs):
            outs, new_reuse = self._drba_eager(*s_in, ts, tuple(s_reuse) if reuse else None, linear)
        n_kernels = _lib.KERNEL_LAUNCHES - k0
        _lib.count(-n_kernels)     # capture records, it does not execute
        passthrough = []
        for o in outs:
            passthrough.append(next((k for k, si in enumerate(s_in) if o is si), -1))
        entry = (s_in, s_reuse, graph, outs, new_reuse, passthrough, n_kernels)
        self._graphs[key] = entry
        self.captures += 1
        return entry

    def _drba_eager(self, I0, I1, I2, ts, reuse=None, linear=False):
        flow10, flow01, f1, f0 = self.calc_flow(I1, I0) if not reuse else reuse
        if reuse is None:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2)
        else:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2, f0=reuse[2])

        # the interpolated frames of a window are independent given the flows: collect them and run
        # IFNet side by side (shared conv launches), then put the results back in `ts` order
        output, reqs, slots = list(), list(), list()
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
                reqs.append((I1, I0, drm['drm_t1_t01'], _nhwc(f1), _nhwc(f0)))
                slots.append(len(output))
                output.append(None)
            elif 1 < t < 2:
                t = t - 1
                drm = calc_drm_rife(t, flow10, flow12, linear, only='drm_t1_t12')
                reqs.append((I1, I2, drm['drm_t1_t12'], _nhwc(f1), _nhwc(f2)))
                slots.append(len(output))
                output.append(None)
        if reqs:
            for k, out in zip(slots, self.ifnet.forward_multi(reqs, self.scale_list)):
                output[k] = out

        # next flow10, flow01 = reverse(current flow12, flow21)
        return output, (flow21, flow12, f2, f1)
