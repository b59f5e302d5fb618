"""CUDA-graph capture / replay of one DRBA window (models/gmfss.py:35-73, models/gmfss_union.py:45-100).

A window is a few hundred to a few thousand kernel launches through ctypes whose sizes depend only on the frame
shape, the timestamp pattern and whether `reuse` is passed -- `WindowGraphs` captures each such combination once
and replays it.  Inputs are copied into static buffers, produced frames are copies of the graph's output buffers,
the returned `reuse` aliases graph memory that the next replay of the same graph overwrites (the caller hands it
straight back, which copies it into the static `reuse` inputs first).  drba_b200.rife.RIFE carries its own, older
copy of this logic for its flat four-tensor `reuse`."""
import torch

from . import _lib


def tensors_of(tree):
    """The tensors of a nested list / tuple, depth first."""
    if torch.is_tensor(tree):
        return [tree]
    out = []
    for x in tree:
        out.extend(tensors_of(x))
    return out


def like_tree(tree):
    """Same nesting with fresh contiguous tensors."""
    if torch.is_tensor(tree):
        return torch.empty_like(tree, memory_format=torch.contiguous_format)
    return type(tree)(like_tree(x) for x in tree)


class WindowGraphs:
    def __init__(self, eager, device):
        self.eager = eager                    # f(I0, I1, I2, ts, reuse, linear) -> (outputs, new_reuse)
        self.device = torch.device(device)
        self._graphs = {}
        self._stream = None

    def __call__(self, I0, I1, I2, ts, reuse, linear):
        frames = (I0, I1, I2)
        for f in frames:
            if not f.is_cuda:
                raise _lib.DrbaError("drba_b200 runs on CUDA tensors only (no CPU fallback)")
        key = (tuple(I0.shape), tuple(float(t) for t in ts), reuse is not None, bool(linear))
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key, frames, ts, reuse, linear)
        s_in, s_reuse, graph, outs, new_reuse, passthrough, n_kernels = entry
        for dst, src in zip(s_in, frames):
            dst.copy_(src)
        if reuse is not None:
            for dst, src in zip(tensors_of(s_reuse), tensors_of(reuse)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)
        graph.replay()
        _lib.count(n_kernels)                 # kernels inside the replayed graph (recorded at capture)
        _lib.check_async("window graph")
        output = [frames[pt] if pt >= 0 else o.clone() for o, pt in zip(outs, passthrough)]
        return output, new_reuse

    def _capture(self, key, frames, ts, reuse, linear):
        dev = self.device
        s_in = [torch.empty_like(f, dtype=torch.float32, memory_format=torch.contiguous_format) for f in frames]
        for dst, src in zip(s_in, frames):
            dst.copy_(src)
        s_reuse = None
        if reuse is not None:
            s_reuse = like_tree(reuse)
            for dst, src in zip(tensors_of(s_reuse), tensors_of(reuse)):
                dst.copy_(src)
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        cs = self._stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):           # warm-up: engine buffers / workspaces are allocated outside the capture
            self.eager(*s_in, ts, s_reuse, linear)
        torch.cuda.current_stream(dev).wait_stream(cs)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        k0 = _lib.KERNEL_LAUNCHES
        with torch.cuda.graph(graph, stream=cs):
            outs, new_reuse = self.eager(*s_in, ts, s_reuse, linear)
        n_kernels = _lib.KERNEL_LAUNCHES - k0
        _lib.count(-n_kernels)                # capture records, it does not execute
        passthrough = [next((k for k, si in enumerate(s_in) if o is si), -1) for o in outs]
        entry = (s_in, s_reuse, graph, outs, new_reuse, passthrough, n_kernels)
        self._graphs[key] = entry
        return entry
