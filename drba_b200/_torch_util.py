"""Device-memory plumbing shared by the host-side operator mirrors (torch tensors own
every buffer; the library only sees raw pointers + the current CUDA stream)."""
import torch

from . import _lib


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DrbaError("drba_b200 operators run on CUDA tensors only (no CPU fallback); "
                                 f"got a tensor on {t.device}")


def f32c(t):
    """contiguous float32 view/copy (models/softsplat/softsplat.py:251 upcasts the same way)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Workspace:
    """Zero-invariant scratch memory per (device, stream): cleared once when (re)allocated;
    every drba_* op leaves it zero on exit (include/drba_b200.h)."""

    _pool = {}
    _retired = []      # outgrown buffers stay alive: CUDA graphs captured earlier have their addresses baked in

    @classmethod
    def get(cls, nbytes, device):
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        buf = cls._pool.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                cls._retired.append(buf)
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            _lib.check(_lib.lib().drba_workspace_clear(buf.data_ptr(), buf.numel(), stream_ptr(device)),
                       "drba_workspace_clear")
            cls._pool[key] = buf
        return buf
