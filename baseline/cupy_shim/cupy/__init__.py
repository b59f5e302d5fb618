"""Minimal stand-in for the slice of CuPy that the reference's softsplat uses (models/softsplat/softsplat.py:4,
:17, :24, :231-239, :362-367), so that the UNMODIFIED reference runs its own CUDA kernel on a box without cupy:

    cupy.int32 / cupy.float32                       -> numpy scalars
    cupy.memoize(for_each_device=True)              -> per-device cache decorator
    cupy.cuda.get_cuda_path()                       -> CUDA_HOME
    cupy.RawModule(code=..., options=(...)).get_function(name)(grid=, block=, args=, stream=)

Implemented with NVRTC + the CUDA driver API from cuda-python (cuda.bindings), on torch's primary context.
Only bench.py's GPU reference leg and tests put this directory on sys.path; the product never imports it
(SURVEY.md 8c-7 option ii)."""
import ctypes
import functools
import os

import numpy as np

int32 = np.int32
float32 = np.float32

__version__ = "0.0-drba-bench-shim"


def memoize(for_each_device=False):
    def deco(fn):
        cache = {}

        @functools.wraps(fn)
        def wrapper(*args):
            dev = 0
            if for_each_device:
                try:
                    import torch
                    dev = torch.cuda.current_device()
                except Exception:
                    dev = 0
            key = (dev,) + args
            if key not in cache:
                cache[key] = fn(*args)
            return cache[key]
        return wrapper
    return deco


class _Cuda:
    @staticmethod
    def get_cuda_path():
        for cand in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
            if cand and os.path.isdir(cand):
                return cand
        return None


cuda = _Cuda()


def _check(res, what):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"{what} failed: {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def compile_to_cubin_or_ptx(code, options=(), arch=None):
    """NVRTC-compile `code`; returns (image bytes, is_cubin).  Usable without a GPU when `arch` is given."""
    from cuda.bindings import nvrtc
    if arch is None:
        import torch
        major, minor = torch.cuda.get_device_capability()
        arch = f"sm_{major}{minor}"
    prog = _check(nvrtc.nvrtcCreateProgram(code.encode(), b"drba_ref_kernel.cu", 0, [], []), "nvrtcCreateProgram")
    opts = [f"--gpu-architecture={arch}".encode()]
    for o in options:
        o = o.strip()
        if o.startswith("-I "):          # the reference passes '-I <dir>' as ONE option string
            o = "-I" + o[3:].strip()
        if o:
            opts.append(o.encode())
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        size = _check(nvrtc.nvrtcGetProgramLogSize(prog), "nvrtcGetProgramLogSize")
        log = b" " * size
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC compilation failed:\n" + log.decode(errors="replace"))
    size = _check(nvrtc.nvrtcGetCUBINSize(prog), "nvrtcGetCUBINSize")
    if size:
        img = b" " * size
        _check(nvrtc.nvrtcGetCUBIN(prog, img), "nvrtcGetCUBIN")
        return img, True
    size = _check(nvrtc.nvrtcGetPTXSize(prog), "nvrtcGetPTXSize")
    img = b" " * size
    _check(nvrtc.nvrtcGetPTX(prog, img), "nvrtcGetPTX")
    return img, False


class _Function:
    def __init__(self, module, name):
        from cuda.bindings import driver
        self._driver = driver
        self._fn = _check(driver.cuModuleGetFunction(module, name.encode()), "cuModuleGetFunction")

    def __call__(self, grid, block, args, stream=None, shared_mem=0):
        driver = self._driver
        vals, types = [], []
        for a in args:
            if isinstance(a, (np.int32,)):
                vals.append(int(a)); types.append(ctypes.c_int)
            elif isinstance(a, (np.float32,)):
                vals.append(float(a)); types.append(ctypes.c_float)
            elif isinstance(a, int):       # tensor.data_ptr()
                vals.append(a); types.append(ctypes.c_void_p)
            elif isinstance(a, float):
                vals.append(a); types.append(ctypes.c_float)
            else:
                raise TypeError(f"unsupported kernel argument {type(a)}")
        sptr = getattr(stream, "ptr", 0) if stream is not None else 0
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        err, = driver.cuLaunchKernel(self._fn, grid[0], grid[1], grid[2], block[0], block[1], block[2],
                                     shared_mem, sptr, (tuple(vals), tuple(types)), 0)
        if int(err) != 0:
            raise RuntimeError(f"cuLaunchKernel failed: {err}")


class RawModule:
    def __init__(self, code=None, options=(), **_kw):
        import torch
        from cuda.bindings import driver
        torch.cuda.init()
        torch.cuda.current_stream()        # makes torch's primary context current on this thread
        img, _ = compile_to_cubin_or_ptx(code, options)
        self._module = _check(driver.cuModuleLoadData(img), "cuModuleLoadData")

    def get_function(self, name):
        return _Function(self._module, name)
