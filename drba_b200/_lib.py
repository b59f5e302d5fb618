"""ctypes loader for libdrba_b200.so (the C ABI declared in include/drba_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the
product path raises.  ``build()`` compiles it in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libdrba_b200.so")
CSRC = os.path.join(_HERE, "csrc")

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_Z = _c.c_size_t
_D = _c.c_double
_F = _c.c_float

# name -> (restype, argtypes); must list every symbol include/drba_b200.h declares
SIGNATURES = {
    "drba_version": (_I, []),
    "drba_error_string": (_c.c_char_p, [_I]),
    "drba_workspace_clear": (_I, [_P, _Z, _P]),
    "drba_softsplat_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "drba_softsplat_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "drba_softsplat_f32_variant": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _Z, _I, _P]),
    "drba_rife_invert_flow_workspace_bytes": (_Z, [_I, _I, _I]),
    "drba_rife_invert_flow_f32": (_I, [_P, _P, _I, _I, _I, _P, _Z, _P]),
    "drba_get_drm_t_f32": (_I, [_P, _D, _D, _P, _Z, _P]),
    "drba_drm_workspace_bytes": (_Z, [_I, _I, _I]),
    "drba_drm_rife_f32": (_I, [_D, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "drba_drm_gmfss_f32": (_I, [_D, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "drba_backwarp_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "drba_resize_bilinear_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "drba_conv2d_direct_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P,
                                    _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "drba_conv_tc_f16": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "drba_ifnet_assemble": (_I, [_P, _P, _P, _P, _I, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    "drba_ifnet_upsample": (_I, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "drba_ifnet_blend": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "drba_ifnet_state_flow": (_I, [_P, _P, _I, _I, _P]),
}


class DrbaError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libdrba_b200.so (nvcc, -gencode arch=compute_100a,code=sm_100a)."""
    args = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        args.append("-B")
    res = subprocess.run(args, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise DrbaError("building libdrba_b200.so failed")
    return SO_PATH


_lib = None


def lib():
    """The loaded library; raises DrbaError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise DrbaError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            f"(or `make -C drba_b200/csrc`); there is no CPU fallback")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().drba_error_string(int(rc))
        raise DrbaError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")
