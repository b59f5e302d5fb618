"""ctypes loader for libdrba_b200.so (the C ABI declared in include/drba_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the
product path raises.  ``build()`` compiles it in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRBA_B200_LIB: load another build of the same C ABI (A/B timing of kernel variants on one box)
SO_PATH = os.environ.get("DRBA_B200_LIB") or os.path.join(_HERE, "libdrba_b200.so")
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
    "drba_frame_ingest_u8": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "drba_frame_egress_u8": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "drba_splat_lists_workspace_bytes": (_Z, [_I, _I]),
    "drba_splat_lists_build": (_I, [_P, _P, _I, _I, _I, _P, _Z, _P]),
    "drba_splat_lists_apply_nhwc_f16": (_I, [_P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "drba_splat_lists_apply_nchw_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "drba_splat_lists_release": (_I, [_P, _I, _I, _P]),
    "drba_pack_planes_nhwc_f16": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _I, _P]),
    "drba_unpack_nhwc_f16": (_I, [_P, _I, _P, _I, _I, _I, _I, _F, _F, _P]),
    "drba_gmfss_metric_prep": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "drba_gmfss_scale_flow": (_I, [_P, _P, _P, _F, _I, _I, _I, _P, _P, _P]),
    "drba_gmflow_normalize_img": (_I, [_P, _P, _I, _I, _P]),
    "drba_gmflow_inorm_stats": (_I, [_P, _I, _I, _I, _P, _P]),
    "drba_gmflow_inorm_apply": (_I, [_P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _P]),
    "drba_gmflow_add_position": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "drba_gmflow_window_pack": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "drba_gmflow_softmax_rows": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "drba_gmflow_ln_residual": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "drba_gmflow_soft_readout": (_I, [_P, _I, _I, _I, _P, _I, _I, _F, _P, _P]),
    "drba_gmflow_local_match": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "drba_gmflow_local_propagate": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "drba_gmflow_warp_feature": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "drba_gmflow_upsampler_input": (_I, [_P, _P, _P, _I, _I, _P]),
    "drba_gmflow_convex_upsample": (_I, [_P, _P, _P, _I, _I, _P]),
    "drba_axpby_f32": (_I, [_P, _F, _P, _F, _P, _Z, _P]),
    "drba_gmfss_union_fix_timesteps": (_I, [_P, _P, _P, _P, _Z, _P]),
    "drba_gmfss_union_swap_nhwc_f16": (_I, [_P, _I, _P, _P, _I, _I, _P]),
    "drba_gmfss_union_swap_nchw_f32": (_I, [_P, _P, _I, _P, _P, _I, _I, _P]),
    "drba_conv2d_direct_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P,
                                    _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "drba_conv_tc_f16": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "drba_conv_tc_program_f16": (_I, [_P, _I, _I, _P, _P]),
    "drba_conv_tc_debug_trace": (_I, [_P]),
    "drba_conv_tc_status": (_I, []),
    "drba_check_scene_f32": (_I, [_P, _P, _c.c_longlong, _c.c_longlong, _I, _I, _I, _F, _P, _P, _P]),
    "drba_ifnet_assemble": (_I, [_P, _P, _P, _P, _I, _P, _F, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "drba_ifnet_flow_accum": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _P]),
    "drba_ifnet_assemble_terms": (_I, [_P, _P, _P, _P, _P, _F, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P]),
    "drba_ifnet_flow_sum": (_I, [_P, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P]),
    "drba_ifnet_blend": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _I, _P]),
    "drba_ifnet_block_conv0a_f16": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P]),
}


class ConvLayer(ctypes.Structure):
    """struct drba_conv_layer (include/drba_b200.h)."""
    _fields_ = [("in_", _P * 2), ("res", _P * 2), ("out", _P * 2), ("w", _P), ("bias", _P), ("slope", _P),
                ("H", _I), ("W", _I), ("Cin", _I), ("G", _I), ("T", _I), ("dy", _I * 36), ("dx", _I * 36),
                ("cout_pad", _I), ("cout", _I), ("S", _I), ("OH", _I), ("OW", _I), ("epilogue", _I), ("act", _I),
                ("out_cstride", _I), ("out_os", _I),
                ("res2", _P * 2), ("out1", _P * 2), ("out2", _P * 2), ("act1", _I), ("act2", _I),
                ("slope0", _F), ("slope1", _F), ("slope2", _F), ("bgemm", _I)]


CONV_MAX_LAYERS = 12


class BlockInput(ctypes.Structure):
    """struct drba_ifnet_block_input (include/drba_b200.h)."""
    _fields_ = [("img0", _P), ("img1", _P), ("f0", _P), ("f1", _P), ("timestep", _P), ("timestep_scalar", _F),
                ("flow", _P), ("tmp_prev", _P), ("s_prev", _I), ("out", _P)]


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

# --- launch accounting ---------------------------------------------------------------
# Every host wrapper reports the kernels it launched (count(), exact numbers) so bench.py can
# state `gpu_launches`; with a LaunchProfiler installed each call is also bracketed by CUDA
# events on the launching stream (per-kernel-family time shares and the roofline numbers).
KERNEL_LAUNCHES = 0
PROFILER = None


def count(n):
    global KERNEL_LAUNCHES
    KERNEL_LAUNCHES += n


class LaunchProfiler:
    """Collects (family, ms, flops, bytes) per wrapped call; install with `with LaunchProfiler() as p:`."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global PROFILER
        PROFILER = self
        return self

    def __exit__(self, *exc):
        global PROFILER
        PROFILER = None

    def summary(self):
        import torch
        torch.cuda.synchronize()
        fam = {}
        for name, a, b, flops, nbytes, nk in self.records:
            d = fam.setdefault(name, {"calls": 0, "kernels": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["kernels"] += nk
            d["ms"] += a.elapsed_time(b)
            d["flops"] += flops
            d["bytes"] += nbytes
        return fam


class launch:
    """Context manager around one C-ABI call: counts its kernels, times it when profiling."""

    def __init__(self, name, kernels=1, flops=0.0, nbytes=0.0):
        self.name, self.kernels, self.flops, self.nbytes = name, kernels, flops, nbytes

    def __enter__(self):
        count(self.kernels)
        if PROFILER is not None:
            import torch
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILER is not None:
            self.b.record()
            PROFILER.records.append((self.name, self.a, self.b, self.flops, self.nbytes, self.kernels))


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


def check_async(what):
    """Sticky device-side errors of work already enqueued (the grid-barrier time-out of a persistent conv program,
    include/drba_b200.h: drba_conv_tc_status); a plain host read, no synchronisation."""
    rc = lib().drba_conv_tc_status()
    if rc != 0:
        check(rc, what)


def check(rc, what):
    if rc != 0:
        msg = lib().drba_error_string(int(rc))
        raise DrbaError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")
