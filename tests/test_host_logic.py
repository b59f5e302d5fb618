"""CPU tests of the host-side logic: C-ABI surface, driver loop / timestamp schedule, frame-window
sharding (incl. a world_size-2 gloo run), weight packing.  No CUDA compute is executed."""
import os
import re
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_header_symbol():
    from drba_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    header = open(os.path.join(ROOT, "include", "drba_b200.h")).read()
    declared = set(re.findall(r"DRBA_API\s+[\w\s\*]+?\b(drba_\w+)\s*\(", header))
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()                       # resolves every symbol; raises if one is not exported
    assert L.drba_version() >= 100
    assert b"workspace" in L.drba_error_string(-2)
    # argument errors are reported without touching a device
    assert L.drba_softsplat_f32(None, None, None, None, 1, 1, 4, 4, 9, 0, None, 0, None) == -1
    assert L.drba_softsplat_workspace_bytes(1, 64, 544, 960, 3) > 0


def test_no_cpu_fallback():
    from drba_b200 import _lib
    from drba_b200.rife import RIFE
    from drba_b200.softsplat import softsplat
    from drba_b200.weights import synth_ifnet_state
    with pytest.raises(_lib.DrbaError):
        RIFE(state=synth_ifnet_state(0), device="cpu")
    with pytest.raises(_lib.DrbaError):
        softsplat(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), None, "avg")


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "drba_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "drba_oracle" not in src, f


def test_net_input_size_and_timestamps():
    from drba_b200 import driver, tools
    assert tools.get_valid_net_inp_size(np.zeros((1080, 1920, 3)), 1.0, 64)["dst_size"] == (1088, 1920)
    assert tools.get_valid_net_inp_size(np.zeros((2160, 3840, 3)), 0.5, 128)["dst_size"] == (2304, 3840)
    calc_t = driver.make_calc_t(24.0, 60.0)
    np.testing.assert_allclose(calc_t(0), [0.6, 1.0, 1.4])
    np.testing.assert_allclose(calc_t(1), [0.8, 1.2])
    np.testing.assert_allclose(calc_t(2), [0.6, 1.0, 1.4])
    np.testing.assert_allclose(driver.make_calc_t(24.0, 0, times=2)(5), [0.75, 1.25])
    np.testing.assert_allclose(driver.make_calc_t(24.0, 0, times=3)(5), [2 / 3, 1.0, 4 / 3])


class FakeModel:
    """Deterministic stand-in with the RIFE interface whose outputs depend on `reuse` the way the
    real model's do (a cold start differs from a warm one)."""
    scale, pad_size = 1.0, 64

    def calc_flow(self, a, b, f0=None, f1=None):
        fa = a * 2 + 1 if f0 is None else f0
        fb = b * 2 + 1 if f1 is None else f1
        return a - b, b - a + 0.5, fa, fb

    def inference_ts(self, I0, I1, ts):
        return [I0 if t == 0 else (I1 if t == 1 else I0 * (1 - t) + I1 * t) for t in ts]

    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        flow10, flow01, f1, f0 = self.calc_flow(I1, I0) if not reuse else reuse
        flow12, flow21, f1, f2 = self.calc_flow(I1, I2) if reuse is None else self.calc_flow(I1, I2, f0=reuse[2])
        out = []
        for t in ts:
            if t == 1:
                out.append(I1)
            elif t < 1:
                out.append(I1 * t + I0 * (1 - t) + 0.01 * flow10 + 0.001 * f0)
            else:
                out.append(I1 * (2 - t) + I2 * (t - 1) + 0.01 * flow12 + 0.001 * f2)
        return out, (flow21, flow12, f2, f1)


def _clip(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand((1, 3, 4, 6), generator=g) for _ in range(n)]


def _scene_fn(cuts):
    def scene(a, b):
        return (float(a.sum()), float(b.sum())) in cuts
    return scene


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("with_cuts", [False, True])
def test_shards_reproduce_the_sequential_run(world, with_cuts):
    from drba_b200 import driver
    frames = _clip(23)
    scene = None
    if with_cuts:
        cuts = {(float(frames[i].sum()), float(frames[i + 1].sum())) for i in (0, 5, 6, 11, 21)}
        scene = _scene_fn(cuts)
    m = FakeModel()
    seq = list(driver.interpolate_sequence(m, frames, 24.0, 60.0, check_scene=scene))
    parts = []
    for a, b in driver.shard_ranges(driver.num_iterations(len(frames)), world):
        parts += list(driver.interpolate_shard(m, frames, 24.0, 60.0, check_scene=scene, a=a, b=b))
    assert len(parts) == len(seq)
    for x, y in zip(parts, seq):
        assert torch.equal(x, y)
    # 23 frames at 24 -> 60: 2.5 outputs per source frame interval
    if not with_cuts:
        assert len(seq) == int(round((len(frames) - 1) * 2.5)) + 2 or len(seq) > 50


@pytest.mark.parametrize("n_frames,world", [(4, 4), (3, 4), (2, 4), (2, 1), (5, 8)])
def test_more_ranks_than_iterations(n_frames, world):
    """world > n_iterations: surplus ranks emit nothing; head and tail appear exactly once."""
    from drba_b200 import driver
    frames = _clip(n_frames)
    m = FakeModel()
    seq = list(driver.interpolate_sequence(m, frames, 24.0, 60.0))
    parts = []
    for a, b in driver.shard_ranges(driver.num_iterations(n_frames), world):
        parts += list(driver.interpolate_shard(m, frames, 24.0, 60.0, a=a, b=b))
    assert len(parts) == len(seq)
    for x, y in zip(parts, seq):
        assert torch.equal(x, y)


class FakeGmfss:
    """Stand-in with the GMFSS interface (no calc_flow; `reuse` is a 6-list with swapped pairs)."""
    scale, pad_size = 1.0, 64

    def _reuse(self, a, b):
        return [a - b, b - a + 0.5, a * 2, b * 2, a * 3 + 1, b * 3 + 1]

    def shard_reuse(self, Ia, Ib):
        r = self._reuse(Ia, Ib)
        return [v for pair in zip(r[1::2], r[0::2]) for v in pair]

    def inference_ts(self, I0, I1, ts):
        return [I0 if t == 0 else (I1 if t == 1 else I0 * (1 - t) + I1 * t) for t in ts]

    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        r10 = self._reuse(I1, I0) if reuse is None else reuse
        r12 = self._reuse(I1, I2)
        out = []
        for t in ts:
            if t == 1:
                out.append(I1)
            elif t < 1:
                out.append(I1 * t + I0 * (1 - t) + 0.01 * r10[0] + 0.001 * r10[5])
            else:
                out.append(I1 * (2 - t) + I2 * (t - 1) + 0.01 * r12[0] + 0.001 * r12[5])
        return out, [v for pair in zip(r12[1::2], r12[0::2]) for v in pair]


@pytest.mark.parametrize("world", [2, 3])
def test_gmfss_style_model_shards(world):
    """Models without calc_flow (GMFSS / GMFSS_UNION) rebuild `reuse` through shard_reuse()."""
    from drba_b200 import driver
    frames = _clip(11)
    m = FakeGmfss()
    seq = list(driver.interpolate_sequence(m, frames, 24.0, 60.0))
    parts = []
    for a, b in driver.shard_ranges(driver.num_iterations(len(frames)), world):
        parts += list(driver.interpolate_shard(m, frames, 24.0, 60.0, a=a, b=b))
    assert len(parts) == len(seq)
    for x, y in zip(parts, seq):
        assert torch.equal(x, y)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from drba_b200 import driver
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frames = _clip(n_frames)
    a, b = driver.shard_ranges(driver.num_iterations(n_frames), world)[rank]
    mine = torch.stack(list(driver.interpolate_shard(FakeModel(), frames, 24.0, 60.0, a=a, b=b)))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)     # outputs are concatenated in shard order by the writer
    dist.barrier()
    if rank == 0:
        q.put(torch.cat(gathered))
    dist.destroy_process_group()


def test_two_process_gloo_sharding_matches_sequential():
    import torch.multiprocessing as mp
    from drba_b200 import driver
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n = _free_port(), 17
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seq = torch.stack(list(driver.interpolate_sequence(FakeModel(), _clip(n), 24.0, 60.0)))
    assert torch.equal(got, seq)


def test_tc_weight_packing_matches_conv():
    """The packed/permuted fp16 weights of the tensor-core engine reproduce F.conv2d on the packed
    channel order written by drba_ifnet_assemble (emulated on the CPU)."""
    import torch.nn.functional as F
    from drba_b200.ifnet import _packed_input_channels, _tc_conv3x3, _tc_lastconv
    g = torch.Generator().manual_seed(1)
    for first, cin in ((True, 39), (False, 52)):
        w = torch.randn((8, cin, 3, 3), generator=g)
        b = torch.randn((8,), generator=g)
        x = torch.randn((1, cin, 6, 10), generator=g)
        m = _packed_input_channels(first)
        layer = _tc_conv3x3(w, b, 1, 1, "cpu", in_map=m)
        xp = torch.zeros((1, len(m), 6, 10))
        for pc, rc in enumerate(m):
            if rc >= 0:
                xp[:, pc] = x[:, rc]
        wp = layer.w.float()[0]                                   # [9][cout_pad][cin_pad]
        ref = F.conv2d(x, w.half().float(), b, 1, 1)
        xpad = F.pad(xp, (1, 1, 1, 1))
        got = torch.zeros((1, layer.cout_pad, 6, 10))
        for t in range(9):
            ky, kx = t // 3, t % 3
            got += torch.einsum("oc,bchw->bohw", wp[t], xpad[:, :, ky:ky + 6, kx:kx + 10])
        got = got + layer.b[0].view(1, -1, 1, 1)
        torch.testing.assert_close(got[:, :8], ref, rtol=1e-3, atol=1e-3)
    wt = torch.randn((16, 52, 4, 4), generator=g)
    bt = torch.randn((52,), generator=g)
    x = torch.randn((1, 16, 5, 7), generator=g)
    layer = _tc_lastconv(wt, bt, "cpu")
    ref = F.conv_transpose2d(x, wt.half().float(), bt, 2, 1)
    xpad = F.pad(x, (1, 1, 1, 1))
    for gph in range(4):
        py, px = gph >> 1, gph & 1
        acc = torch.zeros((1, 64, 5, 7))
        for t in range(4):
            dy, dx = layer.dy[gph * 4 + t], layer.dx[gph * 4 + t]
            acc += torch.einsum("oc,bchw->bohw", layer.w.float()[gph, t], xpad[:, :, 1 + dy:6 + dy, 1 + dx:8 + dx])
        acc = acc + layer.b[gph].view(1, -1, 1, 1)
        torch.testing.assert_close(acc[:, :52], ref[:, :, py::2, px::2], rtol=1e-3, atol=1e-3)


def test_gmfss_host_contract_without_gpu():
    """GMFSS wrapper: CPU device is refused (no fallback); synthetic weights cover every tensor the nets read."""
    from drba_b200 import _lib
    from drba_b200.gmfss import GMFSS
    from drba_b200.weights import gmfss_param_shapes, synth_gmfss_state
    st = synth_gmfss_state(3)
    shapes = gmfss_param_shapes()
    assert {k: len(v) for k, v in shapes.items()} == {"feat": 18, "metric": 14, "fusionnet": 133}
    for net, lst in shapes.items():
        for name, shape in lst:
            assert tuple(st[net][name].shape) == tuple(shape)
    with pytest.raises(_lib.DrbaError):
        GMFSS(state=st, device="cpu")


def test_window_graph_reuse_tree_helpers():
    """drba_b200/_graphs.py: the nested `reuse` of GMFSS (flows, metrics, tuples of feature maps: models/gmfss.py:71)
    is flattened depth first and re-created with the same nesting for the static graph inputs."""
    import torch
    from drba_b200._graphs import like_tree, tensors_of
    a, b, c, d = (torch.arange(6, dtype=torch.float32).reshape(2, 3) + k for k in range(4))
    tree = [a, b, (c, d.t())]           # d.t(): non-contiguous on purpose
    flat = tensors_of(tree)
    assert [t.data_ptr() for t in flat] == [a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr()]
    twin = like_tree(tree)
    assert isinstance(twin, list) and isinstance(twin[2], tuple) and len(twin[2]) == 2
    for x, y in zip(tensors_of(twin), flat):
        assert x.shape == y.shape and x.dtype == y.dtype and x.is_contiguous() and x.data_ptr() != y.data_ptr()
        x.copy_(y)
        assert torch.equal(x, y)


def test_ctypes_signatures_match_header_prototypes():
    """Every ctypes signature in drba_b200/_lib.py has as many arguments as the C prototype in
    include/drba_b200.h, pointers where the header has pointers and scalars where it has scalars."""
    import ctypes
    import re
    from drba_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "drba_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = dict(re.findall(r"\b(drba_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        _, argtypes = _lib.SIGNATURES[name]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for decl, ct in zip(args, argtypes):
            is_ptr_c = "*" in decl
            is_ptr_py = ct in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(ct, type(ctypes.POINTER(ctypes.c_int))) \
                or (isinstance(ct, type) and issubclass(ct, ctypes.Array))
            assert is_ptr_c == is_ptr_py, (name, decl, ct)
