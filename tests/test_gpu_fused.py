"""Block input fused with conv0a (csrc/ifnet_fused.cu, drba_ifnet_block_conv0a_f16) against the two kernels it
replaces: drba_ifnet_assemble (NHWC fp16 block input, models/rife_426_heavy/IFNet_HDv3.py:151-155 + :85-88) followed by
the tcgen05 conv engine's conv0a (IFNet_HDv3.py:66-69).  Both paths round the assembled input to fp16 and accumulate
the same 36 MMAs per tile in fp32, so the results are expected to agree to the last bit; the stated tolerance is one
fp16 ulp of the output (FMA contraction may differ between the two assembly kernels)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _job(g, H, W, s_prev, timestep_map, amp):
    dev = "cuda"
    r = lambda *shape: torch.randn(shape, generator=g)
    return {"img0": torch.rand((1, 3, H, W), generator=g).to(dev), "img1": torch.rand((1, 3, H, W), generator=g).to(dev),
            "f0": r(H, W, 16).half().to(dev), "f1": r(H, W, 16).half().to(dev),
            "ts_t": torch.rand((1, 1, H, W), generator=g).to(dev) if timestep_map else None, "ts_s": 0.37,
            "flow": (amp * r(H, W, 4)).to(dev),
            "prev": (r(H // s_prev, W // s_prev, 16).to(dev), 1, s_prev)}


@pytest.mark.parametrize("bi,s,H,W,nj,ts_map,amp", [
    (4, 1, 64, 128, 1, True, 3.0),        # block 4: scale 1, 16 output channels, whole tiles
    (4, 1, 76, 88, 2, False, 6.0),        # ragged tiles in both directions, two jobs, scalar timestep
    (4, 1, 132, 260, 2, True, 40.0),      # flows that leave the frame (border clamp), more tiles than one wave row
    (3, 2, 128, 192, 1, True, 3.0),       # block 3: scale 2 (2 x 2 means), 32 output channels
    (3, 2, 152, 176, 2, False, 10.0),
])
def test_block_conv0a_matches_assemble_plus_conv(bi, s, H, W, nj, ts_map, amp):
    from drba_b200.ifnet import IFNetEngine, _BLOCKS
    from drba_b200.weights import synth_ifnet_state
    eng = IFNetEngine(synth_ifnet_state(3), "cuda", "fp16")
    name, _, c = _BLOCKS[bi]
    g = torch.Generator(device="cpu").manual_seed(100 * bi + H + nj)
    jobs = [_job(g, H, W, 2 * s, ts_map, amp) for _ in range(nj)]
    h, w = H // s, W // s
    layer = eng.tc[f"{name}.conv0a"]
    want = []
    for j in jobs:
        x = torch.empty((h, w, 64), dtype=torch.float16, device="cuda")
        eng._assemble(x, 1, 64, j["img0"], j["img1"], j["f0"], j["f1"], j["ts_t"], j["ts_s"], j["flow"], j["prev"], H, W, s)
        y = torch.full((h // 2, w // 2, c // 2), float("nan"), dtype=torch.float16, device="cuda")
        eng._conv_tc(layer, x, h, w, y, h // 2, w // 2, c // 2)
        want.append(y)
    got = [torch.full((h // 2, w // 2, c // 2), float("nan"), dtype=torch.float16, device="cuda") for _ in jobs]
    eng._block_conv0a(name, jobs, got, H, W, s)
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.isfinite(a.float()).all()
        diff = (a.float() - b.float()).abs()
        tol = 2.0 ** -10 * b.float().abs() + 1e-4
        assert (diff <= tol).all(), f"max diff {diff.max().item()} at {int((diff > tol).sum())} of {diff.numel()} elements"


def test_block_conv0a_rejects_bad_arguments():
    from drba_b200 import _lib
    L = _lib.lib()
    arr = (_lib.BlockInput * 1)()
    assert L.drba_ifnet_block_conv0a_f16(ctypes.addressof(arr), 1, None, None, 16, 64, 64, 1, None) != 0      # no weights
    assert L.drba_ifnet_block_conv0a_f16(ctypes.addressof(arr), 3, None, None, 16, 64, 64, 1, None) != 0      # too many jobs
