"""Scene detection: the CPU oracle (oracle/scene.py) is pinned to the SSIM values the reference
computed (tests/golden/scene_golden.npz); the GPU kernel (csrc/scene.cu through drba_b200.tools) is compared with
both, including decisions at threshold edges.  models/utils/tools.py:27-30, models/pytorch_msssim/__init__.py:83-136."""
import importlib.util
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_golden_scene", os.path.join(ROOT, "tests", "golden", "make_golden_scene.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


def _cases():
    d = np.load(os.path.join(ROOT, "tests", "golden", "scene_golden.npz"))
    for name, _ in _gen.SIZES:
        base, other, noise = (torch.from_numpy(d[f"{name}_{k}"]) for k in ("base", "other", "noise"))
        for tag, x1, x2 in _gen.scene_pairs(base, other, noise):
            yield f"{name}_{tag}", x1, x2, float(d[f"{name}_{tag}_ssim"])


def test_oracle_ssim_matches_reference():
    from oracle import scene
    n = 0
    for key, x1, x2, want in _cases():
        flag, s = scene.check_scene(x1, x2, 0.3)
        assert abs(s - want) <= 2e-6, (key, s, want)
        assert flag == (want < 0.3)
        n += 1
    assert n == 40


@pytest.mark.gpu
def test_kernel_ssim_and_flags_vs_reference_and_oracle():
    """One-kernel check_scene against the reference's SSIM (fixture) -- tolerance 2e-5 absolute on a [-1, 1] score
    (separable fp32 passes vs the reference's 11^3 window; 255-range inputs reach 1e-5) -- and its decisions at the
    CLI's threshold and at thresholds placed just either side of every case's own score."""
    from drba_b200 import tools
    for key, x1, x2, want in _cases():
        a, b = x1.cuda(), x2.cuda()
        ssim, flag = tools.check_scene_async(a, b, 0.3)
        s = float(ssim.item())
        assert abs(s - want) <= 2e-5, (key, s, want)
        if abs(want - 0.3) > 1e-4:
            assert bool(flag.item()) == (want < 0.3), key
            assert tools.check_scene(a, b, 0.3) == (want < 0.3)
        # threshold edges: a threshold 1e-4 above the score must flag a cut, 1e-4 below must not
        assert tools.check_scene(a, b, want + 1e-4) is True, key
        assert tools.check_scene(a, b, want - 1e-4) is False, key


@pytest.mark.gpu
def test_scene_detector_tickets_do_not_block_and_match():
    """SceneDetector: flags land in pinned host memory; tickets can be redeemed later, in any order."""
    from drba_b200 import tools
    det = tools.SceneDetector("cuda", threshold=0.3, depth=8)
    cases = list(_cases())[:12]
    tickets = []
    for i, (key, x1, x2, want) in enumerate(cases):
        tickets.append((det.submit(x1.cuda(), x2.cuda()), want, key))
        if len(tickets) == 6:          # redeem in reverse before the ring wraps
            for t, w, k in reversed(tickets):
                if abs(w - 0.3) > 1e-4:
                    assert det.result(t) == (w < 0.3), k
                assert abs(det.ssim(t) - w) <= 2e-5
            tickets = []


@pytest.mark.gpu
def test_ssim_matlab_wrapper_and_errors():
    from drba_b200 import _lib, tools
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand((1, 3, 32, 32), generator=g), torch.rand((1, 3, 32, 32), generator=g)
    from oracle import scene
    want = float(scene.ssim_matlab(a, b))
    got = float(tools.ssim_matlab(a.cuda(), b.cuda()).item())
    assert abs(got - want) <= 2e-5
    with pytest.raises(_lib.DrbaError):
        tools.check_scene(a, b)                 # CPU tensors: no fallback
    with pytest.raises(ValueError):
        tools.check_scene(a.cuda(), torch.rand((1, 3, 16, 32)).cuda())
