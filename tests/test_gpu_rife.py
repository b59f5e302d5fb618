"""GPU parity of the RIFE path (fp32 engine) against the fp32 oracle restatement and the
golden vectors generated from the reference classes (tests/golden/rife_golden.npz)."""
import numpy as np
import pytest
import torch

from drba_b200.weights import synth_ifnet_state, find_rife_weights, load_ifnet_state

pytestmark = pytest.mark.gpu

# fp32 engine vs fp32 reference: only the summation order of the convolutions differs
TOL = 5e-4


def _run(g, tag, state):
    from drba_b200.rife import RIFE
    I0, I1, I2 = (torch.from_numpy(g[k]).cuda() for k in ("I0", "I1", "I2"))
    m = RIFE(state=state, device="cuda", precision="fp32")
    f10, f01, f1, f0 = m.calc_flow(I1, I0)
    # flows carry a hole decision (mask < 0.999): compare outside disagreeing holes
    for got, key in ((f10, "flow10"), (f01, "flow01")):
        got, want = got.cpu().numpy(), g[f"{tag}_{key}"]
        big = 2.0 * max(want.shape[2], want.shape[3])
        flips = (got == big) != (want == big)
        assert flips.mean() < 1e-3
        np.testing.assert_allclose(got[~flips], want[~flips], rtol=0, atol=2e-3)
    assert tuple(f1.shape) == (1, 16) + tuple(I1.shape[2:])      # the reference's layout (models/rife.py:75)
    feat = f1.cpu().numpy()
    np.testing.assert_allclose(feat, g[f"{tag}_f1"], rtol=0, atol=1e-4)
    y = m.inference_ts(I0, I1, [0.4])[0]
    np.testing.assert_allclose(y.cpu().numpy(), g[f"{tag}_ts0.4"], rtol=0, atol=TOL)
    o1, reuse = m.inference_ts_drba(I0, I1, I2, np.array([0.6, 1.0, 1.4]), None, True)
    assert o1[1] is I1
    np.testing.assert_allclose(o1[0].cpu().numpy(), g[f"{tag}_w0_0.6"], rtol=0, atol=TOL)
    np.testing.assert_allclose(o1[2].cpu().numpy(), g[f"{tag}_w0_1.4"], rtol=0, atol=TOL)
    o2, _ = m.inference_ts_drba(I1, I2, I0, np.array([0.8, 1.2]), reuse, True)
    np.testing.assert_allclose(o2[0].cpu().numpy(), g[f"{tag}_w1_0.8"], rtol=0, atol=TOL)
    np.testing.assert_allclose(o2[1].cpu().numpy(), g[f"{tag}_w1_1.2"], rtol=0, atol=TOL)
    o3, _ = m.inference_ts_drba(I0, I1, I2, np.array([0.7]), None, False)
    np.testing.assert_allclose(o3[0].cpu().numpy(), g[f"{tag}_w0_0.7_nl"], rtol=0, atol=TOL)


def test_rife_fp32_synth_weights(golden_rife):
    _run(golden_rife, "synth", synth_ifnet_state(0))


def test_rife_fp32_real_weights(golden_rife):
    w = find_rife_weights()
    if w is None:
        pytest.skip("reference checkpoint not present on this machine")
    _run(golden_rife, "real", load_ifnet_state(w))


def test_rife_no_cpu_fallback():
    from drba_b200 import _lib
    from drba_b200.rife import RIFE
    with pytest.raises(_lib.DrbaError):
        RIFE(state=synth_ifnet_state(0), device="cpu")
