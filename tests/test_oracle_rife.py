"""Pins the fp32 RIFE restatement (oracle/ifnet.py) to the reference classes' outputs."""
import numpy as np
import pytest
import torch

from drba_b200.weights import synth_ifnet_state, find_rife_weights, load_ifnet_state
from oracle.ifnet import RIFEOracle, ifnet_forward

TOL = 2e-4  # fp32 conv summation order (MKL-DNN vs itself is identical; splat/backwarp restated in C)


def _run(g, tag, state):
    torch.set_grad_enabled(False)
    I0, I1, I2 = (torch.from_numpy(g[k]) for k in ("I0", "I1", "I2"))
    m = RIFEOracle(state)
    f10, f01, f1, f0 = m.calc_flow(I1, I0)
    np.testing.assert_allclose(f10.numpy(), g[f"{tag}_flow10"], rtol=0, atol=TOL)
    np.testing.assert_allclose(f01.numpy(), g[f"{tag}_flow01"], rtol=0, atol=TOL)
    np.testing.assert_allclose(f1.numpy(), g[f"{tag}_f1"], rtol=0, atol=1e-5)
    y = ifnet_forward(m.sd, torch.cat((I0, I1), 1), 0.4, m.scale_list)[0]
    np.testing.assert_allclose(y.numpy(), g[f"{tag}_ts0.4"], rtol=0, atol=TOL)
    o1, reuse = m.inference_ts_drba(I0, I1, I2, np.array([0.6, 1.0, 1.4]), None, True)
    assert o1[1] is I1
    np.testing.assert_allclose(o1[0].numpy(), g[f"{tag}_w0_0.6"], rtol=0, atol=TOL)
    np.testing.assert_allclose(o1[2].numpy(), g[f"{tag}_w0_1.4"], rtol=0, atol=TOL)
    o2, _ = m.inference_ts_drba(I1, I2, I0, np.array([0.8, 1.2]), reuse, True)
    np.testing.assert_allclose(o2[0].numpy(), g[f"{tag}_w1_0.8"], rtol=0, atol=TOL)
    np.testing.assert_allclose(o2[1].numpy(), g[f"{tag}_w1_1.2"], rtol=0, atol=TOL)
    o3, _ = m.inference_ts_drba(I0, I1, I2, np.array([0.7]), None, False)
    np.testing.assert_allclose(o3[0].numpy(), g[f"{tag}_w0_0.7_nl"], rtol=0, atol=TOL)


def test_rife_oracle_synth_weights(golden_rife):
    _run(golden_rife, "synth", synth_ifnet_state(0))


def test_rife_oracle_real_weights(golden_rife):
    w = find_rife_weights()
    if w is None:
        pytest.skip("reference checkpoint not present on this machine")
    _run(golden_rife, "real", load_ifnet_state(w))
