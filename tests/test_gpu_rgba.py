"""Frame consumer on the device: per-frame symmetric colour levels + RdBu colouring at the grid
nodes (bq_frame_rgba_dev) against oracle/colormap_np.py, byte for byte."""

import numpy as np
import pytest

import workloads
from oracle.colormap_np import frame_rgba

pytestmark = pytest.mark.gpu


def test_rgba_frames_equal_the_oracle_bytes(engine):
    import torch

    wl = workloads.config3(frames=5, steps=160)
    W = np.array(engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"]))
    W[3] = 0.0            # abs_max == 0 -> the reference falls back to 5 (animate.py:283-284)
    W[4] *= -3.0          # |amin| > amax
    W_t = torch.from_numpy(W).to(torch.device("cuda", engine.device))
    rgba = engine.frame_rgba_t(W_t).cpu().numpy()
    assert rgba.shape == (5, 160, 160, 4) and rgba.dtype == np.uint8
    for k in range(5):
        assert np.array_equal(rgba[k], frame_rgba(W[k])), k
    # other level counts / the plot() convention for an all-zero frame, a single 2-D frame, values on the levels
    one = engine.frame_rgba_t(W_t[0], levels=17, zero_abs_max=1.0).cpu().numpy()
    assert np.array_equal(one, frame_rgba(W[0], levels=17, zero_abs_max=1.0))
    lev = np.linspace(-2.0, 2.0, 100)
    edge = np.concatenate([lev, np.nextafter(lev, 5), np.nextafter(lev, -5), [1e-300, 0.0]]).reshape(2, -1)
    got = engine.frame_rgba_t(torch.from_numpy(edge).to(W_t.device)).cpu().numpy()
    assert np.array_equal(got, frame_rgba(edge))


def test_frames_to_rgba_dropin_helper(engine):
    from bosonic_qiskit_b200.animate import frames_to_rgba

    rng = np.random.default_rng(3)
    frames = [rng.normal(size=(40, 40)) * 0.1 for _ in range(3)]
    rgba = frames_to_rgba(frames)
    assert rgba.shape == (3, 40, 40, 4)
    for k in range(3):
        assert np.array_equal(rgba[k], frame_rgba(frames[k]))
    assert np.array_equal(frames_to_rgba(frames[1]), frame_rgba(frames[1]))
