"""stereo_widener / stereo_panner / stereo_bus on the GPU through the C ABI vs the oracle and the reference golden
(tolerance 1e-4 relative fp32; observed ~1e-7)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from helpers import SR, peak_err, run_with_grads

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cmp(cuda_device, fg, fo, x, p):
    y, dx, dp = run_with_grads(lambda xx, q: fg(xx, SR, q[0]), x, [p], torch.float32, cuda_device)
    y64, dx64, dp64 = run_with_grads(lambda xx, q: fo(xx, SR, q[0]), x, [p], torch.float64, "cpu")
    assert y.shape == y64.shape
    assert peak_err(y, y64).max() < TOL
    assert (dx.double() - dx64).abs().max() <= TOL * dx64.abs().max() + 1e-12
    assert (dp[0].double() - dp64[0]).abs().max() <= 1e-4 * dp64[0].abs().max() + 1e-12


def test_stereo_golden(cuda_device):
    import dasp_pytorch_b200 as D
    g = load_golden("stereo.npz")
    for tag, fn, pkey, pname in (("wid", D.stereo_widener, "wid_w", "width"), ("pan", D.stereo_panner, "pan_p", "pan"),
                                 ("bus", D.stereo_bus, "bus_s", "send_db")):
        y, dx, dp = run_with_grads(lambda xx, q: fn(xx, SR, q[0]), g[f"{tag}_x"], [g[pkey]], torch.float32, cuda_device)
        assert peak_err(y, g[f"{tag}_y64"]).max() < TOL, tag
        assert peak_err(dx, g[f"{tag}_dx64"]).max() < TOL, tag
        ref = g[f"{tag}_d_{pname}"]
        assert np.abs(dp[0].numpy().reshape(ref.shape) - ref).max() <= 1e-4 * np.abs(ref).max(), tag


@pytest.mark.parametrize("n", [48000, 4097, 3, 2048 * 3 + 4])
def test_stereo_shapes_vs_oracle(cuda_device, n):
    import dasp_pytorch_b200 as D
    g = torch.Generator().manual_seed(n)
    _cmp(cuda_device, D.stereo_widener, oracle.stereo_widener, torch.rand(3, 2, n, generator=g) * 2 - 1, torch.rand(3, generator=g))
    _cmp(cuda_device, D.stereo_panner, oracle.stereo_panner, torch.rand(2, 3, n, generator=g) * 2 - 1,
         torch.rand(2, 3, generator=g) * 0.9 + 0.05)
    _cmp(cuda_device, D.stereo_bus, oracle.stereo_bus, torch.rand(2, 2, 4, n, generator=g) * 2 - 1,
         torch.rand(2, 4, 1, generator=g) * 30 - 24)


def test_stereo_properties_full_size(cuda_device):
    import dasp_pytorch_b200 as D
    torch.manual_seed(0)
    x = torch.rand(256, 2, 48000, device=cuda_device) * 2 - 1
    half = torch.full((256,), 0.5, device=cuda_device)
    assert torch.equal(D.stereo_widener(x, SR, half), x)                        # width 0.5 is the identity (c = 0)
    mono = D.stereo_widener(x, SR, half * 0)                                    # width 0: both outputs = L + R
    assert torch.equal(mono[:, 0], mono[:, 1]) and torch.allclose(mono[:, 0], x[:, 0] + x[:, 1])
    xm = torch.rand(64, 8, 48000, device=cuda_device)
    y = D.stereo_panner(xm, SR, torch.full((64, 8), 0.5, device=cuda_device))
    assert y.shape == (64, 2, 8, 48000) and torch.allclose(y[:, 0], y[:, 1], rtol=1e-6)     # centre pan: equal gains
    xb = torch.rand(32, 2, 8, 48000, device=cuda_device)
    s = D.stereo_bus(xb, SR, torch.zeros(32, 8, 1, device=cuda_device))
    assert torch.allclose(s, xb.sum(2), rtol=1e-5, atol=1e-5)                   # 0 dB sends: plain sum
