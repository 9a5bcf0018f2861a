"""Multi-GPU property: items are independent, so a batch sharded over two devices
gives BIT-IDENTICAL per-item results (outputs and gradients) to the unsharded batch on one device.  Skipped on
single-GPU boxes; run with `gpurun --gpus 2`."""
import pytest
import torch

import oracle
from dasp_pytorch_b200.dist import shard_bounds
from helpers import COMP_RANGES, SR, denorm, eq_ranges

pytestmark = pytest.mark.gpu


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")


def _run(dev, x, eq, comp, rev, drive, noise, L, taps):
    import dasp_pytorch_b200 as D
    c = lambda t: t.to(dev)
    xx = c(x).requires_grad_(True)
    pe = [c(q).requires_grad_(True) for q in eq]
    y = D.parametric_eq(xx, SR, *pe)
    y = D.compressor(y, SR, *[c(q) for q in comp])
    y = D.noise_shaped_reverberation(y, SR, *[c(q) for q in rev], num_samples=L, num_bandpass_taps=taps, noise=c(noise))
    y = D.distortion(y, SR, c(drive))
    (y * y).sum().backward()
    return y.detach().cpu(), xx.grad.cpu(), torch.stack([q.grad.cpu() for q in pe], 1)


def test_two_device_shards_are_bit_identical():
    _need2()
    torch.manual_seed(0)
    bs, n, L, taps = 6, 9000, 5000, 255
    x = torch.rand(bs, 2, n) * 2 - 1
    eq = denorm(torch.rand(bs, 18), eq_ranges())
    comp = denorm(torch.rand(bs, 6).clamp(min=0.05), COMP_RANGES)
    rev = [torch.rand(bs) for _ in range(25)]
    drive = torch.rand(bs * 2) * 12
    noise = oracle.reverb_noise(bs, L, taps, 3)
    full = _run("cuda:0", x, eq, comp, rev, drive, noise, L, taps)
    parts = []
    for r in range(2):
        lo, hi = shard_bounds(bs, 2, r)
        parts.append(_run(f"cuda:{r}", x[lo:hi], [q[lo:hi] for q in eq], [q[lo:hi] for q in comp],
                          [q[lo:hi] for q in rev], drive[2 * lo:2 * hi], noise[2 * lo:2 * hi], L, taps))
    for i in range(3):
        assert torch.equal(torch.cat([p[i] for p in parts]), full[i]), i
