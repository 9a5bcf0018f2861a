"""Shared helpers for the GPU parity tests: parameter ranges of the reference Processor
classes (modules.py:136-155, 179-186, 204-230), seeded inputs, oracle runs with gradients."""
import torch

SR = 44100


def eq_ranges(sr=SR):
    g, q = (-20.0, 20.0), (0.1, 6.0)
    hi = (sr // 2) - 1000
    fr = [(20, 2000), (80, 2000), (2000, 8000), (8000, 12000), (12000, hi), (4000, hi)]
    out = []
    for f in fr:
        out += [g, f, q]
    return out


COMP_RANGES = [(-60.0, 0.0), (1.0, 20.0), (5.0, 100.0), (5.0, 100.0), (0.0, 12.0), (0.0, 12.0)]
REVERB_RANGES = [(0.0, 1.0)] * 25


def denorm(p01, ranges, dtype=torch.float32):
    """(bs, P) in [0,1] -> list of P tensors (bs,), denormalised in fp32 like modules.py:13-14."""
    p01 = torch.as_tensor(p01).float()
    return [(p01[:, i] * (hi - lo) + lo).to(dtype) for i, (lo, hi) in enumerate(ranges)]


def run_with_grads(fn, x, params, dtype, device):
    """y, dx, [dparam] for loss = mean(y^2); tensors created on `device` in `dtype`."""
    xx = torch.as_tensor(x).to(device=device, dtype=dtype).clone().requires_grad_(True)
    pp = [torch.as_tensor(p).to(device=device, dtype=dtype).clone().requires_grad_(True) for p in params]
    y = fn(xx, pp)
    y.pow(2).mean().backward()
    return y.detach().cpu(), xx.grad.detach().cpu(), [None if p.grad is None else p.grad.detach().cpu() for p in pp]


def peak_err(a, b):
    """per-item max|a-b| / max|b|"""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    bs = b.shape[0]
    return (a - b).reshape(bs, -1).abs().amax(1) / b.reshape(bs, -1).abs().amax(1).clamp_min(1e-30)


def param_grad_err(got, ref):
    """per-item |got-ref| / max_over_params|ref| for lists of (bs,) gradients."""
    keep = [i for i, r in enumerate(ref) if r is not None]
    g = torch.stack([torch.as_tensor(got[i]).double().reshape(-1) for i in keep], 1)
    r = torch.stack([torch.as_tensor(ref[i]).double().reshape(-1) for i in keep], 1)
    return ((g - r).abs() / r.abs().amax(1, keepdim=True).clamp_min(1e-30)).amax(1)
