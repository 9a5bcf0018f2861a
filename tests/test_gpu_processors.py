"""Processor classes on the GPU: the packed fast path of process_normalized equals the explicit functional
calls, keeps the reference's error behaviour, and the whole chain can be captured in a CUDA graph."""
import pytest
import torch

from helpers import COMP_RANGES, SR, denorm, eq_ranges

pytestmark = pytest.mark.gpu


def test_process_normalized_fast_path_matches_functional(cuda_device):
    import dasp_pytorch_b200 as D
    torch.manual_seed(0)
    bs, n = 5, 6000
    x = torch.rand(bs, 2, n, device=cuda_device) * 2 - 1
    # EQ
    p = torch.rand(bs, 18, device=cuda_device, requires_grad=True)
    y = D.ParametricEQ(SR).process_normalized(x, p)
    cols = [q.to(cuda_device) for q in denorm(p.detach().cpu(), eq_ranges())]
    y_ref = D.parametric_eq(x, SR, *cols)
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-5)
    # ... and against the oracle fed the reference's own denormalisation (modules.py:13-14) of the same tensor
    import oracle
    from helpers import peak_err
    y_orc = oracle.parametric_eq(x.detach().cpu().double(), SR, *[c.cpu().double() for c in cols], fsm_tail=1 << 16)
    assert peak_err(y.detach().cpu(), y_orc).max() < 1e-4
    y.pow(2).mean().backward()
    assert p.grad is not None and p.grad.shape == (bs, 18) and bool(p.grad.abs().sum() > 0)
    # compressor (release column receives zero gradient: unused upstream)
    pc = torch.rand(bs, 6, device=cuda_device).clamp(min=0.05).requires_grad_(True)
    yc = D.Compressor(SR).process_normalized(x, pc)
    cols = [q.to(cuda_device) for q in denorm(pc.detach().cpu(), COMP_RANGES)]
    assert torch.allclose(yc, D.compressor(x, SR, *cols), rtol=1e-4, atol=1e-6)
    yc.pow(2).mean().backward()
    assert float(pc.grad[:, 3].abs().max()) == 0.0 and bool(pc.grad[:, 0].abs().sum() > 0)
    # reverb: same seed -> same device noise on both paths
    pr = torch.rand(bs, 25, device=cuda_device)
    torch.manual_seed(5)
    yr = D.NoiseShapedReverb(SR).process_normalized(x, pr)
    torch.manual_seed(5)
    yr_ref = D.noise_shaped_reverberation(x, SR, *[pr[:, i] for i in range(25)])
    assert torch.equal(yr, yr_ref)
    # gain / distortion go through the generic (name-based) path
    pg = torch.rand(bs, 1, device=cuda_device)
    assert torch.allclose(D.Gain(SR).process_normalized(x, pg), D.gain(x, SR, pg[:, 0] * 48 - 24), rtol=1e-6, atol=0)
    xm = x[:, :1].contiguous()
    assert torch.allclose(D.Distortion(sample_rate=SR).process_normalized(xm, pg), D.distortion(xm, SR, pg[:, 0] * 24), atol=1e-6)


def test_process_normalized_errors_and_repointing(cuda_device):
    import dasp_pytorch_b200 as D
    x = torch.rand(2, 2, 512, device=cuda_device)
    eq = D.ParametricEQ(SR)
    with pytest.raises(ValueError):
        eq.process_normalized(x, torch.rand(2, 17, device=cuda_device))
    bad = torch.rand(2, 18, device=cuda_device)
    bad[1, 4] = 1.2
    with pytest.raises(ValueError, match="band0_cutoff_freq"):
        eq.process_normalized(x, bad)
    # re-pointing process_fn falls back to the reference's name-based dispatch
    seen = {}
    eq.process_fn = lambda xx, sr, **kw: seen.update(kw) or xx
    eq.process_normalized(x, torch.rand(2, 18, device=cuda_device))
    assert list(seen) == list(eq.param_ranges)


def test_chain_cuda_graph_capture(cuda_device):
    """fwd+bwd of eq -> compressor -> reverb -> distortion captured once and replayed: no host work, no
    allocation, no synchronisation happens inside the C-ABI calls (the reverb's cuFFT plans and filter-bank
    spectra are created by the warm-up)."""
    import dasp_pytorch_b200 as D
    torch.manual_seed(1)
    bs, n, L, taps = 4, 8192, 6000, 255
    x = (torch.rand(bs, 2, n, device=cuda_device) * 2 - 1).requires_grad_(True)
    eq = [q.to(cuda_device) for q in denorm(torch.rand(bs, 18), eq_ranges())]
    comp = [q.to(cuda_device) for q in denorm(torch.rand(bs, 6).clamp(min=0.05), COMP_RANGES)]
    rev = [torch.rand(bs, device=cuda_device).requires_grad_(True) for _ in range(25)]
    drive = torch.rand(bs * 2, device=cuda_device) * 12
    noise = torch.randn(bs * 2, 12, L + taps - 1, device=cuda_device)

    def step():
        y = D.parametric_eq(x, SR, *eq)
        y = D.compressor(y, SR, *comp)
        y = D.noise_shaped_reverberation(y, SR, *rev, num_samples=L, num_bandpass_taps=taps, noise=noise)
        y = D.distortion(y, SR, drive)
        loss = y.pow(2).mean()
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            x.grad = None
            for r in rev:
                r.grad = None
            ref_loss = step()
    torch.cuda.current_stream().wait_stream(side)
    ref_gx, ref_gr = x.grad.clone(), rev[24].grad.clone()

    x.grad = None
    for r in rev:
        r.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = step()
    with torch.no_grad():
        x.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(loss, ref_loss, rtol=1e-6)
    assert torch.allclose(x.grad, ref_gx, rtol=1e-5, atol=1e-9)
    assert torch.allclose(rev[24].grad, ref_gr, rtol=1e-5, atol=1e-9)
    # new input values in the static buffers -> new results on replay
    with torch.no_grad():
        x.mul_(0.5)
    graph.replay()
    torch.cuda.synchronize()
    assert not torch.allclose(x.grad, ref_gx)


def test_processor_chain_graph_capture_with_device_noise(cuda_device):
    """the chain the reference trains (examples/style_transfer.py:150-154: eq -> comp -> reverb ->
    gain, all through Processor.process_normalized) captured fwd+bwd in ONE CUDA graph with the DEFAULT device-noise
    reverb: no host read happens under capture (device-side range check), and -- like the reference's torch.randn --
    every replay draws FRESH noise, because the Philox key is a device word re-drawn by torch's graph-safe generator.
    A replay right after torch.manual_seed(s) equals the eager step after the same seed."""
    import dasp_pytorch_b200 as D
    bs, n, L, taps = 4, 8192, 6000, 255
    torch.manual_seed(3)
    x = (torch.rand(bs, 2, n, device=cuda_device) * 2 - 1).requires_grad_(True)
    p = torch.rand(bs, 18 + 6 + 25 + 1, device=cuda_device)
    p[:, 22].clamp_(min=0.05)                                          # knee > 0
    p.requires_grad_(True)
    eq, comp, gain = D.ParametricEQ(SR), D.Compressor(SR), D.Gain(SR)
    rev = D.NoiseShapedReverb(SR)
    import functools
    from dasp_pytorch_b200 import functional as F
    rev._packed_path = (functools.partial(F.noise_shaped_reverberation_packed, num_samples=L, num_bandpass_taps=taps),
                        rev.process_fn)

    def step():
        y = eq.process_normalized(x, p[:, :18])
        y = comp.process_normalized(y, p[:, 18:24])
        y = rev.process_normalized(y, p[:, 24:49])
        y = gain.process_normalized(y, p[:, 49:50])
        loss = y.pow(2).mean()
        loss.backward()
        return y, loss

    def eager(seed):
        x.grad = None; p.grad = None
        torch.manual_seed(seed)
        y, loss = step()
        return y.detach().clone(), loss.detach().clone(), x.grad.clone(), p.grad.clone()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            eager(0)
    torch.cuda.current_stream().wait_stream(side)
    ref = eager(11)

    x.grad = None; p.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ys, losss = step()
    outs = []
    for seed in (11, None):
        with torch.no_grad():
            x.grad.zero_(); p.grad.zero_()
        if seed is not None:
            torch.manual_seed(seed)
        graph.replay()
        torch.cuda.synchronize()
        outs.append((ys.clone(), losss.clone(), x.grad.clone(), p.grad.clone()))
    # replay after manual_seed(11) == eager after manual_seed(11)
    for a, b in zip(outs[0], ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), float((a - b).abs().max())
    # the next replay drew different noise: wet part differs, nothing is NaN
    assert torch.isfinite(outs[1][0]).all()
    assert float((outs[1][0] - outs[0][0]).abs().max()) > 1e-3 * float(outs[0][0].abs().max())
    assert not (eq.range_violation() or comp.range_violation() or rev.range_violation())
    # device-side range check under capture: an out-of-range value poisons its item and raises the flag, no exception
    with torch.no_grad():
        p[1, 20] = 1.5
    graph.replay()
    torch.cuda.synchronize()
    assert comp.range_violation() and not comp.range_violation()        # reported once, then reset
    assert torch.isnan(ys[1]).any() and torch.isfinite(ys[0]).all()
    with torch.no_grad():
        p[1, 20] = 0.5
    with pytest.raises(ValueError, match="attack_ms"):                  # eager: the reference's ValueError
        bad = p.detach().clone(); bad[0, 20] = -0.2
        comp.process_normalized(x.detach(), bad[:, 18:24])
