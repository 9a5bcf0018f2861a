"""CPU restatement of dasp_pytorch.functional's audio-processor hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Each function restates
the arithmetic of one reference symbol (cited ``file:line`` relative to the
upstream repo ``csteinmetz1/dasp-pytorch`` @ c9ae0126) on CPU torch tensors, in
whatever float dtype the caller passes (fp32 = "what the reference prints",
fp64 = the arbiter).  Everything is built from
differentiable torch ops so ``torch.autograd`` on the oracle gives the
gradient every CUDA backward kernel is checked against.

Pinning: the reference ships no tests and no golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the authoring
container by ``oracle/make_golden.py`` (which imports ``/root/reference``) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures everywhere (the GPU box has no ``/root/reference``).

``expander`` has no reference implementation (``functional.py:402-403`` is a
stub that raises) -- PARITY UNPINNED for that one op: it is pinned only to the
definition written here.
"""

from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import scipy.signal
import torch

LN10_OVER_20 = math.log(10.0) / 20.0

# --------------------------------------------------------------------------------------
# pointwise processors
# --------------------------------------------------------------------------------------


def gain(x: torch.Tensor, sample_rate, gain_db: torch.Tensor) -> torch.Tensor:
    """y = x * 10^(gain_db/20), one gain per batch item (functional.py:10-29).

    ``gain_db`` must hold exactly ``bs`` elements (the reference does
    ``gain_db.view(bs, 1, 1)``, functional.py:25).
    """
    bs = x.shape[0]
    k = torch.pow(10.0, gain_db.reshape(bs, 1, 1) / 20.0)
    return x * k


def distortion(x: torch.Tensor, sample_rate, drive_db: torch.Tensor) -> torch.Tensor:
    """y = tanh(x * 10^(drive_db/20)) (functional.py:65-78).

    The reference reshapes the drive to ``(bs, chs, -1)`` (functional.py:78), so
    it must hold ``bs*chs`` elements (one per row), or ``bs`` when chs == 1.
    """
    bs, chs, _ = x.shape
    k = torch.pow(10.0, drive_db.reshape(bs, chs, -1) / 20.0)
    return torch.tanh(x * k)


# --------------------------------------------------------------------------------------
# stereo mixing processors
# --------------------------------------------------------------------------------------


def stereo_widener(x: torch.Tensor, sample_rate, width: torch.Tensor) -> torch.Tensor:
    """Mid/side widener (functional.py:580-604), restated without the in-place ops.

    mid=(L+R)/sqrt2 * 2(1-w), side=(L-R)/sqrt2 * 2w, left=(mid+side)/sqrt2, right=(mid-side)/sqrt2.
    ``width`` is reshaped to ``(bs, 1)`` (the reference's broadcast only works for that shape).
    """
    bs = x.shape[0]
    w = width.reshape(bs, 1)
    r2 = math.sqrt(2.0)
    mid = (x[:, 0] + x[:, 1]) / r2 * (2.0 * (1.0 - w))
    side = (x[:, 0] - x[:, 1]) / r2 * (2.0 * w)
    return torch.stack(((mid + side) / r2, (mid - side) / r2), dim=-2)


def stereo_panner(x: torch.Tensor, sample_rate, pan: torch.Tensor) -> torch.Tensor:
    """Constant-power-ish panner (functional.py:607-636): returns (bs, 2, tracks, N)."""
    bs, tracks, _ = x.shape
    theta = pan.reshape(bs, tracks) * (math.pi / 2)
    lg = torch.sqrt(((math.pi / 2) - theta) * (2 / math.pi) * torch.cos(theta))
    rg = torch.sqrt(theta * (2 / math.pi) * torch.sin(theta))
    gains = torch.stack((lg, rg), dim=1).unsqueeze(-1)          # (bs, 2, tracks, 1)
    return x.unsqueeze(1) * gains


def stereo_bus(x: torch.Tensor, sample_rate, send_db: torch.Tensor) -> torch.Tensor:
    """Stereo bus (functional.py:32-62): (bs, 2, tracks, N) x sends in dB -> (bs, 2, N)."""
    bs, chs, tracks, _ = x.shape
    assert chs == 2
    s = torch.pow(10.0, send_db.reshape(bs, 1, tracks, 1) / 20.0)
    return (x * s).sum(dim=2)


# --------------------------------------------------------------------------------------
# biquad design + parametric EQ
# --------------------------------------------------------------------------------------

EQ_SECTION_KINDS = ("low_shelf", "peaking", "peaking", "peaking", "peaking", "high_shelf")


def biquad_section(gain_db, cutoff_freq, q_factor, sample_rate, kind: str):
    """RBJ-cookbook biquad, normalised by a0 (signal.py:242-306).

    Inputs are ``(bs, 1)`` tensors; returns ``b, a`` each ``(bs, 3)`` with
    ``a[:, 0] == 1``.  Shelf/peaking formulas: signal.py:261-281.
    """
    amp = torch.pow(10.0, gain_db / 40.0)
    w0 = 2.0 * math.pi * (cutoff_freq / sample_rate)
    alpha = torch.sin(w0) / (2.0 * q_factor)
    cw = torch.cos(w0)
    if kind == "peaking":
        num = (1.0 + alpha * amp, -2.0 * cw, 1.0 - alpha * amp)
        den = (1.0 + alpha / amp, -2.0 * cw, 1.0 - alpha / amp)
    elif kind in ("low_shelf", "high_shelf"):
        # the two shelves differ only in the sign of the cos(w0) terms
        sgn = 1.0 if kind == "low_shelf" else -1.0
        s = 2.0 * torch.sqrt(amp) * alpha
        ap1, am1 = amp + 1.0, amp - 1.0
        num = (
            amp * (ap1 - sgn * am1 * cw + s),
            sgn * 2.0 * amp * (am1 - sgn * ap1 * cw),
            amp * (ap1 - sgn * am1 * cw - s),
        )
        den = (
            ap1 + sgn * am1 * cw + s,
            -sgn * 2.0 * (am1 + sgn * ap1 * cw),
            ap1 + sgn * am1 * cw - s,
        )
    else:
        raise ValueError(f"unknown biquad kind {kind!r}")
    a0 = den[0]
    b = torch.cat([c / a0 for c in num], dim=-1)
    a = torch.cat([c / a0 for c in den], dim=-1)
    return b, a


def eq_sections(sample_rate, params):
    """18 EQ parameters -> SOS tensor ``(bs, 6, 6)`` rows ``[b0 b1 b2 1 a1 a2]``.

    Section order low-shelf, band0..band3 (peaking), high-shelf
    (functional.py:213-265).  ``params`` is the list of 18 tensors in signature
    order (gain, cutoff, q per section).
    """
    assert len(params) == 18
    rows = []
    for k, kind in enumerate(EQ_SECTION_KINDS):
        g, f, q = (p.reshape(-1, 1) for p in params[3 * k : 3 * k + 3])
        b, a = biquad_section(g, f, q, sample_rate, kind)
        rows.append(torch.cat([b, a], dim=-1))
    return torch.stack(rows, dim=1)


def _fsm_size(n: int, tail: int = 0) -> int:
    """FFT size of the frequency-sampling method: 2^ceil(log2(2n-1)) (signal.py:109,150).

    ``tail`` > 0 is a test-only extension: enlarge the grid to at least ``n + tail`` points so
    that an impulse response up to ``tail`` samples long does not wrap around (the reference's
    own grid leaves ``n_fft - n`` samples, which time-aliases long decays at small ``n``).
    With a large tail the result is the true zero-state IIR *and* stays differentiable.
    """
    size = 1 << max(0, math.ceil(math.log2(max(1, 2 * n - 1))))
    if tail > 0:
        size = max(size, 1 << math.ceil(math.log2(n + tail)))
    return size


def sos_frequency_sampling(sos: torch.Tensor, x: torch.Tensor, tail: int = 0) -> torch.Tensor:
    """Cascade filtering the way the reference does it: sample H on an FFT grid.

    H = prod_k rfft(b_k, n)/rfft(a_k, n) (signal.py:7-11, 14-32), one response per
    batch item shared by all channels (signal.py:157-158), then
    irfft(rfft(x, n) * H)[..., :N] (signal.py:35-39, 161-164).
    """
    n = _fsm_size(x.shape[-1], tail)
    resp = None
    for k in range(sos.shape[1]):
        hk = torch.fft.rfft(sos[:, k, :3], n) / torch.fft.rfft(sos[:, k, 3:], n)
        resp = hk if resp is None else resp * hk
    while resp.dim() < x.dim():
        resp = resp.unsqueeze(1)
    y = torch.fft.irfft(torch.fft.rfft(x, n) * resp, n)
    return y[..., : x.shape[-1]]


def sos_recursion_truth(sos: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Independent truth: zero-state time-domain recursion in fp64 (scipy sosfilt).

    Not differentiable; used to show the frequency-sampling result equals the true
    IIR and to arbitrate the fp32 noise of the FSM path.
    """
    s = sos.detach().double().numpy()
    xin = x.detach().double().numpy()
    out = np.empty_like(xin)
    for b in range(xin.shape[0]):
        out[b] = scipy.signal.sosfilt(s[b if s.shape[0] > 1 else 0], xin[b], axis=-1)
    return torch.from_numpy(out)


def parametric_eq(x: torch.Tensor, sample_rate, *params, method: str = "fsm", fsm_tail: int = 0) -> torch.Tensor:
    """Six-section parametric EQ (functional.py:118-272).

    ``params``: the 18 tensors of the reference signature, any shape with ``bs``
    elements, or one element broadcast over the batch
    (examples/virtual_analog.py:204-206).  Integer cutoffs are accepted
    (examples/demo.py:44) -- they are promoted by the arithmetic like in torch.
    ``method="fsm"`` follows the reference; ``"recursion"`` is the fp64 scipy truth.
    """
    params = [p if torch.is_tensor(p) else torch.as_tensor(p) for p in params]
    sos = eq_sections(sample_rate, params)
    if method == "recursion":
        return sos_recursion_truth(sos, x).to(x.dtype)
    sos = sos.to(x.dtype) if not sos.is_floating_point() else sos
    return sos_frequency_sampling(sos.type_as(x), x, fsm_tail)


# --------------------------------------------------------------------------------------
# dynamics: compressor (reference) and expander (defined here, parity unpinned)
# --------------------------------------------------------------------------------------


def _attack_coefficient(attack_ms, sample_rate):
    """alpha = exp(-ln 9 / (sr * attack_ms / 1000)) (functional.py:339-342)."""
    return torch.exp(-math.log(9.0) / (sample_rate * (attack_ms / 1e3)))


def _one_pole_fsm(gc: torch.Tensor, alpha: torch.Tensor, tail: int = 0) -> torch.Tensor:
    """s[n] = alpha s[n-1] + (1-alpha) gc[n] by frequency sampling.

    b = [1-alpha, 0], a = [1, -alpha] (functional.py:372-379) pushed through the
    same FFT-grid division as the EQ (signal.py:95-133).  ``gc`` is ``(bs, 1, N)``.
    """
    bs = gc.shape[0]
    al = alpha.reshape(bs, 1)
    zero = torch.zeros_like(al)
    b = torch.cat([1.0 - al, zero], dim=-1)
    a = torch.cat([torch.ones_like(al), -al], dim=-1)
    n = _fsm_size(gc.shape[-1], tail)
    resp = (torch.fft.rfft(b, n) / torch.fft.rfft(a, n)).unsqueeze(1)
    return torch.fft.irfft(torch.fft.rfft(gc, n) * resp, n)[..., : gc.shape[-1]]


def one_pole_recursion_truth(gc: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """fp64 scipy lfilter truth for the attack smoother (not differentiable)."""
    g = gc.detach().double().numpy()
    al = alpha.detach().double().reshape(-1).numpy()
    out = np.empty_like(g)
    for i in range(g.shape[0]):
        out[i] = scipy.signal.lfilter([1.0 - al[i]], [1.0, -al[i]], g[i], axis=-1)
    return torch.from_numpy(out)


def _dynamics(x, sample_rate, threshold_db, ratio, attack_ms, knee_db, makeup_gain_db,
              eps, lookahead_samples, curve, smoother, fsm_tail=0):
    bs, chs, n = x.shape
    side = x.sum(dim=1, keepdim=True)                       # functional.py:328
    t = threshold_db.reshape(bs, 1, 1)
    r = ratio.reshape(bs, 1, 1)
    w = knee_db.reshape(bs, 1, 1)
    m = makeup_gain_db.reshape(bs, 1, 1)
    alpha = _attack_coefficient(attack_ms.reshape(bs, 1, 1), sample_rate)
    level_db = 20.0 * torch.log10(side.abs().clamp(min=eps))  # functional.py:347
    gc = curve(level_db, t, r, w)                           # static gain computer, dB
    if smoother == "recursion":
        sm = one_pole_recursion_truth(gc, alpha).to(x.dtype)
    else:
        sm = _one_pole_fsm(gc, alpha, fsm_tail)             # functional.py:380
    if lookahead_samples > 0:                               # functional.py:383-385
        delayed = torch.zeros_like(x)
        delayed[..., lookahead_samples:] = x[..., : n - lookahead_samples]
        x = delayed
    return x * torch.pow(10.0, (sm + m) / 20.0)             # functional.py:388-394


def _compressor_curve(level_db, t, r, w):
    """Soft-knee downward compression, expressed as gain g_c = x_sc - x_db.

    functional.py:350-369: unchanged below T-W/2; quadratic knee for
    T-W/2 <= x_db <= T+W/2; slope 1/R above.  Built with ``where`` instead of the
    reference's masked writes; the values (and the fact that W == 0 poisons the
    gradient with 0/0) are the same.
    """
    slope = 1.0 / r - 1.0
    d = level_db - t + w / 2.0
    knee = slope * d * d / (2.0 * w)
    above = (t - level_db) * (1.0 - 1.0 / r)
    in_knee = (level_db >= t - w / 2.0) & (level_db <= t + w / 2.0)
    over = level_db > t + w / 2.0
    out = torch.zeros_like(level_db)
    out = torch.where(in_knee, knee.expand_as(level_db), out)
    out = torch.where(over, above.expand_as(level_db), out)
    return out


def compressor(x, sample_rate, threshold_db, ratio, attack_ms, release_ms, knee_db,
               makeup_gain_db, eps: float = 1e-8, lookahead_samples: int = 0,
               smoother: str = "fsm", fsm_tail: int = 0):
    """Feed-forward compressor, attack-only smoothing (functional.py:275-399).

    ``release_ms`` is accepted and unused, exactly like the reference
    (functional.py:333,340,343-344); its gradient is None.
    """
    return _dynamics(x, sample_rate, threshold_db, ratio, attack_ms, knee_db, makeup_gain_db,
                     eps, lookahead_samples, _compressor_curve, smoother, fsm_tail)


def _expander_curve(level_db, t, r, w):
    """Soft-knee downward EXPANSION gain (dB), Giannoulis et al. 2012 static curve.

    g = (R-1)(x_db - T)                 for x_db <  T - W/2
      = (1-R)(x_db - T - W/2)^2 / (2W)  for T - W/2 <= x_db <= T + W/2
      = 0                               for x_db >  T + W/2
    (continuous at both knee edges).  PARITY UNPINNED: the reference has no expander.
    """
    d = level_db - t - w / 2.0
    knee = (1.0 - r) * d * d / (2.0 * w)
    below = (r - 1.0) * (level_db - t)
    in_knee = (level_db >= t - w / 2.0) & (level_db <= t + w / 2.0)
    under = level_db < t - w / 2.0
    out = torch.zeros_like(level_db)
    out = torch.where(in_knee, knee.expand_as(level_db), out)
    out = torch.where(under, below.expand_as(level_db), out)
    return out


def expander(x, sample_rate, threshold_db, ratio, attack_ms, release_ms, knee_db,
             makeup_gain_db, eps: float = 1e-8, lookahead_samples: int = 0,
             smoother: str = "fsm", fsm_tail: int = 0):
    """Downward expander with the compressor's signature and smoother (new op)."""
    return _dynamics(x, sample_rate, threshold_db, ratio, attack_ms, knee_db, makeup_gain_db,
                     eps, lookahead_samples, _expander_curve, smoother, fsm_tail)


# --------------------------------------------------------------------------------------
# noise-shaped reverberation
# --------------------------------------------------------------------------------------

OCTAVE_CENTRES = (31.5, 63.0, 125.0, 250.0, 500.0, 1000.0, 2000.0, 4000.0, 8000.0, 16000.0)


@lru_cache(maxsize=16)
def _filterbank_np(num_taps: int, sample_rate: float) -> np.ndarray:
    rows = [scipy.signal.firwin(num_taps, 12, fs=sample_rate)]           # signal.py:60-64
    for fc in OCTAVE_CENTRES:                                            # signal.py:69-78
        lo = fc / np.sqrt(2)
        hi = min(fc * np.sqrt(2), 0.999 * sample_rate / 2)
        rows.append(scipy.signal.firwin(num_taps, [lo, hi], fs=sample_rate, pass_zero=False))
    rows.append(scipy.signal.firwin(num_taps, 18000, fs=sample_rate, pass_zero=False))  # :84
    # cast to fp32 like the reference (signal.py:65,79,85); the time flip it applies is a
    # no-op for these symmetric (linear-phase) filters.
    return np.stack(rows).astype(np.float32)[:, ::-1].copy()


def octave_filterbank(num_taps: int, sample_rate: float) -> torch.Tensor:
    """12 Hamming-window FIRs: LP 12 Hz, 10 octave band-passes, HP 18 kHz (signal.py:42-92).

    Returns ``(12, num_taps)`` float32 (the reference returns ``(12, 1, num_taps)``).
    """
    return torch.from_numpy(_filterbank_np(int(num_taps), float(sample_rate)))


def reverb_noise(bs: int, num_samples: int, num_bandpass_taps: int, seed: int) -> torch.Tensor:
    """The one RNG draw of the reference call (functional.py:547-548), reproduced.

    ``torch.manual_seed(seed)`` then ``randn(bs*2, 12, num_samples + taps - 1)`` fp32 on
    the CPU generator; row ``b*2 + c`` belongs to item b, channel c (functional.py:558).
    """
    torch.manual_seed(seed)
    return torch.randn(bs * 2, 12, num_samples + num_bandpass_taps - 1)


def _conv_valid_fft(sig: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """'valid' linear convolution along the last dim via FFT (== conv1d with flipped taps)."""
    n_sig, n_tap = sig.shape[-1], taps.shape[-1]
    n = 1 << math.ceil(math.log2(n_sig + n_tap - 1))
    full = torch.fft.irfft(torch.fft.rfft(sig, n) * torch.fft.rfft(taps, n), n)
    return full[..., n_tap - 1 : n_sig]


def noise_shaped_reverberation(x, sample_rate, *params, num_samples: int = 65536,
                               num_bandpass_taps: int = 1023, noise=None,
                               method: str = "fft"):
    """Filtered-noise reverb (functional.py:406-577).

    ``params`` = 12 band gains, 12 band decays, mix (25 tensors of ``bs`` elements).
    ``noise`` is the ``(bs*2, 12, num_samples + taps - 1)`` white-noise tensor the
    reference draws internally (functional.py:548); pass ``reverb_noise(...)`` to
    reproduce a seeded reference call.  ``method="direct"`` uses time-domain
    ``conv1d`` like the reference (functional.py:551-556, 570-572) and is what the
    CPU baseline times; ``"fft"`` is the mathematically identical fast form
     used for checking at sizes the direct form cannot finish.
    """
    assert num_bandpass_taps % 2 == 1, "num_bandpass_taps must be odd"   # functional.py:487
    assert len(params) == 25
    bs, chs, n = x.shape
    assert chs <= 2, "only mono/stereo signals are supported"            # functional.py:490
    if chs == 1:                                                          # functional.py:493-495
        x = x.repeat(1, 2, 1)
    gains = torch.stack([p.reshape(bs) for p in params[0:12]], dim=1).reshape(bs, 1, 12, 1)
    decays = torch.stack([p.reshape(bs) for p in params[12:24]], dim=1).reshape(bs, 1, 12, 1)
    mix = params[24].reshape(bs, 1, 1)
    fb = octave_filterbank(num_bandpass_taps, sample_rate).to(x.dtype)    # (12, taps)
    if noise is None:
        noise = torch.randn(bs * 2, 12, num_samples + num_bandpass_taps - 1)
    noise = noise.to(x.dtype)

    if method == "direct":
        shaped = torch.nn.functional.conv1d(noise, fb.unsqueeze(1), groups=12)
    else:
        # conv1d is a correlation; the filters are symmetric so either orientation is equal,
        # but flip anyway to restate conv1d exactly.
        shaped = _conv_valid_fft(noise, torch.flip(fb, dims=[-1]).unsqueeze(0))
    shaped = shaped.reshape(bs, 2, 12, num_samples)                      # functional.py:558

    # the reference builds the time axis in fp32 and then casts it (functional.py:561)
    t = torch.linspace(0, 1, steps=num_samples, dtype=torch.float32).to(x.dtype)
    env = torch.exp(-(decays * 10.0 + 1.0) * t.reshape(1, 1, 1, -1))     # functional.py:562-563
    ir = (shaped * env * gains).mean(dim=2)                              # functional.py:564-567

    if method == "direct":
        # one grouped correlation over all (item, channel) rows == the reference's vmap(conv1d(groups=2))
        xp = torch.nn.functional.pad(x, (num_samples - 1, 0)).reshape(1, bs * 2, n + num_samples - 1)
        wet = torch.nn.functional.conv1d(xp, torch.flip(ir, dims=[-1]).reshape(bs * 2, 1, num_samples),
                                         groups=bs * 2).reshape(bs, 2, n)
    else:
        m = 1 << math.ceil(math.log2(n + num_samples - 1))
        wet = torch.fft.irfft(torch.fft.rfft(x, m) * torch.fft.rfft(ir, m), m)[..., :n]
    return (1.0 - mix) * x + mix * wet                                    # functional.py:575
