"""numpy restatements of the algebra behind the reverb kernels (csrc/reverb.cu), at toy sizes.  They do not call the
library (that is what the `-m gpu` tests do); they pin the identities the kernels rely on, each next to the kernel
it explains, so that a maintainer can change a kernel against an executable statement of what it must compute."""
import numpy as np

RNG = np.random.default_rng(7)


def test_only_the_first_min_L_N_taps_reach_the_output():
    """y[n] = sum_{t<=n} IR[t] x[n-t], n < N  (reference functional.py:570-572 crops the convolution to N samples):
    taps beyond N never contribute, so the kernels synthesise Leff = min(L, N) taps only."""
    n, L = 50, 80
    x, ir = RNG.standard_normal(n), RNG.standard_normal(L)
    assert np.allclose(np.convolve(x, ir)[:n], np.convolve(x, ir[:n])[:n])


def test_packed_left_right_spectrum_and_its_mirror():
    """spectral_unit: left/right band signals are real, their spectra Hermitian; the packed spectrum G_l + i G_r at the
    mirror bin n1 - j equals conj(S_l) + i conj(S_r), so one draw per canonical bin fills two bins."""
    n1 = 48
    fl, fr = RNG.standard_normal(n1), RNG.standard_normal(n1)
    sl, sr = np.fft.fft(fl), np.fft.fft(fr)
    packed = np.fft.fft(fl + 1j * fr)
    j = np.arange(1, n1 // 2)
    assert np.allclose(packed[j], sl[j] + 1j * sr[j])
    assert np.allclose(packed[n1 - j], np.conj(sl[j]) + 1j * np.conj(sr[j]))


def test_polyphase_split_of_the_long_inverse_transform():
    """spectral_unit + ifft_shape_kernel: an inverse DFT of n1 = R nb points done as R-point DFTs over the bin classes,
    a twiddle, and R inverse DFTs of nb points:  f[R a + b] = sum_{j1<nb} Q_b[j1] e^{2 pi i j1 a / nb},
    Q_b[j1] = e^{2 pi i j1 b / n1} sum_{j2<R} G[j1 + nb j2] e^{2 pi i j2 b / R}."""
    for R, nb in ((1, 16), (3, 16), (6, 8), (7, 4)):
        n1 = R * nb
        G = RNG.standard_normal(n1) + 1j * RNG.standard_normal(n1)
        f = np.fft.ifft(G) * n1                                    # unnormalised inverse, like the kernels
        j1 = np.arange(nb)
        for b in range(R):
            q = np.exp(2j * np.pi * j1 * b / n1) * sum(G[j1 + nb * j2] * np.exp(2j * np.pi * j2 * b / R) for j2 in range(R))
            assert np.allclose(np.fft.ifft(q) * nb, f[R * np.arange(nb) + b])
        # the class step of the mirror class nb - j1 is e^{2 pi i / R} conj(step of class j1)
        assert np.allclose(np.exp(2j * np.pi * (nb - j1) / n1), np.exp(2j * np.pi / R) * np.conj(np.exp(2j * np.pi * j1 / n1)))


def test_periodic_filtered_noise_has_the_fir_autocovariance_on_the_window():
    """Device-noise mode: white noise filtered by the P+1-tap FIR h has autocovariance r[d] = sum_k h[k] h[k+d].  A
    PERIODIC white sequence of n1 >= Leff + P points filtered circularly has the circular autocovariance of h, which
    equals r[d] for every lag |d| <= n1 - P - 1, i.e. for all pairs of samples inside a window of Leff points."""
    P, leff = 6, 20
    n1 = leff + P
    h = RNG.standard_normal(P + 1)
    hp = np.zeros(n1)
    hp[:P + 1] = h
    circ = np.real(np.fft.ifft(np.abs(np.fft.fft(hp)) ** 2))        # covariance of the circularly filtered process
    lin = np.correlate(h, h, mode="full")[P:]                       # r[0..P]
    for d in range(leff):                                           # every lag between two samples of the window
        assert np.isclose(circ[d], lin[d] if d <= P else 0.0)


def test_uniformly_partitioned_overlap_save_convolution():
    """x_fft_kernel / partition_mac_kernel / ifft_mix_kernel: block i of y is the last B samples of
    IFFT(sum_{j<=i} X[i-j] H[j]) with X[i] = FFT(x[(i-1)B : (i+1)B]) and H[j] = FFT(IR[jB : (j+1)B] zero padded)."""
    B, n, leff = 8, 45, 27
    x, ir = RNG.standard_normal(n), RNG.standard_normal(leff)
    I, J = -(-n // B), -(-leff // B)
    xp = np.concatenate([np.zeros(B), x, np.zeros((I + 1) * B - n)])
    X = [np.fft.fft(xp[i * B:(i + 2) * B]) for i in range(I)]       # xp index i*B is sample (i-1)*B
    H = [np.fft.fft(np.concatenate([np.pad(ir, (0, J * B - leff))[j * B:(j + 1) * B], np.zeros(B)])) for j in range(J)]
    y = np.concatenate([np.real(np.fft.ifft(sum(X[i - j] * H[j] for j in range(J) if j <= i)))[B:] for i in range(I)])[:n]
    assert np.allclose(y, np.convolve(x, ir)[:n])


def test_two_real_convolutions_through_one_packed_complex_transform():
    """partition_mac_kernel: left and right channel have different IRs, so the packed spectra Z = FFT(a + i b) are
    untangled through A[f] = (Z[f] + conj(Z[-f]))/2, B[f] = (Z[f] - conj(Z[-f]))/(2i), multiplied per channel and
    re-packed as L + i R; the inverse transform then carries both channels in its real / imaginary part."""
    m = 32
    xl, xr, hl, hr = (RNG.standard_normal(m) for _ in range(4))

    def untangle(z):
        zm = np.conj(np.roll(z[::-1], 1))
        return (z + zm) / 2, (z - zm) / 2j

    XL, XR = untangle(np.fft.fft(xl + 1j * xr))
    HL, HR = untangle(np.fft.fft(hl + 1j * hr))
    y = np.fft.ifft(XL * HL + 1j * (XR * HR))
    circ = lambda a, b: np.real(np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)))
    assert np.allclose(y.real, circ(xl, hl)) and np.allclose(y.imag, circ(xr, hr))


def test_envelope_is_geometric_in_the_tap_index():
    """ifft_shape_kernel evaluates exp(-(10 decay + 1) tt) at every 4th of a thread's taps and steps geometrically in
    between: tt = linspace(0, 1, L) is linear in the tap index up to fp32 rounding (reference functional.py:561)."""
    L, R, c, t = 96000, 6, 2, 137
    tt = np.linspace(0.0, 1.0, L, dtype=np.float32)
    rr = np.float32(-(10.0 * 0.7 + 1.0))
    tau = R * (t + 512 * np.arange(16)) + c
    exact = np.exp(rr * tt[tau], dtype=np.float32)
    rho = np.exp(rr * np.float32(1.0 / (L - 1)) * np.float32(512 * R), dtype=np.float32)
    stepped = exact.copy()
    for q in range(16):
        if q % 4:
            stepped[q] = stepped[q - 1] * rho
    assert np.max(np.abs(stepped / exact - 1)) < 2e-6
