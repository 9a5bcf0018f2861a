"""numpy restatements (fp64, toy sizes) of the algebra behind the recurrence kernels csrc/biquad.cu and
csrc/dynamics.cu: the sigma-form biquad, the chunked scan with matrix powers, and the state-space adjoint used by the
backward.  They do not call the library; they pin what the kernels must compute."""
import numpy as np
from scipy.signal import lfilter

RNG = np.random.default_rng(11)


def sigma_coeffs(b0, b1, b2, a1, a2):
    sg = -a1 / 2
    be1 = b1 - a1 * b0
    return sg, sg * sg - a2, be1, (b2 - a2 * b0) + sg * be1, b0


def sigma_run(u, cf, s=(0.0, 0.0)):
    sg, q, be1, B2, b0 = cf
    s1, s2 = s
    y, states = np.empty_like(u), np.empty((len(u), 2))
    for n, un in enumerate(u):
        states[n] = (s1, s2)
        y[n] = s1 + b0 * un
        s1, s2 = sg * s1 + s2 + be1 * un, q * s1 + sg * s2 + B2 * un
    return y, states, (s1, s2)


def test_sigma_form_is_the_direct_form_biquad():
    """biquad.cu header: s1' = sg s1 + s2 + be1 u, s2' = q s1 + sg s2 + B2 u, y = s1 + b0 u with sg = -a1/2,
    q = sg^2 - a2, be1 = b1 - a1 b0, B2 = (b2 - a2 b0) + sg be1 realises (b0 + b1 z^-1 + b2 z^-2)/(1 + a1 z^-1 + a2 z^-2)."""
    b, a = np.array([0.9, -1.7, 0.85]), np.array([1.0, -1.92, 0.93])
    u = RNG.standard_normal(300)
    y, _, _ = sigma_run(u, sigma_coeffs(b[0], b[1], b[2], a[1], a[2]))
    assert np.allclose(y, lfilter(b, a, u), atol=1e-10)


def test_chunked_scan_with_matrix_powers():
    """eq_fwd_kernel: every thread runs its E samples from a ZERO state, the end states are combined by a scan whose
    operator is the matrix power A^E (A is constant in time), and the outputs are fixed up with y[j] += (A^j c_in)_1."""
    cf = sigma_coeffs(0.9, -1.7, 0.85, -1.92, 0.93)
    sg, q = cf[0], cf[1]
    A = np.array([[sg, 1.0], [q, sg]])
    E, T = 5, 8
    u = RNG.standard_normal(E * T)
    y_ref, _, _ = sigma_run(u, cf)
    ends, local = [], []
    for t in range(T):
        y, _, end = sigma_run(u[t * E:(t + 1) * E], cf)
        local.append(y)
        ends.append(np.array(end))
    AE = np.linalg.matrix_power(A, E)
    carry, out = np.zeros(2), []
    for t in range(T):                                              # (the kernel does this as a log-step scan)
        fix = np.array([(np.linalg.matrix_power(A, j) @ carry)[0] for j in range(E)])
        out.append(local[t] + fix)
        carry = AE @ carry + ends[t]
    assert np.allclose(np.concatenate(out), y_ref, atol=1e-10)


def test_state_space_adjoint_gradients():
    """eq_bwd_kernel: lam[n] = A^T lam[n+1] + (g[n], 0); gu[n] = be1 lam1[n+1] + B2 lam2[n+1] + b0 g[n];
    d sg = sum lam[n+1].s[n], d q = sum lam2[n+1] s1[n], d be1 = sum lam1[n+1] u[n], d B2 = sum lam2[n+1] u[n],
    d b0 = sum g[n] u[n] -- checked against central finite differences of L = sum w y."""
    cf = np.array(sigma_coeffs(0.9, -1.7, 0.85, -1.6, 0.7))
    n = 60
    u, w = RNG.standard_normal(n), RNG.standard_normal(n)
    _, S, _ = sigma_run(u, cf)
    sg, q, be1, B2, b0 = cf
    lam = np.zeros((n + 1, 2))                                      # lam[n] = dL/d state entering sample n
    for k in range(n - 1, -1, -1):
        l1, l2 = lam[k + 1]
        lam[k] = (sg * l1 + q * l2 + w[k], l1 + sg * l2)
    gu = be1 * lam[1:, 0] + B2 * lam[1:, 1] + b0 * w
    grads = np.array([np.sum(lam[1:, 0] * S[:, 0] + lam[1:, 1] * S[:, 1]), np.sum(lam[1:, 1] * S[:, 0]),
                      np.sum(lam[1:, 0] * u), np.sum(lam[1:, 1] * u), np.sum(w * u)])
    loss = lambda c, uu: float(np.sum(w * sigma_run(uu, c)[0]))
    eps = 1e-6
    for i in range(5):
        d = np.zeros(5)
        d[i] = eps
        assert np.isclose(grads[i], (loss(cf + d, u) - loss(cf - d, u)) / (2 * eps), rtol=1e-5, atol=1e-7), i
    for k in (0, 7, n - 1):
        d = np.zeros(n)
        d[k] = eps
        assert np.isclose(gu[k], (loss(cf, u + d) - loss(cf, u - d)) / (2 * eps), rtol=1e-5, atol=1e-7)


def test_one_pole_smoother_as_an_affine_scan():
    """dynamics.cu: the attack smoother y[n] = a y[n-1] + (1 - a) g[n] (reference functional.py:343-376 via
    lfilter_via_fsm) is the affine map v -> a v + c; chunks compose as (a^E, local end value), so the same
    zero-state-pass + power scan + fix-up  y[j] += a^(j+1) carry  applies with scalars instead of 2x2 matrices."""
    a, E, T = 0.93, 7, 6
    g = RNG.standard_normal(E * T)
    ref = lfilter([1 - a], [1.0, -a], g)
    carry, out = 0.0, []
    for t in range(T):
        loc = lfilter([1 - a], [1.0, -a], g[t * E:(t + 1) * E])
        out.append(loc + carry * a ** np.arange(1, E + 1))
        carry = a ** E * carry + loc[-1]
    assert np.allclose(np.concatenate(out), ref, atol=1e-12)
