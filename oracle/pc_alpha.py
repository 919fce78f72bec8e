"""alpha-generalisation of the ``DiagFFTPC`` apply (an EXTENSION: the upstream operator has no alpha).

Oracle code (test infrastructure).  **Parity unpinned by construction**: Control_Wave_PC.py is the
alpha = 1 block circulant of Wu & Liu and contains no Gamma_alpha scaling anywhere, so there is no
upstream behaviour to match for alpha != 1.  BASELINE config 5 / north_star nevertheless ask for a
"Gamma_alpha time-weight scaling" and an alpha sweep; the definition used here is the procedure-defined
one of SURVEY H1, which reduces exactly to :491-553 at alpha = 1:

    Gamma = diag(a^j), a = alpha^(1/N_t), j = time index
    x~ = Gamma x (both fields)  ->  ifft_t (:500-501)  ->  per frequency k the 2x2-block solve with the
    alpha-shifted symbols  l1 = (1 - a z)^2,  l2 = 1 + a^2 z^2,  z = e^{2 pi i k / N_t}  (:387-388 with
    z -> a z) and their conjugates in the adjoint block (:418-419)  ->  fft_t (:547-548)  ->  Gamma^-1.

Written as a matrix this is the inverse of

    P_alpha = [ C1a (x) M + dt^2/2 C2a (x) K          -c I (x) M             ]
              [ c I (x) M                   D1a (x) M + dt^2/2 D2a (x) K     ]

where C.a are the alpha-circulants of the stencils (1,-2,1) and (1,0,1) (wrap-around entries times alpha)
and D.a = Gamma^-1 (Ct.)^T Gamma with Ct. the ordinary circulant of the scaled stencil (c_m a^m): an
upper-triangular Toeplitz matrix with entries c_m a^(2m) and wrap-around entries c_m a^(2m) / alpha.  The
adjoint block is therefore NOT the transpose of the state block's alpha-circulant (no single Gamma scaling
block-diagonalises that pair), and the preconditioner degrades as alpha -> 0 (measured in
tests/test_oracle_alpha.py: GMRES 5 / 12 / 18 / 25 iterations for alpha = 1 / 0.5 / 0.1 / 0.01).

Three routes: ``ExplicitAlphaPC`` (the sparse matrix above, SuperLU), ``BlockAlphaPC`` (Gamma, FFT, one
sparse 2x2-block LU per frequency) and ``DiagFFTPCAlpha`` (the decoupled closed form the CUDA kernels
regenerate: one complex-symmetric Toeplitz tridiagonal matrix per frequency, two right-hand sides).
"""
import numpy as np
import scipy.fft as sfft
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem1d
from .pc_fast import thomas_toeplitz


def alpha_circulant(stencil, N_t, alpha):
    """Lower-Toeplitz matrix of ``stencil`` (c_0, c_1, ...) with wrap-around entries times alpha."""
    C = np.zeros((N_t, N_t))
    for m, v in enumerate(stencil):
        for i in range(N_t):
            j = i - m
            if j >= 0:
                C[i, j] += v
            else:
                C[i, j + N_t] += alpha * v
    return C


def adjoint_block(stencil, N_t, alpha):
    """Gamma^-1 (circulant of c_m a^m)^T Gamma: entries c_m a^(2m) above the diagonal, wraps / alpha."""
    a = alpha ** (1.0 / N_t)
    D = np.zeros((N_t, N_t))
    for m, v in enumerate(stencil):
        for i in range(N_t):
            j = i + m
            if j < N_t:
                D[i, j] += v * a ** (2 * m)
            else:
                D[i, j - N_t] += v * a ** (2 * m) / alpha
    return D


class ExplicitAlphaPC:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, alpha=1.0):
        self.N_x, self.N_t, self.n, self.m = N_x, N_t, N_x + 1, N_x - 1
        dt = T / N_t
        c = dt ** 2 / np.sqrt(gamma)
        Mi = fem1d.interior(fem1d.mass_full(N_x))
        Ki = fem1d.interior(fem1d.stiff_full(N_x))
        C1, C2 = (sp.csr_matrix(alpha_circulant(s, N_t, alpha)) for s in ((1.0, -2.0, 1.0), (1.0, 0.0, 1.0)))
        D1, D2 = (sp.csr_matrix(adjoint_block(s, N_t, alpha)) for s in ((1.0, -2.0, 1.0), (1.0, 0.0, 1.0)))
        It = sp.identity(N_t, format="csr")
        Puu = sp.kron(Mi, C1) + dt ** 2 / 2 * sp.kron(Ki, C2)
        Ppp = sp.kron(Mi, D1) + dt ** 2 / 2 * sp.kron(Ki, D2)
        self.P = sp.bmat([[Puu, -c * sp.kron(Mi, It)], [c * sp.kron(Mi, It), Ppp]], format="csc")
        self._lu = spla.splu(self.P)

    def apply(self, x):
        n, N_t, m = self.n, self.N_t, self.m
        x = np.asarray(x, dtype=complex).reshape(2, n, N_t)
        rhs = x[:, 1:-1, :].reshape(-1)
        sol = self._lu.solve(rhs.real) + 1j * self._lu.solve(rhs.imag)
        y = np.zeros((2, n, N_t), dtype=complex)
        y[:, 1:-1, :] = sol.reshape(2, m, N_t)
        return y.reshape(-1)


def symbols(N_t, alpha):
    """l1(k), l2(k) with z -> a z, formed without cancellation near theta = 0 and theta = pi/2."""
    k = np.arange(N_t)
    th = 2 * np.pi * k / N_t
    la = np.log(alpha) / N_t
    a = np.exp(la)
    one_m_a = -np.expm1(la)            # 1 - a
    one_m_a2 = -np.expm1(2 * la)       # 1 - a^2
    cth, sth = np.cos(th), np.sin(th)
    q = (one_m_a + 2 * a * np.sin(th / 2) ** 2) - 1j * a * sth        # 1 - a z
    l1 = q * q
    l2 = (one_m_a2 + 2 * a * a * cth * cth) + 1j * (2 * a * a * sth * cth)
    return l1, l2, a


class BlockAlphaPC:
    """Gamma scaling + FFT + one sparse LU of the coupled 2x2-block system per frequency."""

    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, alpha=1.0):
        self.N_x, self.N_t, self.n, self.m = N_x, N_t, N_x + 1, N_x - 1
        dt = T / N_t
        c = dt ** 2 / np.sqrt(gamma)
        Mi = fem1d.interior(fem1d.mass_full(N_x)).tocsc()
        Ki = fem1d.interior(fem1d.stiff_full(N_x)).tocsc()
        l1, l2, a = symbols(N_t, alpha)
        self.gam = a ** np.arange(N_t)
        self._lu = []
        for k in range(N_t):
            Pk = sp.bmat([[l1[k] * Mi + dt ** 2 / 2 * l2[k] * Ki, -c * Mi],
                          [c * Mi, np.conj(l1[k]) * Mi + dt ** 2 / 2 * np.conj(l2[k]) * Ki]], format="csc")
            self._lu.append(spla.splu(Pk))

    def apply(self, x):
        n, N_t, m = self.n, self.N_t, self.m
        x = np.asarray(x, dtype=complex).reshape(2, n, N_t)
        xh = sfft.ifft(x * self.gam, axis=2)
        w = np.zeros_like(xh)
        for k in range(N_t):
            s = self._lu[k].solve(np.concatenate([xh[0, 1:-1, k], xh[1, 1:-1, k]]))
            w[0, 1:-1, k], w[1, 1:-1, k] = s[:m], s[m:]
        return (sfft.fft(w, axis=2) / self.gam).reshape(-1)


def decoupled_coeffs(N_x, N_t, T, gamma, alpha):
    """Per-frequency quantities of the decoupled form (what ``make_coef<true>`` regenerates in-kernel).

    With l2 = |l2| e^{i phi}, mu = l1 e^{-i phi}, beta = Im mu, d = sqrt(beta^2 + c^2):
        T_k  = s M_int + kap K_int,  s = Re mu + i d,  kap = dt^2/2 |l2|     (T_- = conj T_+)
        rho_+ = g_+ uh + e ph,   rho_- = g_- uh - e ph,   g_+- = (d +- beta) / (2 d),  e = i c e^{i phi} / (2 d)
        zeta_+ = T_k^-1 rho_+,   zeta_- = conj(T_k^-1 conj rho_-)
        wh_u = e^{-i phi} (zeta_+ + zeta_-),   wh_p = i [ (beta - d) zeta_+ + (beta + d) zeta_- ] / c
    """
    dt = T / N_t
    c = dt ** 2 / np.sqrt(gamma)
    h = 1.0 / N_x
    l1, l2, a = symbols(N_t, alpha)
    al2 = np.abs(l2)
    eiphi = l2 / al2
    mu = l1 * np.conj(eiphi)
    beta = mu.imag
    d = np.sqrt(beta * beta + c * c)
    s = mu.real + 1j * d
    kap = dt ** 2 / 2 * al2
    return dict(off=s * (h / 6) - kap / h, diag=s * (2 * h / 3) + 2 * kap / h, gp=(d + beta) / (2 * d),
                gm=(d - beta) / (2 * d), e=1j * c * eiphi / (2 * d), eic=np.conj(eiphi), bmd=(beta - d) / c,
                bpd=(beta + d) / c, gam=a ** np.arange(N_t))


class DiagFFTPCAlpha:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, alpha=1.0, workers=None):
        self.N_x, self.N_t, self.n = N_x, N_t, N_x + 1
        self.cf = decoupled_coeffs(N_x, N_t, T, gamma, alpha)
        self.workers = workers

    def apply(self, x):
        cf, n, N_t = self.cf, self.n, self.N_t
        x = np.asarray(x, dtype=complex).reshape(2, n, N_t)
        xh = sfft.ifft(x * cf["gam"], axis=2, workers=self.workers)
        rp = cf["gp"] * xh[0] + cf["e"] * xh[1]
        rm = cf["gm"] * xh[0] - cf["e"] * xh[1]
        zp, zm = np.zeros_like(rp), np.zeros_like(rm)
        zp[1:-1] = thomas_toeplitz(cf["off"], cf["diag"], rp[1:-1])
        zm[1:-1] = np.conj(thomas_toeplitz(cf["off"], cf["diag"], np.conj(rm[1:-1])))
        w = np.empty_like(xh)
        w[0] = cf["eic"] * (zp + zm)
        w[1] = 1j * (cf["bmd"] * zp + cf["bpd"] * zm)
        return (sfft.fft(w, axis=2, workers=self.workers) / cf["gam"]).reshape(-1)
