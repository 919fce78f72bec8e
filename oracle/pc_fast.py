"""Division-free decoupled form of the ``DiagFFTPC`` apply (same operator).

Oracle code (test infrastructure); also the timed CPU baseline of ``bench.py``.

Algebra (derived from Control_Wave_PC.py:387-388, :418-419, :460-473, :491-553;
closed forms of pre_cond.py:32-38).  With theta = 2 pi k / N_t, z = e^{i theta}:

    lambda_2 = 2 cos(theta) z,   lambda_1 = -4 sin^2(theta/2) z,
    sigma = sign(cos theta),     c = dt^2 / sqrt(gamma),   kappa = dt^2 cos(theta)

the unitary eigenvector matrix of the 2x2 block is
S_k = 2^{-1/2} [[1, 1], [-i sigma / z, +i sigma / z]] and the whole apply becomes

    (uh, ph)  = ifft_t(x_u), ifft_t(x_p)                         (:500-501)
    rho_+     = ( uh / z + i sigma ph ) / 2
    rho_-     = ( uh / z - i sigma ph ) / 2
    zeta_+    = Tt_k^{-1} rho_+ ,   zeta_- = conj(Tt_k)^{-1} rho_-   (interior nodes)
    Tt_k      = (s_re + i c sigma) M_int + kappa K_int,   s_re = -4 sin^2(theta/2)
    wh_u      = zeta_+ + zeta_- ,   wh_p = -i sigma z (zeta_+ - zeta_-)
    y         = fft_t(wh)                                         (:547-548)

``Tt_k`` is complex-symmetric Toeplitz tridiagonal with off-diagonal
``a = (s_re + i c sigma) h/6 - kappa/h`` and diagonal ``b = (s_re + i c sigma) 2h/3 +
2 kappa / h``.  No division by lambda_2 occurs, so N_t divisible by 4
(lambda_2(N_t/4) = 0) needs no special case.  Boundary-node outputs are 0.
"""
import os

import numpy as np
import scipy.fft as sfft

from . import eigs


def tridiag_coeffs(N_x, N_t, T, gamma, dtype=np.float64):
    """(a_k, b_k) of Tt_k for every k (complex arrays of length N_t)."""
    cf = eigs.closed_form(N_t, T / N_t if dtype == np.float64 else dtype(T) / dtype(N_t),
                          gamma, dtype)
    h = dtype(1) / dtype(N_x)
    s = cf["s_re"] + 1j * (cf["c"] * cf["sigma"])
    a = s * (h / dtype(6)) - cf["kappa"] / h
    b = s * (dtype(2) * h / dtype(3)) + dtype(2) * cf["kappa"] / h
    return a, b, cf


def thomas_toeplitz(a, b, rhs):
    """Solve tridiag(a_k, b_k, a_k) z = rhs for every k.

    ``rhs`` has shape (m, N_t): row = interior node, column = frequency.
    Vectorised over the frequency axis; LU without pivoting (what a banded
    direct solve of a Toeplitz tridiagonal does).
    """
    m = rhs.shape[0]
    cp = np.empty_like(rhs)
    d = np.empty_like(rhs)
    inv = 1 / b
    cp[0] = a * inv
    d[0] = rhs[0] * inv
    for i in range(1, m):
        inv = 1 / (b - a * cp[i - 1])
        cp[i] = a * inv
        d[i] = (rhs[i] - a * d[i - 1]) * inv
    for i in range(m - 2, -1, -1):
        d[i] -= cp[i] * d[i + 1]
    return d


class DiagFFTPCFast:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, dtype=np.float64, workers=None,
                 solver=None):
        self.N_x, self.N_t, self.n = N_x, N_t, N_x + 1
        self.T, self.gamma, self.dtype = T, gamma, dtype
        self.ctype = np.complex128 if dtype == np.float64 else np.clongdouble
        self.a, self.b, cf = tridiag_coeffs(N_x, N_t, T, gamma, dtype)
        self.z, self.sigma = cf["z"], cf["sigma"]
        self.workers = workers or len(os.sched_getaffinity(0))
        self.solver = solver or thomas_toeplitz

    def forward_stage(self, x):
        """ifft in time + rotation: returns (rho_plus, rho_minus), each (n, N_t)."""
        x = np.asarray(x, dtype=self.ctype).reshape(2, self.n, self.N_t)
        xh = sfft.ifft(x, axis=2, workers=self.workers)
        uz = xh[0] * np.conj(self.z)
        ip = (1j * self.sigma) * xh[1]
        return (uz + ip) / 2, (uz - ip) / 2

    def solve_stage(self, rp, rm):
        zp = np.zeros_like(rp)
        zm = np.zeros_like(rm)
        zp[1:-1] = self.solver(self.a, self.b, rp[1:-1])
        zm[1:-1] = np.conj(self.solver(self.a, self.b, np.conj(rm[1:-1])))
        return zp, zm

    def backward_stage(self, zp, zm):
        w = np.empty((2, self.n, self.N_t), dtype=self.ctype)
        w[0] = zp + zm
        w[1] = (-1j * self.sigma * self.z) * (zp - zm)
        return sfft.fft(w, axis=2, workers=self.workers)

    def stage_columns(self, ks, xh_cols):
        """The per-frequency stage (:445-540) on a SAMPLE of frequencies: ``xh_cols`` = columns ``ks`` of
        ifft_t(x), shape (2, n, len(ks)); returns the same columns of wh.  Frequencies decouple, so a handful of
        columns can be solved in 80-bit arithmetic (``dtype=np.longdouble``) at any BASELINE size."""
        ks = np.asarray(ks)
        xh = np.asarray(xh_cols, dtype=self.ctype)
        z, sg, a, b = self.z[ks], self.sigma[ks], self.a[ks], self.b[ks]
        uz = xh[0] * np.conj(z)
        ip = (1j * sg) * xh[1]
        rp, rm = (uz + ip) / 2, (uz - ip) / 2
        zp, zm = np.zeros_like(rp), np.zeros_like(rm)
        zp[1:-1] = thomas_toeplitz(a, b, rp[1:-1])
        zm[1:-1] = np.conj(thomas_toeplitz(a, b, np.conj(rm[1:-1])))
        return np.stack([zp + zm, (-1j * sg * z) * (zp - zm)])

    def apply_threaded(self, x):
        """Same operator with every stage on all host threads: scipy.fft with ``workers`` and the fused C
        stage of oracle/csrc/pc_solve.c (``oracle_pc_stage``).  The timed CPU baseline of bench.py; float64 only."""
        from . import csolve
        x = np.asarray(x, dtype=np.complex128).reshape(2, self.n, self.N_t)
        xh = sfft.ifft(x, axis=2, workers=self.workers)
        csolve.pc_stage_c(self.a, self.b, self.z, self.sigma, xh)
        return sfft.fft(xh, axis=2, workers=self.workers, overwrite_x=True).reshape(-1)

    def apply(self, x):
        rp, rm = self.forward_stage(x)
        zp, zm = self.solve_stage(rp, rm)
        return self.backward_stage(zp, zm).reshape(-1)
