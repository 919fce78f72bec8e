"""Line-by-line CPU restatement of ``DiagFFTPC`` (Control_Wave_PC.py:376-558).

Oracle code (test infrastructure).  Vectors use the PETSc layout of the
reference: ``x.reshape(2, n, N_t)`` -- field (u, p) slowest, node, then time
fastest (``dat.data`` has shape ``(n, N_t)`` and the FFTs run along
``axis=1``, :496-501).

Firedrake objects are replaced by their algebraic meaning on the uniform 1-D
P1 mesh:

* ``xf.riesz_representation()`` (:506)  -> solve with the full mass matrix
  (default L2 Riesz map, no boundary conditions),
* assembling the form ``L`` (:445-457)  -> ``M_full (S^-1 (x) I) f``,
* ``solv_w.solve()`` with ``bcs`` (:482-484, :512) -> per frequency k two
  interior solves ``(Sigma_i(k) M + dt^2/2 K) w = rhs`` with boundary rows
  replaced by identity and right-hand side 0 (homogeneous Dirichlet, :44-45),
* the two ``interpolate`` loops (:516-540) -> pointwise 2x2 multiply by S_k
  and division by lambda_2 / conj(lambda_2).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.fft import fft, ifft

from . import eigs, fem1d


class DiagFFTPCRefRoute:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, riesz_roundtrip=True):
        self.N_x, self.N_t = N_x, N_t
        self.n = N_x + 1
        self.dt = T / N_t                                   # :24, :367
        self.gamma = gamma
        self.riesz_roundtrip = riesz_roundtrip
        self.initialize()

    # -- :380-484 ---------------------------------------------------------
    def initialize(self):
        N_t, dt, gamma = self.N_t, self.dt, self.gamma
        self.Lambda_1, self.Lambda_2 = eigs.lambdas(N_t)   # :387-388
        self.Sigma, self.S, self.SI = eigs.eig_numpy(N_t, dt, gamma)  # :415-436
        self.M = fem1d.mass_full(self.N_x).tocsc()
        self.K = fem1d.stiff_full(self.N_x).tocsc()
        self.Mi = fem1d.interior(self.M).tocsc()
        self.Ki = fem1d.interior(self.K).tocsc()
        self._Mlu = spla.splu(self.M.astype(complex))
        # LHS form D :460-473 with bcs :482 -> one interior LU per (k, field)
        self._lu = [[spla.splu((self.Sigma[k, f] * self.Mi
                                + dt ** 2 / 2 * self.Ki).tocsc())
                     for f in range(2)] for k in range(N_t)]

    # -- :491-553 ---------------------------------------------------------
    def apply(self, x):
        n, N_t = self.n, self.N_t
        x = np.asarray(x, dtype=complex).reshape(2, n, N_t)
        u_array, p_array = x[0], x[1]                        # :496-497
        vu1 = ifft(u_array, axis=1)                          # :500
        vp1 = ifft(p_array, axis=1)                          # :501
        if self.riesz_roundtrip:
            # f = M^-1 xhat (:506), RHS assembly multiplies by M again (:449-457)
            fu = self._Mlu.solve(vu1)
            fp = self._Mlu.solve(vp1)
        else:
            fu, fp = vu1, vp1
        w = np.zeros((2, n, N_t), dtype=complex)
        for k in range(N_t):
            SI = self.SI[k]
            r1 = SI[0, 0] * fu[:, k] + SI[0, 1] * fp[:, k]   # :449, :456
            r2 = SI[1, 0] * fu[:, k] + SI[1, 1] * fp[:, k]   # :450, :457
            if self.riesz_roundtrip:
                r1 = self.M @ r1
                r2 = self.M @ r2
            # Dirichlet rows: identity, rhs 0  (:482, bcs :44-45)
            w[0, 1:-1, k] = self._lu[k][0].solve(r1[1:-1])   # :512
            w[1, 1:-1, k] = self._lu[k][1].solve(r2[1:-1])
        y = np.zeros((2, n, N_t), dtype=complex)
        with np.errstate(all="ignore"):
            for k in range(N_t):
                S = self.S[k]
                yu = S[0, 0] * w[0, :, k] + S[0, 1] * w[1, :, k]  # :526
                yp = S[1, 0] * w[0, :, k] + S[1, 1] * w[1, :, k]  # :527
                y[0, :, k] = yu / self.Lambda_2[k]                # :537
                y[1, :, k] = yp / np.conj(self.Lambda_2[k])       # :538
        y[0] = fft(y[0], axis=1)                              # :547
        y[1] = fft(y[1], axis=1)                              # :548
        return y.reshape(-1)

    def applyTranspose(self, x):                              # :557-558
        raise NotImplementedError
