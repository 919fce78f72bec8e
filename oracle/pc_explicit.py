"""Explicit sparse block-circulant preconditioner matrix P and its LU.

Oracle code (test infrastructure); semantic ground truth at small sizes.

``DiagFFTPC`` (Control_Wave_PC.py:376-558) is, algebraically, the inverse of the
all-at-once matrix of ``Build_L`` (:86-179, ``pc=True`` branches) with the
lower-Toeplitz time stencils B1 = (1,-2,1), B2 = (1,0,1) replaced by the
circulants C1, C2 (first columns (1,-2,1,0..), (1,0,1,0..); mat_test.ipynb
cells 5-9) and the first/last-row half weights of the coupling replaced by 1:

    P = [ C1 (x) M + dt^2/2 C2 (x) K        -c I (x) M               ]
        [ c I (x) M                 C1^T (x) M + dt^2/2 C2^T (x) K   ]

with c = dt^2/sqrt(gamma) and homogeneous Dirichlet rows (identity, output 0).
Ordering is the PETSc layout: index(field, node j, time i) = (f*n + j)*N_t + i.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem1d


def circulants(N_t):
    c1 = np.zeros(N_t)
    c2 = np.zeros(N_t)
    for off, v in ((0, 1.0), (1, -2.0), (2, 1.0)):
        c1[off % N_t] += v
    for off, v in ((0, 1.0), (2, 1.0)):
        c2[off % N_t] += v
    return sp.csr_matrix(sla.circulant(c1)), sp.csr_matrix(sla.circulant(c2))


class ExplicitPC:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0):
        self.N_x, self.N_t, self.n = N_x, N_t, N_x + 1
        dt = T / N_t
        c = dt ** 2 / np.sqrt(gamma)
        m = N_x - 1
        Mi = fem1d.interior(fem1d.mass_full(N_x))
        Ki = fem1d.interior(fem1d.stiff_full(N_x))
        C1, C2 = circulants(N_t)
        It = sp.identity(N_t, format="csr")
        # node-major, time-fastest ordering -> kron(space, time)
        Puu = sp.kron(Mi, C1) + dt ** 2 / 2 * sp.kron(Ki, C2)
        Ppp = sp.kron(Mi, C1.T) + dt ** 2 / 2 * sp.kron(Ki, C2.T)
        Pup = -c * sp.kron(Mi, It)
        Ppu = c * sp.kron(Mi, It)
        self.P = sp.bmat([[Puu, Pup], [Ppu, Ppp]], format="csc")
        self._lu = spla.splu(self.P)
        self.m = m

    def apply(self, x):
        n, N_t, m = self.n, self.N_t, self.m
        x = np.asarray(x, dtype=complex).reshape(2, n, N_t)
        rhs = x[:, 1:-1, :].reshape(-1)
        sol = self._lu.solve(rhs.real) + 1j * self._lu.solve(rhs.imag)
        y = np.zeros((2, n, N_t), dtype=complex)
        y[:, 1:-1, :] = sol.reshape(2, m, N_t)
        return y.reshape(-1)
