"""1-D uniform P1 (CG1) mass and stiffness matrices on the unit interval.

Oracle code (test infrastructure).  Restates what Firedrake assembles for
``fd.UnitIntervalMesh(N_x)`` with ``CG1`` (Control_Wave_PC.py:17, :33, :42):
``n = N_x + 1`` nodes at ``x_j = j / N_x``, element length ``h = 1 / N_x``.

    M = h/6 * tridiag(1, 4, 1)   (boundary diagonal entries h/3)
    K = 1/h * tridiag(-1, 2, -1) (boundary diagonal entries 1/h)

Homogeneous Dirichlet conditions on both ends (:44-45) remove rows/columns 0
and N_x; the interior blocks are exactly Toeplitz.
"""
import numpy as np
import scipy.sparse as sp


def mass_full(N_x, dtype=np.float64):
    n = N_x + 1
    h = dtype(1) / dtype(N_x)
    d = np.full(n, 4, dtype=dtype)
    d[0] = d[-1] = 2
    o = np.ones(n - 1, dtype=dtype)
    return sp.diags([o, d, o], [-1, 0, 1], format="csr") * (h / 6)


def stiff_full(N_x, dtype=np.float64):
    n = N_x + 1
    h = dtype(1) / dtype(N_x)
    d = np.full(n, 2, dtype=dtype)
    d[0] = d[-1] = 1
    o = -np.ones(n - 1, dtype=dtype)
    return sp.diags([o, d, o], [-1, 0, 1], format="csr") * (1 / h)


def interior(A):
    """Drop the two Dirichlet rows / columns."""
    return A[1:-1, 1:-1].tocsr()


def stencil(N_x):
    """Interior Toeplitz entries (m_off, m_diag, k_off, k_diag)."""
    h = 1.0 / N_x
    return h / 6.0, 2.0 * h / 3.0, -1.0 / h, 2.0 / h
