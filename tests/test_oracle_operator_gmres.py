"""All-at-once operator, manufactured right-hand side and PETSc-semantic GMRES of the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle.gmres import gmres
from oracle.operator import AllAtOnce
from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_operator_structure_small():
    """A differs from the block-circulant P only in the wrapped time couplings and the two
    half-weighted coupling rows: check on explicit matrices."""
    N_x, N_t, gamma = 6, 7, 0.25
    op = AllAtOnce(N_x, N_t, 2.0, gamma, bug138=False)
    A = op.dense()
    n = N_x + 1
    P = ExplicitPC(N_x, N_t, 2.0, gamma).P.toarray()
    idx = np.array([(f * n + j) * N_t + i for f in range(2) for j in range(1, n - 1) for i in range(N_t)])
    Aint = A[np.ix_(idx, idx)]
    D = Aint - P
    # per spatial mode the difference has rank 4 -> overall rank <= 4 * (N_x - 1)
    assert np.linalg.matrix_rank(D, tol=1e-12) <= 4 * (N_x - 1)
    # boundary rows are identity
    b0 = np.array([(f * n + j) * N_t + i for f in range(2) for j in (0, n - 1) for i in range(N_t)])
    assert np.allclose(A[np.ix_(b0, b0)], np.eye(b0.size))
    # and interior rows ignore boundary columns
    assert np.abs(A[np.ix_(idx, b0)]).max() == 0


def test_bug138_only_touches_last_state_row():
    N_x, N_t, gamma = 5, 6, 0.01
    a = AllAtOnce(N_x, N_t, 2.0, gamma, bug138=True).dense()
    b = AllAtOnce(N_x, N_t, 2.0, gamma, bug138=False).dense()
    rows = np.nonzero(np.abs(a - b).max(axis=1))[0]
    n = N_x + 1
    expect = {(0 * n + j) * N_t + (N_t - 1) for j in range(1, n - 1)}
    assert set(rows.tolist()) == expect
    assert np.array_equal(AllAtOnce(N_x, N_t, 2.0, 1.0, True).dense(), AllAtOnce(N_x, N_t, 2.0, 1.0, False).dense())


@pytest.mark.parametrize("N_x,N_t,gamma", [(80, 81, 1.0), (32, 64, 1.0), (32, 128, 1e-2)])
def test_manufactured_rhs_converges_in_five_iterations(N_x, N_t, gamma):
    op = AllAtOnce(N_x, N_t, 2.0, gamma)
    pc = DiagFFTPCFast(N_x, N_t, 2.0, gamma)
    x, its, hist, reason = gmres(op.matvec, pc.apply, op.rhs(), rtol=1e-5)
    assert reason == "CONVERGED_RTOL" and its == 5
    rel_hist = np.array(hist) / hist[0]
    assert rel_hist[-1] < 1e-5 and np.all(rel_hist[1:5] > 1e-2)
    assert np.abs(x.imag).max() < 1e-8 * np.abs(x.real).max()


def test_solution_approximates_analytic_state():
    errs = []
    for N in (20, 40):
        op = AllAtOnce(N, N, 2.0, 1.0)
        x, *_ = gmres(op.matvec, DiagFFTPCFast(N, N).apply, op.rhs(), rtol=1e-9)
        ua, _ = op.analytic()
        errs.append(np.abs(x.reshape(2, N + 1, N)[0].real - ua).max())
    assert errs[1] < errs[0] < 0.2


def test_random_rhs_iteration_counts():
    N_x, N_t = 32, 64
    op = AllAtOnce(N_x, N_t)
    pc = DiagFFTPCFast(N_x, N_t)
    b = np.random.default_rng(0).standard_normal((2, N_x + 1, N_t))
    b[:, 0] = b[:, -1] = 0
    _, its, hist, reason = gmres(op.matvec, pc.apply, b.reshape(-1) + 0j, rtol=1e-7)
    assert reason == "CONVERGED_RTOL" and 55 <= its <= 63       # survey: 59


def test_restart_and_max_it():
    N_x, N_t = 16, 24
    op = AllAtOnce(N_x, N_t)
    pc = DiagFFTPCFast(N_x, N_t)
    b = np.random.default_rng(0).standard_normal(2 * (N_x + 1) * N_t) + 0j
    _, its_full, _, _ = gmres(op.matvec, pc.apply, b, rtol=1e-8, restart=300)
    x, its_r, _, reason = gmres(op.matvec, pc.apply, b, rtol=1e-8, restart=10, max_it=2000)
    assert reason == "CONVERGED_RTOL" and its_r >= its_full
    bb = b.reshape(2, N_x + 1, N_t).copy()
    assert np.linalg.norm(pc.apply(op.matvec(x) - b)) <= 1.01e-8 * np.linalg.norm(pc.apply(b))
    _, its_m, _, reason_m = gmres(op.matvec, pc.apply, b, rtol=1e-14, restart=300, max_it=7)
    assert its_m == 7 and reason_m == "DIVERGED_ITS"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gmres_*.npz"))))
def test_gmres_golden(path):
    g = np.load(path)
    N_x, N_t, gamma = int(g["N_x"]), int(g["N_t"]), float(g["gamma"])
    op = AllAtOnce(N_x, N_t, float(g["T"]), gamma)
    assert np.allclose(op.rhs(), g["b"], rtol=1e-13, atol=1e-13 * np.abs(g["b"]).max())
    assert np.allclose(op.matvec(g["xs"]), g["Axs"], rtol=1e-12, atol=1e-12 * np.abs(g["Axs"]).max())
    x, its, hist, _ = gmres(op.matvec, DiagFFTPCFast(N_x, N_t, float(g["T"]), gamma).apply, g["b"], rtol=1e-7)
    assert its == int(g["its"])
    assert np.allclose(np.array(hist)[:-1], g["hist"][:-1], rtol=1e-6)


def test_gmres_with_pc_reproduces_the_direct_lu_baseline():
    # upstream's own end-to-end check: the pc=True run (GMRES + DiagFFTPC, :567) against the pc=False direct
    # MUMPS solve (:573-577) of the same problem
    from oracle.gmres import gmres
    from oracle.operator import AllAtOnce
    from oracle.pc_fast import DiagFFTPCFast
    N_x, N_t = 16, 24
    op = AllAtOnce(N_x, N_t)
    direct = op.direct_solve()
    x, its, _, reason = gmres(op.matvec, DiagFFTPCFast(N_x, N_t).apply, op.rhs() + 0j, rtol=1e-12)
    assert reason == "CONVERGED_RTOL"
    assert np.linalg.norm(x - direct) / np.linalg.norm(direct) < 1e-9
    assert np.linalg.norm(op.matvec(direct) - op.rhs()) / np.linalg.norm(op.rhs()) < 1e-12
