"""The alpha extension of the oracle (oracle/pc_alpha.py): three routes agree, alpha = 1 is the upstream
operator, and the GMRES iteration counts of SURVEY H1 are reproduced.  Parity unpinned by construction:
the reference has no alpha."""
import glob
import os

import numpy as np
import pytest

from oracle.gmres import gmres
from oracle.operator import AllAtOnce
from oracle.pc_alpha import BlockAlphaPC, DiagFFTPCAlpha, ExplicitAlphaPC, alpha_circulant
from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def rand_x(N_x, N_t, seed=0):
    rng = np.random.default_rng(seed)
    size = 2 * (N_x + 1) * N_t
    return rng.standard_normal(size) + 1j * rng.standard_normal(size)


@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 13, 1.0), (12, 16, 1.0), (20, 32, 1e-2)])
@pytest.mark.parametrize("alpha", [0.5, 1e-1, 1e-2, 1e-4, 1e-6])
def test_three_alpha_routes_agree(N_x, N_t, gamma, alpha):
    x = rand_x(N_x, N_t)
    e = ExplicitAlphaPC(N_x, N_t, 2.0, gamma, alpha).apply(x)
    tol = 1e-11 if alpha >= 1e-4 else 2e-10            # P_alpha's adjoint block carries 1/alpha entries
    assert rel(BlockAlphaPC(N_x, N_t, 2.0, gamma, alpha).apply(x), e) < tol
    assert rel(DiagFFTPCAlpha(N_x, N_t, 2.0, gamma, alpha).apply(x), e) < tol


def test_alpha_one_is_the_upstream_operator():
    N_x, N_t = 16, 13
    x = rand_x(N_x, N_t)
    e1 = ExplicitPC(N_x, N_t).apply(x)
    assert rel(ExplicitAlphaPC(N_x, N_t, alpha=1.0).apply(x), e1) < 1e-13
    assert rel(DiagFFTPCAlpha(N_x, N_t, alpha=1.0).apply(x), e1) < 1e-12       # N_t odd: lambda_2 != 0
    # continuity: alpha -> 1 approaches the reference preconditioner
    assert rel(DiagFFTPCAlpha(N_x, N_t, alpha=1 - 1e-9).apply(x), DiagFFTPCFast(N_x, N_t).apply(x)) < 1e-6


def test_alpha_circulant_is_diagonalised_by_gamma_and_the_dft():
    # Gamma C_alpha Gamma^-1 is the plain circulant of the scaled stencil: eigenvalues (1 - a z)^2
    N, alpha = 12, 1e-3
    a = alpha ** (1.0 / N)
    C = alpha_circulant((1.0, -2.0, 1.0), N, alpha)
    G = np.diag(a ** np.arange(N))
    Ct = G @ C @ np.linalg.inv(G)
    lam = np.fft.fft(Ct[:, 0])                              # eigenvalues of a circulant = DFT of its first column
    z = np.exp(-2j * np.pi * np.arange(N) / N)
    assert np.abs(lam - (1 - a * z) ** 2).max() < 1e-13
    assert np.abs(Ct - np.roll(np.roll(Ct, 1, 0), 1, 1)).max() < 1e-13   # Ct is circulant


def test_alpha_sweep_iteration_counts():
    # SURVEY H1 (N_x = 20, N_t = 32, gamma = 1, rtol 1e-7): 5, 12, 18, 25 iterations; for alpha <= 1e-3 the
    # preconditioned residual still "converges" while the true residual does not
    N_x, N_t = 20, 32
    op = AllAtOnce(N_x, N_t)
    b = op.rhs() + 0j
    counts = {}
    for alpha in (1.0, 0.5, 1e-1, 1e-2, 1e-4):
        pc = DiagFFTPCFast(N_x, N_t) if alpha == 1.0 else DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha)
        x, its, _, reason = gmres(op.matvec, pc.apply, b, rtol=1e-7)
        counts[alpha] = (its, np.linalg.norm(op.matvec(x) - b) / np.linalg.norm(b))
    assert [counts[a][0] for a in (1.0, 0.5, 1e-1, 1e-2)] == [5, 12, 18, 25]
    assert counts[1.0][1] < 1e-10 and counts[1e-4][1] > 1.0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "alpha_*.npz"))))
def test_decoupled_route_matches_alpha_golden(path):
    g = np.load(path)
    pc = DiagFFTPCAlpha(int(g["N_x"]), int(g["N_t"]), float(g["T"]), float(g["gamma"]), float(g["alpha"]))
    assert rel(pc.apply(g["x"]), g["y"]) < 1e-10


def test_alpha_fixtures_exist():
    assert len(glob.glob(os.path.join(GOLDEN, "alpha_*.npz"))) >= 3
