"""The oracle against outputs of EXECUTED upstream code.

tests/golden/refsetup_*.npz hold what Code/Control_Wave_PC.py:387-436 (the eigen-setup of
``DiagFFTPC.initialize``: Lambda_1, Lambda_2 and the per-frequency numpy ``eig`` / ``inv`` loop) computes when
those lines are run unmodified (tests/golden/make_reference_setup_golden.py).  They pin ``oracle/eigs.py`` bit
for bit, and -- injected into the line-by-line route -- every other route of the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import eigs
from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast
from oracle.pc_ref_route import DiagFFTPCRefRoute

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refsetup_*.npz")))


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _eq(a, b):
    # bit-for-bit where finite; the same non-finite pattern where upstream divides by lambda_2 = 0
    fin = np.isfinite(a) & np.isfinite(b)
    return np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(a[fin], b[fin])


def test_fixtures_exist():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES)
def test_oracle_eigen_setup_equals_executed_upstream_code(path):
    g = np.load(path)
    N_t, T, gamma = int(g["N_t"]), float(g["T"]), float(g["gamma"])
    l1, l2 = eigs.lambdas(N_t)
    assert _eq(l1, g["Lambda_1"]) and _eq(l2, g["Lambda_2"])
    Sig, S, SI = eigs.eig_numpy(N_t, T / N_t, gamma)
    assert _eq(Sig[:, 0], g["Sigma_1"]) and _eq(Sig[:, 1], g["Sigma_2"])
    for (i, j), name in (((0, 0), "S11"), ((0, 1), "S12"), ((1, 0), "S21"), ((1, 1), "S22")):
        assert _eq(S[:, i, j], g[name]), name
        assert _eq(SI[:, i, j], g["SI" + name[1:]]), name


@pytest.mark.parametrize("path", [p for p in FIXTURES if "refsetup_16_" not in p])
def test_routes_with_the_upstream_arrays_injected(path):
    # the line-by-line route driven by the arrays upstream computed, against the routes that never see them
    g = np.load(path)
    N_t, T, gamma = int(g["N_t"]), float(g["T"]), float(g["gamma"])
    N_x = 20
    pc = DiagFFTPCRefRoute(N_x, N_t, T, gamma)
    pc.Lambda_1, pc.Lambda_2 = g["Lambda_1"], g["Lambda_2"]
    pc.S = np.stack([np.stack([g["S11"], g["S12"]], -1), np.stack([g["S21"], g["S22"]], -1)], -2)
    pc.SI = np.stack([np.stack([g["SI11"], g["SI12"]], -1), np.stack([g["SI21"], g["SI22"]], -1)], -2)
    assert np.array_equal(pc.Sigma, np.stack([g["Sigma_1"], g["Sigma_2"]], -1))     # the LU factors in use
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2 * (N_x + 1) * N_t) + 1j * rng.standard_normal(2 * (N_x + 1) * N_t)
    y = pc.apply(x)
    assert rel(DiagFFTPCFast(N_x, N_t, T, gamma).apply(x), y) < 2e-12
    assert rel(ExplicitPC(N_x, N_t, T, gamma).apply(x), y) < 2e-12


def test_closed_forms_against_the_upstream_eigenvalues():
    # what the CUDA kernels regenerate (make_coef): Sigma_+- = Re(l1/l2) +- i c / |l2|, against upstream's eig
    g = np.load(os.path.join(GOLDEN, "refsetup_81_1.npz"))
    N_t, T, gamma = int(g["N_t"]), float(g["T"]), float(g["gamma"])
    cf = eigs.closed_form(N_t, T / N_t, gamma)
    l1, l2 = g["Lambda_1"], g["Lambda_2"]
    plus = np.real(l1 / l2) + 1j * cf["c"] / np.abs(l2)
    for k in range(N_t):
        got = sorted([g["Sigma_1"][k], g["Sigma_2"][k]], key=lambda v: v.imag)
        assert abs(got[1] - plus[k]) < 1e-12 * max(1, abs(plus[k])) and abs(got[0] - np.conj(plus[k])) < 1e-12 * max(1, abs(plus[k]))
