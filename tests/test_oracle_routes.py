"""The oracle's three independent routes to P^-1 x agree, and match the committed fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast, thomas_toeplitz
from oracle.pc_ref_route import DiagFFTPCRefRoute

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [(16, 13, 1.0), (20, 81, 1.0), (16, 13, 0.01), (16, 16, 1.0), (24, 64, 1e-4), (12, 24, 1.0)]


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def rand_x(N_x, N_t, seed=0):
    rng = np.random.default_rng(seed)
    size = 2 * (N_x + 1) * N_t
    return rng.standard_normal(size) + 1j * rng.standard_normal(size)


@pytest.mark.parametrize("N_x,N_t,gamma", CASES)
def test_three_routes_agree(N_x, N_t, gamma):
    x = rand_x(N_x, N_t)
    e = ExplicitPC(N_x, N_t, 2.0, gamma).apply(x)
    r = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(x)
    f = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x)
    assert rel(r, e) < 2e-12
    assert rel(f, e) < 2e-12
    assert rel(f, r) < 2e-12


def test_riesz_roundtrip_is_algebraically_void():
    x = rand_x(20, 27)
    a = DiagFFTPCRefRoute(20, 27, riesz_roundtrip=True).apply(x)
    b = DiagFFTPCRefRoute(20, 27, riesz_roundtrip=False).apply(x)
    assert rel(a, b) < 1e-12


@pytest.mark.parametrize("N_x,N_t,gamma", CASES[:3])
def test_boundary_rows_are_exactly_zero(N_x, N_t, gamma):
    y = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(rand_x(N_x, N_t)).reshape(2, N_x + 1, N_t)
    assert np.all(y[:, 0, :] == 0) and np.all(y[:, -1, :] == 0)
    yr = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(rand_x(N_x, N_t)).reshape(2, N_x + 1, N_t)
    assert np.abs(yr[:, [0, -1], :]).max() == 0


def test_real_input_gives_real_output():
    N_x, N_t = 24, 32
    x = np.random.default_rng(1).standard_normal(2 * (N_x + 1) * N_t)
    y = DiagFFTPCFast(N_x, N_t).apply(x)
    assert np.abs(y.imag).max() < 1e-12 * np.abs(y.real).max()


def test_long_double_route():
    N_x, N_t = 64, 64
    x = rand_x(N_x, N_t)
    f = DiagFFTPCFast(N_x, N_t).apply(x)
    l = DiagFFTPCFast(N_x, N_t, dtype=np.longdouble).apply(x)
    assert l.dtype == np.clongdouble
    assert rel(f, l.astype(complex)) < 1e-11


def test_c_helper_matches_numpy_thomas():
    from oracle import csolve
    N_x, N_t = 96, 40
    x = rand_x(N_x, N_t)
    a = DiagFFTPCFast(N_x, N_t).apply(x)
    b = DiagFFTPCFast(N_x, N_t, solver=csolve.thomas_toeplitz_c).apply(x)
    assert rel(b, a) < 1e-12


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "pc_apply_*.npz"))))
def test_fast_route_matches_golden(path):
    g = np.load(path)
    pc = DiagFFTPCFast(int(g["N_x"]), int(g["N_t"]), float(g["T"]), float(g["gamma"]))
    assert rel(pc.apply(g["x"]), g["y"]) < 1e-11
    assert rel(pc.apply(g["x_real"]), g["y_real"]) < 1e-11


def test_golden_fixtures_exist():
    assert len(glob.glob(os.path.join(GOLDEN, "pc_apply_*.npz"))) >= 4
    assert len(glob.glob(os.path.join(GOLDEN, "gmres_*.npz"))) >= 2
