"""The oracle against EXECUTED upstream code for the whole path.

tests/golden/upstream_apply_*.npz and upstream_operator_*.npz are produced by
tests/golden/make_reference_executed_golden.py: the two classes of ``Code/Control_Wave_PC.py`` executed unmodified
(``Optimal_Control_Wave_Equation`` :13-179 with ``Build_f/g/Initial_Condition/L``; ``DiagFFTPC.initialize`` / ``apply``
:380-553) with ``fd`` bound to tests/golden/firedrake_standin.py (P1 matrices on the uniform interval, affine UFL forms,
Dirichlet rows, sparse LU -- nothing that knows about the preconditioner).  Here every route of the oracle is pinned
to what those executed lines computed: the apply (complex and real input), the all-at-once operator and right-hand
side including the 1/2-weight rows and the :138 quirk, the direct baseline, and the GMRES iteration."""
import glob
import os

import numpy as np
import pytest

from oracle import csolve, eigs
from oracle.gmres import gmres
from oracle.operator import AllAtOnce
from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast
from oracle.pc_ref_route import DiagFFTPCRefRoute

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
APPLY = sorted(glob.glob(os.path.join(GOLDEN, "upstream_apply_*.npz")))
OPER = sorted(glob.glob(os.path.join(GOLDEN, "upstream_operator_*.npz")))


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def cfg(g):
    return int(g["N_x"]), int(g["N_t"]), float(g["T"]), float(g["gamma"])


def test_executed_upstream_fixtures_exist():
    assert len(APPLY) >= 5 and len(OPER) >= 5
    assert any("80_81_1" in p for p in APPLY)                 # the upstream script's own constants, :335-339


@pytest.mark.parametrize("path", APPLY)
def test_every_oracle_route_reproduces_the_executed_upstream_apply(path):
    g = np.load(path)
    N_x, N_t, T, gamma = cfg(g)
    x, y = g["x"], g["y"]
    routes = {
        "closed form, numpy Thomas": DiagFFTPCFast(N_x, N_t, T, gamma).apply,
        "closed form, C Thomas": DiagFFTPCFast(N_x, N_t, T, gamma, solver=csolve.thomas_toeplitz_c).apply,
        "closed form, threaded (the timed CPU baseline)": DiagFFTPCFast(N_x, N_t, T, gamma, workers=2).apply_threaded,
        "line-by-line restatement of :380-553": DiagFFTPCRefRoute(N_x, N_t, T, gamma).apply,
    }
    if 2 * (N_x + 1) * N_t <= 4000:
        routes["explicit block-circulant matrix, SuperLU"] = ExplicitPC(N_x, N_t, T, gamma).apply
    for name, apply in routes.items():
        assert rel(apply(x), y) < 1e-11, (name, rel(apply(x), y))
        assert rel(apply(g["x_real"] + 0j).real, g["y_real"]) < 1e-11, name
    # what the executed upstream lines themselves show
    Y = y.reshape(2, N_x + 1, N_t)
    assert np.abs(Y[:, [0, -1], :]).max() == 0.0                       # Dirichlet rows of the result
    assert float(g["y_real_imag_max"]) <= 1e-11 * np.abs(g["y_real"]).max()   # real in, real out (up to rounding)
    assert str(g["apply_transpose"]) == "NotImplementedError"          # :557-558
    l1, l2 = eigs.lambdas(N_t)
    assert np.array_equal(l1, g["Lambda_1"]) and np.array_equal(l2, g["Lambda_2"])


@pytest.mark.parametrize("path", OPER)
def test_operator_and_right_hand_side_reproduce_the_executed_upstream_forms(path):
    g = np.load(path)
    N_x, N_t, T, gamma = cfg(g)
    op = AllAtOnce(N_x, N_t, T, gamma)
    assert float(g["b_imag_max"]) == 0.0
    assert rel(op.rhs(), g["b"]) < 1e-14                               # Build_f / Build_g / Build_Initial_Condition
    assert rel(op.matvec(g["v"]), g["Av"]) < 1e-14                     # Build_L incl. :117, :143 and the :138 quirk
    if gamma != 1.0:                                                   # without the quirk the operator is another one
        assert rel(AllAtOnce(N_x, N_t, T, gamma, bug138=False).matvec(g["v"]), g["Av"]) > 1e-6
    # the pc=False branch (:573-577) of the executed system against the oracle's
    assert rel(op.direct_solve(), g["direct"]) < 1e-9
    assert float(g["direct_imag_max"]) <= 1e-12 * np.abs(g["direct"]).max()


@pytest.mark.parametrize("path", OPER)
def test_gmres_on_the_executed_operator_with_the_executed_preconditioner(path):
    # KSP itself is not runnable; the fixture ran oracle.gmres on the EXECUTED operator with the EXECUTED apply.
    # The oracle's own operator and preconditioner must give the same iteration.
    g = np.load(path)
    N_x, N_t, T, gamma = cfg(g)
    op = AllAtOnce(N_x, N_t, T, gamma)
    x, its, hist, reason = gmres(op.matvec, DiagFFTPCFast(N_x, N_t, T, gamma).apply, op.rhs() + 0j, rtol=1e-7)
    assert reason == str(g["gmres_reason"]) == "CONVERGED_RTOL"
    assert its == int(g["gmres_its"])
    assert np.allclose(np.array(hist)[:-1], g["gmres_hist"][:-1], rtol=1e-8)
    assert float(g["gmres_err_vs_direct"]) < 1e-9 and rel(x, g["direct"]) < 1e-9
    if gamma == 1.0:
        assert its == 5                                                # the 5-step termination of the manufactured problem


def test_standin_matrices_are_the_p1_matrices_of_the_oracle():
    # the one thing the stand-in and the oracle both state independently
    import sys
    sys.path.insert(0, GOLDEN)
    import firedrake_standin as fd
    from oracle import fem1d
    mesh = fd.UnitIntervalMesh(12)
    assert np.allclose(mesh.M.toarray(), fem1d.mass_full(12).toarray() if hasattr(fem1d.mass_full(12), "toarray")
                       else fem1d.mass_full(12), rtol=1e-15, atol=0)
    assert np.allclose(mesh.K.toarray(), fem1d.stiff_full(12).toarray() if hasattr(fem1d.stiff_full(12), "toarray")
                       else fem1d.stiff_full(12), rtol=1e-15, atol=0)
    # exactness for P1: integral of a linear function times a hat function
    x = mesh.coords
    assert np.isclose(np.ones_like(x) @ (mesh.M @ x), 0.5) and np.isclose(x @ (mesh.K @ x), 1.0)


def test_executed_upstream_script_flow_for_the_accuracy_study():
    # tests/golden/upstream_accuracy.json: equ.solve(...) + equ.write(...) of the upstream script executed for
    # N = 5 ... 70 (both branches).  Non-binding numbers (write / plot.py are outside the hot path); pinned here:
    # the solve itself -- 5 GMRES iterations at every N, pc and LU branches agreeing, and the oracle's direct
    # solution having the norm the executed solve produced.
    import json
    tab = json.load(open(os.path.join(GOLDEN, "upstream_accuracy.json")))["table"]
    assert sorted(int(k) for k in tab) == list(range(5, 75, 5))
    for k, row in tab.items():
        N = int(k)
        assert row["gmres_its"] == 5 and row["gmres_reason"] == "CONVERGED_RTOL"
        assert abs(row["pc_u_norm"] - row["lu_u_norm"]) < 1e-9 * row["lu_u_norm"]
        if N < 25:       # write() reads nodal entry 25 (:281-282): the committed script cannot run these sizes
            assert str(row["pc"]).startswith("IndexError") and str(row["lu"]).startswith("IndexError")
        else:
            assert abs(row["pc"] - row["lu"]) < 1e-8 * row["lu"]
            assert row["lu"] > 5 * row["plot_py"]    # the committed script does not reproduce plot.py:5-18
    for N in (10, 25, 40):
        u = AllAtOnce(N, N, 2.0, 1.0).direct_solve().reshape(2, N + 1, N)[0]
        assert abs(np.linalg.norm(u) - tab[str(N)]["lu_u_norm"]) < 1e-9 * tab[str(N)]["lu_u_norm"]
