"""The CUDA path against EXECUTED upstream code -- GPU box only.

tests/golden/upstream_apply_*.npz / upstream_operator_*.npz hold what the two classes of ``Code/Control_Wave_PC.py``
computed when they were executed unmodified against tests/golden/firedrake_standin.py (see
tests/golden/make_reference_executed_golden.py; the CPU suite pins the oracle to the same files to 1e-11 ... 1e-14).
Tolerances: the apply as in the other golden tests (1e-10: two fp64 routes through systems of condition ~12/h^2);
right-hand side and matvec 1e-12 (rounding only); GMRES within +-1 iteration of the executed iteration (north star).
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from optimal_control_paradiag_b200 import DiagFFTPC, ParaDiagHandle, petsc_shim  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
APPLY = sorted(glob.glob(os.path.join(GOLDEN, "upstream_apply_*.npz")))
OPER = sorted(glob.glob(os.path.join(GOLDEN, "upstream_operator_*.npz")))


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("path", APPLY)
def test_cuda_apply_reproduces_the_executed_upstream_apply(path):
    g = np.load(path)
    N_x, N_t = int(g["N_x"]), int(g["N_t"])
    with ParaDiagHandle(N_x, N_t, T=float(g["T"]), gamma=float(g["gamma"])) as h:
        y = h.pc_apply_host(g["x"])                                   # DiagFFTPC.apply, :491-553
        assert rel(y, g["y"]) < 1e-10
        assert np.abs(y.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
        yr = h.pc_apply(torch.tensor(g["x_real"] + 0j, device=DEV)).cpu().numpy()
        assert rel(yr.real, g["y_real"]) < 1e-10
        if h.real_path_supported:                                     # float64 vectors, half spectrum
            y64 = h.pc_apply_real(torch.tensor(g["x_real"], device=DEV)).cpu().numpy()
            assert rel(y64, g["y_real"]) < 1e-10


def test_cuda_apply_through_the_pc_class_reproduces_the_executed_upstream_apply():
    g = np.load([p for p in APPLY if "80_81_1" in p][0])              # the upstream script's own constants
    DiagFFTPC.configure(N_x=int(g["N_x"]), N_t=int(g["N_t"]), T=float(g["T"]), gamma=float(g["gamma"]))
    try:
        pc = petsc_shim.PC()
        pc.setPythonContext(DiagFFTPC())
        pc.setUp()
        xv, yv = petsc_shim.Vec(g["x"].copy()), petsc_shim.Vec.zeros(g["x"].size)
        pc.apply(xv, yv)
        assert rel(yv.getArray(), g["y"]) < 1e-10
        pc.destroy()
    finally:
        DiagFFTPC._defaults = {}


@pytest.mark.parametrize("path", OPER)
def test_cuda_operator_rhs_and_gmres_reproduce_the_executed_upstream_system(path):
    g = np.load(path)
    N_x, N_t = int(g["N_x"]), int(g["N_t"])
    with ParaDiagHandle(N_x, N_t, T=float(g["T"]), gamma=float(g["gamma"])) as h:
        b = h.build_rhs()                                             # Build_f / Build_g / Build_Initial_Condition
        assert rel(b.cpu().numpy().real, g["b"]) < 1e-12
        Av = h.matvec(torch.tensor(g["v"], device=DEV)).cpu().numpy()  # Build_L (matrix-free Jacobian action)
        assert rel(Av, g["Av"]) < 1e-12
        x, its, hist, reason = h.gmres(b, rtol=1e-7)                  # the solve of :347-359 / :567
        assert reason == "CONVERGED_RTOL"
        assert abs(its - int(g["gmres_its"])) <= 1
        m = min(its, int(g["gmres_its"]))
        assert np.allclose(np.asarray(hist)[:m], g["gmres_hist"][:m], rtol=1e-5)
        assert rel(x.cpu().numpy().real, g["direct"]) < 1e-7          # = the pc=False direct solve, :573-577
