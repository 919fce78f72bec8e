"""Regenerates tests/golden/upstream_apply_*.npz and upstream_operator_*.npz by EXECUTING the upstream classes.

``Code/Control_Wave_PC.py`` cannot run here as a script (no Firedrake / PETSc / MUMPS), but its two classes are plain
Python that only talks to Firedrake.  This script reads their source from the upstream checkout AT GENERATION TIME
(nothing of the reference is copied into this repository), executes it UNMODIFIED with ``fd`` bound to
``firedrake_standin`` (P1 mass / stiffness on the uniform interval, affine UFL forms, homogeneous Dirichlet rows, sparse
LU -- see that module's header for exactly what it supplies) and stores what the upstream code computes:

* ``upstream_apply_<N_x>_<N_t>_<gamma>.npz``: y = DiagFFTPC.apply(pc, x, y) (:491-553, after ``initialize`` :380-484)
  for a complex and a real random x -- the FFT direction and scaling, the eigen-decomposition and its ordering as
  numpy returns it, the S^-1 / S rotations, the conjugated 1/lambda_2, the Riesz round trip and the data layout all come
  from the executed upstream lines;
* ``upstream_operator_<N_x>_<N_t>_<gamma>.npz``: the all-at-once system A U = b that ``Build_f / Build_g /
  Build_Initial_Condition / Build_L`` (:48-179) define (``snes_type ksponly`` from U = 0 solves exactly this): b, A v for
  a random v, the direct solution A^-1 b (upstream's pc=False branch :573-577), and the GMRES history of
  KSPGMRES-as-restated (oracle/gmres.py; PETSc itself is not runnable) with the EXECUTED operator and the EXECUTED
  ``DiagFFTPC.apply`` as preconditioner.

tests/test_reference_executed_golden.py pins every oracle route (and tests/test_gpu_reference_executed.py the CUDA
path) to these.  N_t = 16 and 64 are included on purpose: upstream divides by lambda_2 = 1 + e^{4 pi i k / N_t}, which
is zero at k = N_t/4 in exact arithmetic and ~1e-16 as numpy evaluates it -- the executed code runs through (and GMRES
still takes 5 iterations), and the routes here, which avoid that division, agree with it to 1e-14.

Run from the repo root, where /root/reference exists:  python tests/golden/make_reference_executed_golden.py
"""
import math
import os
import sys
import time

import numpy as np
from scipy.fft import fft, ifft

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import firedrake_standin as fd  # noqa: E402

REF = os.environ.get("PARADIAG_REFERENCE", "/root/reference/Code/Control_Wave_PC.py")
CASES = [(16, 13, 1.0), (20, 81, 1.0), (33, 21, 1e-2), (24, 50, 1e-4), (12, 16, 1.0), (20, 64, 1e-2), (80, 81, 1.0)]


def upstream_segments():
    """Line-index ranges [first, last) of the two class definitions and of the set-up lines between them."""
    with open(REF) as fh:
        lines = fh.read().splitlines()
    find = lambda pred, start=0: next(i for i in range(start, len(lines)) if pred(lines[i]))
    c1 = find(lambda l: l.startswith("class Optimal_Control_Wave_Equation"))
    c1_end = find(lambda l: l.startswith("# the control test problem"), c1)
    setup = find(lambda l: l.startswith("# setup variables that we wish to use in the pc class"), c1_end)
    c2 = find(lambda l: l.startswith("class DiagFFTPC"), setup)
    c2_end = find(lambda l: l.startswith("if pc:"), c2)
    return lines, (c1, c1_end), (setup, c2), (c2, c2_end)


def run_case(N_x, N_t, gamma, T=2.0):
    lines, seg_problem, seg_setup, seg_pc = upstream_segments()
    ns = {"fd": fd, "np": np, "math": math, "time": time, "fft": fft, "ifft": ifft,
          "pc": True, "complex": True, "T": T, "N_t": N_t, "N_x": N_x, "gamma": gamma, "dim": 1,
          "__name__": "upstream_executed"}
    exec(compile("\n".join(lines[seg_problem[0]:seg_problem[1]]), REF, "exec"), ns)      # class Optimal_Control_...
    ns["equ"] = ns["Optimal_Control_Wave_Equation"](N_x, T, N_t, gamma, dim=1)           # :343
    setup_src = [l for l in lines[seg_setup[0]:seg_setup[1]] if not l.lstrip().startswith("#")]
    exec(compile("\n".join(setup_src), REF, "exec"), ns)                                  # V, W, bigv, vu, vp, dt, bcs, pc
    exec(compile("\n".join(lines[seg_pc[0]:seg_pc[1]]), REF, "exec"), ns)                 # class DiagFFTPC
    equ = ns["equ"]
    n = N_x + 1
    size = 2 * n * N_t
    rng = np.random.default_rng(N_x * 1000 + N_t)

    # ---- the preconditioner: initialize once, apply to a complex and to a real vector
    pcobj = ns["DiagFFTPC"]()
    pcobj.initialize(None)
    pcobj.update(None)

    def upstream_apply(x):
        xv, yv = fd.HostVec(x), fd.HostVec(np.zeros(size))
        pcobj.apply(None, xv, yv)
        return yv.array.copy()

    x = rng.standard_normal(size) + 1j * rng.standard_normal(size)
    x_real = rng.standard_normal(size)
    y, y_real = upstream_apply(x), upstream_apply(x_real + 0j)
    try:
        pcobj.applyTranspose(None, None, None)
        transpose = "implemented"
    except NotImplementedError:
        transpose = "NotImplementedError"
    tag = f"{N_x}_{N_t}_{gamma:g}"
    np.savez_compressed(os.path.join(HERE, f"upstream_apply_{tag}.npz"), N_x=N_x, N_t=N_t, T=T, gamma=gamma,
                        x=x, y=y, x_real=x_real, y_real=y_real.real, y_real_imag_max=np.abs(y_real.imag).max(),
                        apply_transpose=transpose, Lambda_1=pcobj.Lambda_1, Lambda_2=pcobj.Lambda_2)

    # ---- the all-at-once operator and right-hand side
    equ.Build_f()
    equ.Build_g()
    equ.Build_Initial_Condition()
    equ.Build_L()
    A, b = fd.NonlinearVariationalProblem(equ.L, equ.U, bcs=equ.bcs).affine_system()
    v = rng.standard_normal(size) + 1j * rng.standard_normal(size)
    import scipy.sparse.linalg as spla
    direct = spla.splu(A.tocsc()).solve(b)
    from oracle.gmres import gmres                                   # KSPGMRES as restated; operator and PC executed
    xg, its, hist, reason = gmres(lambda z: A @ z, upstream_apply, b, rtol=1e-7)
    np.savez_compressed(os.path.join(HERE, f"upstream_operator_{tag}.npz"), N_x=N_x, N_t=N_t, T=T, gamma=gamma,
                        b=b.real, b_imag_max=np.abs(b.imag).max(), v=v, Av=A @ v, direct=direct.real,
                        direct_imag_max=np.abs(direct.imag).max(), gmres_err_vs_direct=np.linalg.norm(xg - direct)
                        / np.linalg.norm(direct), gmres_its=its, gmres_hist=np.array(hist), gmres_reason=reason,
                        f=equ.f.data.real, g=equ.g.data.real, u_0=equ.u_0.data.real, u_1=equ.u_1.data.real)
    return its, reason, float(np.linalg.norm(xg - direct) / np.linalg.norm(direct))


def main():
    for case in CASES:
        its, reason, err = run_case(*case)
        print(f"N_x={case[0]} N_t={case[1]} gamma={case[2]:g}: GMRES {its} its ({reason}), "
              f"|x_gmres - x_direct| / |x_direct| = {err:.2e}")
    print("executed-upstream fixtures written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
