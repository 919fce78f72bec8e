"""Regenerates tests/golden/*.npz.

The upstream repository ships no golden vectors and cannot run here (no Firedrake / PETSc /
MUMPS), so these fixtures are produced by the oracle's LINE-BY-LINE restatement of
``DiagFFTPC`` (oracle/pc_ref_route.py: numpy eig/inv per frequency, shifted solves, 1/lambda_2,
scipy fft -- the same third-party routines upstream calls) and by the oracle GMRES on the
restated operator.  They pin (a) the oracle's other routes and (b) the CUDA path against
regressions; they are NOT outputs of the upstream code ("parity unpinned").

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.gmres import gmres  # noqa: E402
from oracle.operator import AllAtOnce  # noqa: E402
from oracle.pc_fast import DiagFFTPCFast  # noqa: E402
from oracle.pc_ref_route import DiagFFTPCRefRoute  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PC_CASES = [(16, 13, 1.0), (20, 81, 1.0), (16, 16, 1.0), (24, 64, 1e-4), (33, 20, 1e-2)]
GMRES_CASES = [(80, 81, 1.0), (32, 64, 1e-2), (24, 48, 1.0)]


def main():
    for (N_x, N_t, gamma) in PC_CASES:
        rng = np.random.default_rng(0)
        size = 2 * (N_x + 1) * N_t
        x = rng.standard_normal(size) + 1j * rng.standard_normal(size)
        y = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(x)
        xr = np.random.default_rng(1).standard_normal(size) + 0j
        yr = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(xr)
        np.savez_compressed(os.path.join(HERE, f"pc_apply_{N_x}_{N_t}_{gamma:g}.npz"),
                            N_x=N_x, N_t=N_t, T=2.0, gamma=gamma, x=x, y=y, x_real=xr, y_real=yr)
    for (N_x, N_t, gamma) in GMRES_CASES:
        op = AllAtOnce(N_x, N_t, 2.0, gamma)
        pc = DiagFFTPCFast(N_x, N_t, 2.0, gamma)
        b = op.rhs()
        sol, its, hist, reason = gmres(op.matvec, pc.apply, b, rtol=1e-7, restart=300, max_it=1000)
        rng = np.random.default_rng(0)
        br = rng.standard_normal((2, N_x + 1, N_t))
        br[:, 0] = br[:, -1] = 0
        _, its_r, hist_r, _ = gmres(op.matvec, pc.apply, br.reshape(-1) + 0j, rtol=1e-7, restart=300, max_it=1000)
        xs = np.random.default_rng(2).standard_normal(b.size) + 1j * np.random.default_rng(3).standard_normal(b.size)
        np.savez_compressed(os.path.join(HERE, f"gmres_{N_x}_{N_t}_{gamma:g}.npz"),
                            N_x=N_x, N_t=N_t, T=2.0, gamma=gamma, b=b, x=sol, its=its, hist=np.array(hist),
                            its_random=its_r, hist_random=np.array(hist_r), xs=xs, Axs=op.matvec(xs))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
