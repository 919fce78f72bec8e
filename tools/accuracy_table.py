"""The upstream accuracy study (Control_Wave_PC.py:583-631, results hard-coded in plot.py:5-18) on the B200 path:
max-over-time nodal 2-norm error of the computed state against the analytic one, N_x = N_t = N = 5 ... 70, T = 2,
gamma = 1, once with GMRES + DiagFFTPC and once with the direct-LU baseline (pc=False branch).

NON-BINDING: SURVEY section 4 found upstream's published numbers not reproducible from the committed script by any
restatement (its ``write()`` pairs u_sol[i-2] with t = i tau); the table reports the deviation, it asserts nothing.
GPU box only:  python tools/accuracy_table.py > profiles/r02_accuracy_table.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import Optimal_Control_Wave_Equation, default_parameters  # noqa: E402

# plot.py:5-18 (numerical results published by upstream; data, not code)
UPSTREAM = {5: 9.0425e-01, 10: 2.1949e-01, 15: 2.0741e-01, 20: 7.4347e-02, 25: 8.4479e-02, 30: 4.0363e-02,
            35: 4.7018e-02, 40: 2.6207e-02, 45: 3.0604e-02, 50: 1.8751e-02, 55: 2.1843e-02, 60: 1.4264e-02,
            65: 1.6556e-02, 70: 1.1320e-02}


def main():
    print(f"{'N':>4} {'GMRES+PC its':>12} {'err GMRES+PC':>14} {'err direct LU':>14} {'|diff|':>10} {'upstream':>12} {'ours/upstream':>14}")
    for N in range(5, 75, 5):
        equ = Optimal_Control_Wave_Equation(N, 2, N, 1)
        u, _ = equ.solve(parameters=default_parameters, complex=True, verbose=False)
        e_g, its = equ.error_norm(u), equ.ksp_its
        u, _ = equ.solve(parameters=None, complex=True, verbose=False)
        e_d = equ.error_norm(u)
        equ.handle.close()
        print(f"{N:>4} {its:>12} {e_g:>14.4e} {e_d:>14.4e} {abs(e_g - e_d):>10.1e} {UPSTREAM[N]:>12.4e} {e_g / UPSTREAM[N]:>14.2f}")
    print("(errors: max over time levels of the nodal 2-norm of u_h - u, u_h[:, i] at t = (i+1) tau; upstream column: "
          "plot.py:5-18, not reproducible from the committed script -- SURVEY section 4)")


if __name__ == "__main__":
    main()
