"""Regenerates tests/golden/upstream_accuracy.json by EXECUTING the upstream script flow.

Upstream's accuracy study (Code/Control_Wave_PC.py:583-631: for N = 5 ... 70, ``equ.solve(...)`` then ``equ.write(...)``,
the returned error norms hard-coded in Code/plot.py:5-18) run here, per N, exactly as the script would run it: the two
class definitions, the solver options :347-359 and the set-up lines :361-372 are read from the upstream checkout at
generation time and executed unmodified with ``fd`` bound to ``firedrake_standin``; then

    u_sol, p_sol = equ.solve(parameters=parameters, complex=True); error = equ.write(u_sol, p_sol)     # pc = True, :567-570
    u_sol, p_sol = equ.solve(complex=True);                        error = equ.write(u_sol, p_sol, name="lu")   # :576-577

``write`` reads entry 25 of the nodal arrays (:281-282, ``xi = x_list[4]``), so the committed script can only run for
N_x >= 25: the study's first four sizes raise IndexError, which is recorded as such.  PETSc's KSPGMRES is not runnable;
the stand-in's solver runs oracle/gmres.py (its restatement) with PETSc's default ksp_rtol = 1e-5 -- the options
leave it open -- restart 300 and max_it 1000 as the options say.  The numbers of plot.py are read from that file.

NON-BINDING by SURVEY section 4 (``write`` / ``error_plot`` / ``plot.py`` are outside the hot path); the point of the
fixture is to settle where the deviation of every restatement from plot.py comes from.

Run from the repo root, where /root/reference exists:  python tests/golden/make_reference_accuracy_golden.py
"""
import contextlib
import io
import json
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
from make_reference_executed_golden import REF, upstream_segments  # noqa: E402
from oracle.gmres import gmres  # noqa: E402

PLOT = os.path.join(os.path.dirname(REF), "plot.py")


class VTKFile:                                   # firedrake.output.VTKFile: the files are not wanted here
    def __init__(self, name):
        self.name = name

    def write(self, *functions):
        pass


def ksp(matvec, pc_apply, b, options):
    assert options["ksp_type"] == "gmres"
    x, its, hist, reason = gmres(matvec, pc_apply, b, rtol=1e-5, atol=1e-50, restart=int(options["ksp_gmres_restart"]),
                                 max_it=int(options["ksp_max_it"]))
    return x, its, hist, reason


def script_namespace(N):
    lines, seg_problem, seg_setup, seg_pc = upstream_segments()
    ns = {"fd": fd, "np": np, "math": math, "time": time, "fft": fft, "ifft": ifft, "VTKFile": VTKFile,
          "pc": True, "complex": True, "T": 2, "N_t": N, "N_x": N, "gamma": 1, "dim": 1, "__name__": "upstream_executed"}
    exec(compile("\n".join(lines[seg_problem[0]:seg_problem[1]]), REF, "exec"), ns)
    ns["equ"] = ns["Optimal_Control_Wave_Equation"](N, 2, N, 1, dim=1)
    first = next(i for i, l in enumerate(lines) if l.startswith("parameters = {"))
    last = next(i for i in range(first, len(lines)) if lines[i].startswith("}"))
    exec(compile("\n".join(lines[first:last + 1]), REF, "exec"), ns)                      # solver options :347-359
    setup_src = [l for l in lines[seg_setup[0]:seg_setup[1]] if not l.lstrip().startswith("#")]
    exec(compile("\n".join(setup_src), REF, "exec"), ns)
    exec(compile("\n".join(lines[seg_pc[0]:seg_pc[1]]), REF, "exec"), ns)
    fd.NonlinearVariationalSolver.python_pcs = {"DiagFFTPC": ns["DiagFFTPC"]}
    fd.NonlinearVariationalSolver.ksp = staticmethod(ksp)
    return ns


def run(N):
    out = {}
    for branch in ("pc", "lu"):
        ns = script_namespace(N)
        equ = ns["equ"]
        with contextlib.redirect_stdout(io.StringIO()):
            if branch == "pc":
                u_sol, p_sol = equ.solve(parameters=ns["parameters"], complex=True)
                out["gmres_its"], _, out["gmres_reason"] = fd.NonlinearVariationalSolver.last
            else:
                u_sol, p_sol = equ.solve(complex=True)
            try:
                out[branch] = float(equ.write(u_sol, p_sol) if branch == "pc" else equ.write(u_sol, p_sol, name="lu"))
            except IndexError as ex:
                out[branch] = f"IndexError: {ex}"
        out[branch + "_u_norm"] = float(np.linalg.norm(u_sol.data))
    return out


def main():
    pl = {}
    with open(PLOT) as fh:
        src = fh.read().splitlines()
    first = next(i for i, l in enumerate(src) if l.startswith("a = np.arange"))
    last = next(i for i in range(first, len(src)) if src[i].rstrip().endswith("])"))
    exec(compile("\n".join(src[first:last + 1]), PLOT, "exec"), {"np": np}, pl)
    published = {int(a): float(b) for a, b in zip(pl["a"], pl["b"])}
    table = {}
    for N in range(5, 75, 5):
        table[N] = run(N)
        table[N]["plot_py"] = published[N]
        print(N, table[N], flush=True)
    with open(os.path.join(HERE, "upstream_accuracy.json"), "w") as fh:
        json.dump({"source": "Control_Wave_PC.py classes, options and set-up lines executed against firedrake_standin; "
                             "plot.py:5-18 read from the upstream file",
                   "ksp": "oracle.gmres (restated KSPGMRES), rtol 1e-5 (PETSc default), restart 300, max_it 1000",
                   "table": table}, fh, indent=1)
    print("written", os.path.join(HERE, "upstream_accuracy.json"))


if __name__ == "__main__":
    sys.exit(main())
