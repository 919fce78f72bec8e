"""Regenerates tests/golden/refnotebook_*.npz and refprecond_*.npz by EXECUTING upstream code.

Two more pieces of the reference run without Firedrake / PETSc:

* ``Code/mat_test.ipynb`` -- the only known-answer checks upstream holds (cells 1-12: FFT convention, circulant
  eigenvalues, the analytic 2x2 diagonalisation S, Sigma, Lambda).  The code cells are taken from the notebook's
  JSON AT GENERATION TIME and executed unmodified, cell by cell, in one namespace (as Jupyter would); only the
  literal ``N_t = 5`` of cell 1 and the hard-wired 5-point stencil of cell 8 are re-parameterised for the other
  sizes (the stencil by padding with zeros, which is what the literal does for N_t = 5).
* ``Code/pre_cond.py:32-38`` -- the closed forms Lambda_1, Lambda_2, S1, S2, Gamma, Sigma_1, Sigma_2 of the abandoned
  first draft of the PC (``fd.sqrt`` / ``fd.abs`` there are numpy's ``sqrt`` / ``abs`` on these arrays).

Nothing of the reference is copied into this repository: the lines are read from the upstream checkout, executed,
and only their NUMERICAL OUTPUTS are stored.  tests/test_reference_notebook_golden.py pins ``oracle/eigs.py``
(``lambdas``, ``closed_form``, ``eig_numpy``) and the FFT conventions of the oracle / CUDA path to them.

Run from the repo root, where /root/reference exists:  python tests/golden/make_reference_notebook_golden.py
"""
import ast
import contextlib
import io
import json
import os
import sys
import textwrap
import types

import numpy as np

NB = os.environ.get("PARADIAG_REFERENCE_NB", "/root/reference/Code/mat_test.ipynb")
PRE = os.environ.get("PARADIAG_REFERENCE_PRE", "/root/reference/Code/pre_cond.py")
HERE = os.path.dirname(os.path.abspath(__file__))
NB_CASES = [5, 8, 13, 81]            # 5 = the notebook's own size; 8: N_t % 4 == 0 (lambda_2 = 0 up to rounding)
PRE_CASES = [(5, 2.0, 1.0), (81, 2.0, 1.0), (64, 2.0, 1e-4), (16, 2.0, 1.0)]
PRE_FIRST, PRE_LAST = 32, 38


def run_notebook(N_t):
    nb = json.load(open(NB))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    ns = {}
    values = {}
    with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):   # cell 12 prints a matrix
        for idx, src in enumerate(cells):
            if not src.strip():
                continue
            if "N_t = 5" in src:
                src = src.replace("N_t = 5", f"N_t = {N_t}")
            if "circulant([1,-2,1,0,0])" in src:
                src = src.replace("circulant([1,-2,1,0,0])", "circulant([1,-2,1] + [0] * (N_t - 3))")
            tree = ast.parse(src)
            tail = None
            if tree.body and isinstance(tree.body[-1], ast.Expr):  # a trailing expression = the cell's displayed value
                tail = ast.Expression(tree.body.pop().value)
            exec(compile(tree, f"{NB}:cell{idx}", "exec"), ns)
            if tail is not None:
                values[idx] = eval(compile(tail, f"{NB}:cell{idx}", "eval"), ns)
    keep = ("Lambda_1", "Lambda_2", "S1", "S2", "Gamma", "Sigma_1", "Sigma_2", "S", "Sigma", "Lambda", "A", "B", "C",
            "C1", "LHS", "RHS", "E", "tau", "gamma", "T")
    if N_t > 16:      # keep the fixtures small: vectors and norms only at the larger sizes
        keep = tuple(k for k in keep if np.asarray(ns.get(k, 0)).ndim < 2)
    out = {k: np.asarray(ns[k]) for k in keep if k in ns}
    out["SSh_minus_2I_norm"] = np.linalg.norm(np.asarray(values.get(2)) - 2 * np.eye(2 * N_t))   # cell 2: S S^H = 2 I
    out["norm_B_minus_C"] = np.asarray(values.get(7))      # cell 7
    out["norm_A_minus_C1"] = np.asarray(values.get(9))     # cell 9
    out["norm_E"] = np.asarray(values.get(12))             # cell 12
    return out


def run_precond(N_t, T, gamma):
    lines = open(PRE).read().splitlines()[PRE_FIRST - 1:PRE_LAST]
    assert lines[0].lstrip().startswith("self.Lambda_1 = "), lines[0]
    code = compile(textwrap.dedent("\n".join(lines)), PRE, "exec")
    me = types.SimpleNamespace(N_t=N_t, gamma=gamma)
    ns = {"np": np, "fd": types.SimpleNamespace(sqrt=np.sqrt, abs=np.abs), "self": me, "dt": T / N_t}
    with np.errstate(all="ignore"):
        exec(code, ns)
    return {k: np.asarray(getattr(me, k)) for k in ("Lambda_1", "Lambda_2", "S1", "S2", "Gamma", "Sigma_1", "Sigma_2")}


def main():
    for N_t in NB_CASES:
        np.savez_compressed(os.path.join(HERE, f"refnotebook_{N_t}.npz"), N_t=N_t, **run_notebook(N_t))
    for (N_t, T, gamma) in PRE_CASES:
        np.savez_compressed(os.path.join(HERE, f"refprecond_{N_t}_{gamma:g}.npz"), N_t=N_t, T=T, gamma=gamma,
                            upstream_lines=np.array([PRE_FIRST, PRE_LAST]), **run_precond(N_t, T, gamma))
    print("notebook / pre_cond fixtures written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
