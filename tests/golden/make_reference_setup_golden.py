"""Regenerates tests/golden/refsetup_*.npz by EXECUTING upstream code.

The one part of ``DiagFFTPC`` that runs without Firedrake / PETSc is the eigen-setup of
``DiagFFTPC.initialize`` (Code/Control_Wave_PC.py:387-436): plain numpy on the module globals ``N_t, dt,
gamma``.  This script reads those lines from the upstream checkout AT GENERATION TIME (nothing of the
reference is copied into this repository), executes them unmodified against a dummy ``self`` and stores what
they compute -- Lambda_1, Lambda_2, the eigenvalues Sigma_1/2 and the eigenvector matrices S, S^-1 that the
upstream loop keeps in S11..SI22.  tests/test_reference_setup_golden.py pins the oracle (``oracle/eigs.py``,
and through it the line-by-line route and the closed forms the CUDA kernels regenerate) to these outputs.

Run from the repo root, where /root/reference exists:  python tests/golden/make_reference_setup_golden.py
"""
import os
import sys
import textwrap
import types

import numpy as np

REF = os.environ.get("PARADIAG_REFERENCE", "/root/reference/Code/Control_Wave_PC.py")
HERE = os.path.dirname(os.path.abspath(__file__))
FIRST, LAST = 387, 436          # "self.Lambda_1 = ..." through the end of the per-frequency eig loop
CASES = [(81, 2.0, 1.0), (16, 2.0, 1.0), (64, 2.0, 1e-4), (5, 2.0, 1.0)]
KEEP = ("Lambda_1", "Lambda_2", "S11", "S12", "S21", "S22", "SI11", "SI12", "SI21", "SI22")


def main():
    lines = open(REF).read().splitlines()[FIRST - 1:LAST]
    assert lines[0].lstrip().startswith("self.Lambda_1 = "), lines[0]
    code = compile(textwrap.dedent("\n".join(lines)), REF, "exec")
    for (N_t, T, gamma) in CASES:
        ns = {"np": np, "N_t": N_t, "dt": T / N_t, "gamma": gamma, "self": types.SimpleNamespace()}
        with np.errstate(all="ignore"):      # N_t % 4 == 0: lambda_2 = 0 up to rounding, as upstream
            exec(code, ns)
        out = {k: getattr(ns["self"], k) for k in KEEP}
        out.update(Sigma_1=ns["Sigma_1"], Sigma_2=ns["Sigma_2"])
        np.savez_compressed(os.path.join(HERE, f"refsetup_{N_t}_{gamma:g}.npz"), N_t=N_t, T=T, gamma=gamma,
                            upstream_lines=np.array([FIRST, LAST]), **out)
    print("reference set-up fixtures written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
