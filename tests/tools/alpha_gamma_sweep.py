"""BASELINE config 5: alpha x gamma sweep on N_x = 4096, N_t = 4096 (GPU box).

For every (alpha, gamma): GMRES iterations / reason / time-to-solution on the manufactured right-hand side
(Build_f/g/IC), the TRUE relative residual at exit, and the PC-apply error against the CPU oracle on a size the
oracle finishes in seconds (same alpha, gamma).  alpha = 1 is the upstream operator; alpha != 1 is the extension
defined in oracle/pc_alpha.py (no upstream behaviour exists for it: parity unpinned).
Writes one JSON object per line to stdout.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_alpha import DiagFFTPCAlpha
from oracle.pc_fast import DiagFFTPCFast

N_x = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N_t = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 120
dev = "cuda:0"
SN_x, SN_t = 256, 256     # oracle-sized accuracy check


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


for gamma in (1.0, 1e-2, 1e-4, 1e-6):
    for alpha in (1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6):
        rec = {"N_x": N_x, "N_t": N_t, "gamma": gamma, "alpha": alpha}
        with ParaDiagHandle(SN_x, SN_t, gamma=gamma, alpha=alpha) as hs:
            rng = np.random.default_rng(0)
            x = rng.standard_normal(hs.size) + 1j * rng.standard_normal(hs.size)
            ora = DiagFFTPCFast(SN_x, SN_t, 2.0, gamma) if alpha == 1.0 else DiagFFTPCAlpha(SN_x, SN_t, 2.0, gamma, alpha)
            rec["apply_err_vs_oracle_256x256"] = rel(hs.pc_apply(torch.tensor(x, device=dev)).cpu().numpy(), ora.apply(x))
        with ParaDiagHandle(N_x, N_t, gamma=gamma, alpha=alpha) as h:
            b = h.build_rhs()
            xt = torch.randn(h.size, dtype=torch.complex128, device=dev)
            yt = torch.empty_like(xt)
            for _ in range(3):
                h.pc_apply(xt, yt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                h.pc_apply(xt, yt)
            e1.record()
            torch.cuda.synchronize()
            rec["apply_ms"] = e0.elapsed_time(e1) / 10
            del xt, yt
            t0 = time.perf_counter()
            sol, its, hist, reason = h.gmres(b, rtol=1e-7, max_it=max_it)
            torch.cuda.synchronize()
            rec.update(gmres_its=its, gmres_reason=reason, gmres_seconds=time.perf_counter() - t0,
                       precond_residual_drop=float(hist[-1] / hist[0]),
                       true_rel_residual=float(torch.linalg.norm(h.matvec(sol) - b) / torch.linalg.norm(b)))
        print(json.dumps(rec), flush=True)
