"""Developer probe: GMRES residual histories at large sizes, GPU vs CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
from oracle.operator import AllAtOnce
from oracle.gmres import gmres as ogmres
from oracle import csolve
for (Nx, Nt, do_cpu) in [(1024, 1024, True), (4096, 1024, True), (4096, 4096, False), (16384, 4096, False)]:
    with ParaDiagHandle(Nx, Nt) as h:
        b = h.build_rhs()
        for rtol in (1e-5, 1e-7):
            t = time.time()
            x, its, hist, reason = h.gmres(b, rtol=rtol, max_it=80)
            torch.cuda.synchronize()
            r = h.matvec(x) - b
            print(f"gpu ({Nx},{Nt}) rtol {rtol}: its {its} {reason} {time.time()-t:.2f}s true res {float(torch.linalg.norm(r)/torch.linalg.norm(b)):.1e} hist {['%.1e' % (v/hist[0]) for v in hist[:9]]}", flush=True)
        if do_cpu:
            op = AllAtOnce(Nx, Nt); pc = DiagFFTPCFast(Nx, Nt, solver=csolve.thomas_toeplitz_c)
            for rtol in (1e-5, 1e-7):
                xo, io, ho, ro = ogmres(op.matvec, pc.apply, op.rhs(), rtol=rtol, max_it=80)
                print(f"cpu ({Nx},{Nt}) rtol {rtol}: its {io} {ro} hist {['%.1e' % (v/ho[0]) for v in ho[:9]]}", flush=True)
