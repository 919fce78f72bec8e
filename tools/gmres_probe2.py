"""Developer probe: which component makes the device GMRES need more iterations at (4096,1024)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
from oracle.operator import AllAtOnce
from oracle.gmres import gmres as ogmres
from oracle import csolve
Nx, Nt = 4096, 1024
op = AllAtOnce(Nx, Nt); pc = DiagFFTPCFast(Nx, Nt, solver=csolve.thomas_toeplitz_c)
with ParaDiagHandle(Nx, Nt) as h:
    bg = h.build_rhs(); b_gpu = bg.cpu().numpy(); b_cpu = op.rhs() + 0j
    pc_gpu = lambda v: h.pc_apply_host(np.ascontiguousarray(v, dtype=np.complex128)).copy()
    def mv_gpu(v):
        return h.matvec(torch.tensor(np.ascontiguousarray(v, dtype=np.complex128), device="cuda:0")).cpu().numpy()
    fmt = lambda hist: ['%.1e' % (v / hist[0]) for v in hist[:10]]
    for name, mv, p, b in (("mvCPU pcCPU bCPU", op.matvec, pc.apply, b_cpu), ("mvCPU pcGPU bCPU", op.matvec, pc_gpu, b_cpu),
                           ("mvGPU pcCPU bCPU", mv_gpu, pc.apply, b_cpu), ("mvGPU pcGPU bCPU", mv_gpu, pc_gpu, b_cpu),
                           ("mvGPU pcGPU bGPU", mv_gpu, pc_gpu, b_gpu)):
        x, its, hist, reason = ogmres(mv, p, b, rtol=1e-7, max_it=30)
        print(name, its, fmt(hist), flush=True)
    x, its, hist, reason = h.gmres(bg, rtol=1e-7, max_it=30)
    print("device GMRES", its, fmt(hist))
    x, its, hist, reason = h.gmres(torch.tensor(b_cpu, device="cuda:0"), rtol=1e-7, max_it=30)
    print("device GMRES, bCPU", its, fmt(hist))
