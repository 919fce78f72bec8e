"""Developer probe: where does the PC-apply error on SMOOTH vectors come from? (vs 80-bit oracle)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, scipy.fft as sfft
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
from oracle.operator import AllAtOnce
dev = "cuda:0"
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for (Nx, Nt) in [(1024, 1024), (4096, 1024), (4096, 256)]:
    n = Nx + 1
    b = AllAtOnce(Nx, Nt).rhs() + 0j
    rng = np.random.default_rng(0)
    xr = rng.standard_normal(b.size) + 1j * rng.standard_normal(b.size)
    f64 = DiagFFTPCFast(Nx, Nt); ld = DiagFFTPCFast(Nx, Nt, dtype=np.longdouble)
    with ParaDiagHandle(Nx, Nt) as h:
        for name, x in (("smooth", b), ("random", xr)):
            truth = ld.apply(x)
            print(f"({Nx},{Nt}) {name}: full apply  cuda-ld {rel(h.pc_apply_host(x), truth):.2e}  oracle64-ld {rel(f64.apply(x), truth):.2e}", flush=True)
            # stage 1 alone
            xt = torch.tensor(x, device=dev); w = torch.empty_like(xt)
            h.stage_fft(xt, w, 2 * n, True)
            xh_ld = sfft.ifft(x.astype(np.clongdouble).reshape(2, n, Nt), axis=2)
            print(f"      ifft: cuda-ld {rel(w.cpu().numpy().reshape(2,n,Nt), xh_ld):.2e} scipy64-ld {rel(sfft.ifft(x.reshape(2,n,Nt),axis=2), xh_ld):.2e}")
            # stage 2 alone, fed with the rounded 80-bit spectrum
            xh = xh_ld.astype(np.complex128)
            w2 = torch.tensor(xh.reshape(-1), device=dev)
            h.stage_solve(w2)
            def solve_stage(pc, xh_):
                uz = xh_[0] * np.conj(pc.z); ip = (1j * pc.sigma) * xh_[1]
                zp, zm = pc.solve_stage((uz + ip) / 2, (uz - ip) / 2)
                return np.stack([zp + zm, (-1j * pc.sigma * pc.z) * (zp - zm)])
            s_ld = solve_stage(ld, xh.astype(np.clongdouble)); s_64 = solve_stage(f64, xh)
            print(f"      solve: cuda-ld {rel(w2.cpu().numpy().reshape(2,n,Nt), s_ld):.2e} oracle64-ld {rel(s_64, s_ld):.2e}")
            # per-frequency worst
            d = w2.cpu().numpy().reshape(2, n, Nt) - s_ld
            pk = np.sqrt((np.abs(d) ** 2).sum(axis=(0, 1))) / np.sqrt((np.abs(s_ld) ** 2).sum(axis=(0, 1)) + 1e-300)
            d64 = s_64 - s_ld
            pk64 = np.sqrt((np.abs(d64) ** 2).sum(axis=(0, 1))) / np.sqrt((np.abs(s_ld) ** 2).sum(axis=(0, 1)) + 1e-300)
            k = int(np.argmax(pk)); print(f"      worst k (cuda) {k}: {float(pk[k]):.2e} (oracle64 there {float(pk64[k]):.2e}); k=0..5 cuda {[f'{float(v):.1e}' for v in pk[:6]]} oracle {[f'{float(v):.1e}' for v in pk64[:6]]}")
