"""Developer smoke: parity of every kernel against the oracle + first timings (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
from oracle.operator import AllAtOnce
from oracle.gmres import gmres as ogmres

dev = "cuda:0"
print(torch.cuda.get_device_name(0), flush=True)

def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)

# 1. FFT alone
for Nt in (13, 81, 64, 96, 100, 128, 256, 512, 1024, 2048, 4096, 8192, 97):
    with ParaDiagHandle(8, Nt) as h:
        rng = np.random.default_rng(0)
        nl = 37
        x = rng.standard_normal((nl, Nt)) + 1j * rng.standard_normal((nl, Nt))
        xt = torch.tensor(x, device=dev).reshape(-1)
        yt = torch.empty_like(xt)
        h.stage_fft(xt, yt, nl, False)
        e1 = rel(yt.cpu().numpy().reshape(nl, Nt), np.fft.fft(x, axis=1))
        h.stage_fft(xt, yt, nl, True)
        e2 = rel(yt.cpu().numpy().reshape(nl, Nt), np.fft.ifft(x, axis=1))
        print(f"fft N_t={Nt}: fwd {e1:.2e} inv {e2:.2e}", flush=True)

# 2. PC apply
for (Nx, Nt, g) in [(16, 13, 1.0), (20, 81, 1.0), (80, 81, 1.0), (16, 16, 1.0), (24, 64, 1e-4), (17, 64, 1.0), (18, 64, 1.0), (34, 64, 1.0), (35,64,1.0),(40, 96, 1e-2),
                    (100, 128, 1.0), (256, 256, 1.0), (1024, 1024, 1.0)]:
    with ParaDiagHandle(Nx, Nt, gamma=g) as h:
        rng = np.random.default_rng(0)
        x = rng.standard_normal(h.size) + 1j * rng.standard_normal(h.size)
        ref = DiagFFTPCFast(Nx, Nt, 2.0, g).apply(x)
        y = h.pc_apply(torch.tensor(x, device=dev)).cpu().numpy()
        yh = h.pc_apply_host(x)
        bnd = np.abs(y.reshape(2, Nx + 1, Nt)[:, [0, -1], :]).max()
        print(f"pc ({Nx},{Nt},{g}): dev {rel(y, ref):.2e} host {rel(yh, ref):.2e} boundary {bnd:.1e}", flush=True)

# 3. matvec / rhs / gmres
for (Nx, Nt, g) in [(16, 13, 1.0), (80, 81, 1.0), (32, 64, 1e-2), (64, 256, 1e-4)]:
    with ParaDiagHandle(Nx, Nt, gamma=g) as h:
        op = AllAtOnce(Nx, Nt, 2.0, g)
        rng = np.random.default_rng(1)
        x = rng.standard_normal(h.size) + 1j * rng.standard_normal(h.size)
        y = h.matvec(torch.tensor(x, device=dev)).cpu().numpy()
        b = h.build_rhs().cpu().numpy()
        print(f"op ({Nx},{Nt},{g}): matvec {rel(y, op.matvec(x)):.2e} rhs {rel(b, op.rhs()):.2e}", flush=True)
        pc = DiagFFTPCFast(Nx, Nt, 2.0, g)
        xo, io, ho, ro = ogmres(op.matvec, pc.apply, op.rhs(), rtol=1e-7)
        xg, ig, hg, rg = h.gmres(h.build_rhs(), rtol=1e-7)
        print(f"   gmres its oracle {io} gpu {ig} {rg}; x diff {rel(xg.cpu().numpy(), xo):.2e}; hist gpu {['%.2e' % v for v in hg]}", flush=True)

# 4. timings
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for (Nx, Nt) in [(80, 81), (1024, 1024), (4096, 4096), (16384, 4096)]:
    with ParaDiagHandle(Nx, Nt) as h:
        x = torch.randn(h.size, dtype=torch.complex128, device=dev)
        y = torch.empty_like(x)
        w = torch.empty_like(x)
        S = 32 * (Nx + 1) * Nt
        t_all = timeit(lambda: h.pc_apply(x, y))
        t_f = timeit(lambda: h.stage_fft(x, w, 2 * (Nx + 1), True))
        t_s = timeit(lambda: h.stage_solve(w))
        t_b = timeit(lambda: h.stage_fft(w, y, 2 * (Nx + 1), False))
        t_cp = timeit(lambda: y.copy_(x))
        t_mv = timeit(lambda: h.matvec(x, y))
        print(f"time ({Nx},{Nt}): apply {t_all:.3f} ms = {6*S/t_all/1e6:.0f} GB/s algorithmic ({6*S/t_all/1e6/6547:.2%} of 6547); "
              f"ifft {t_f:.3f} solve {t_s:.3f} fft {t_b:.3f}; copy {t_cp:.3f} ms = {2*S/t_cp/1e6:.0f} GB/s; matvec {t_mv:.3f}", flush=True)
