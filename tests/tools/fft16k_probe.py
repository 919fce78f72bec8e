"""N_t = 16384 FFT kernel on the GPU box: parity against scipy and the oracle, then timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scipy.fft as sfft
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast

dev = "cuda:0"
N = 16384


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


perm = np.concatenate([np.arange(q, N, 4) for q in range(4)])
nl = 301
rng = np.random.default_rng(0)
x = rng.standard_normal((nl, N)) + 1j * rng.standard_normal((nl, N))
with ParaDiagHandle(8, N) as h:
    xt = torch.tensor(x, device=dev).reshape(-1)
    yt = torch.empty_like(xt)
    h.stage_fft(xt, yt, nl, True)
    e1 = rel(yt.cpu().numpy().reshape(nl, N), sfft.ifft(x, axis=1)[:, perm])
    xp = torch.tensor(np.ascontiguousarray(x[:, perm]), device=dev).reshape(-1)
    h.stage_fft(xp, yt, nl, False)
    e2 = rel(yt.cpu().numpy().reshape(nl, N), sfft.fft(x, axis=1))
    h.stage_fft(xt, yt, nl, True)
    h.stage_fft(yt, yt, nl, False)
    e3 = rel(yt.cpu().numpy().reshape(nl, N), x)
    print(f"ifft {e1:.2e} fft {e2:.2e} in-place round trip {e3:.2e}", flush=True)
with ParaDiagHandle(20, N) as h:
    xv = rng.standard_normal(h.size) + 1j * rng.standard_normal(h.size)
    ref = DiagFFTPCFast(20, N, 2.0, 1.0).apply(xv)
    y = h.pc_apply(torch.tensor(xv, device=dev)).cpu().numpy()
    print(f"pc apply (20, 16384) vs oracle {rel(y, ref):.2e}", flush=True)

# timing: 2 * 8193 lines = 4.3 GB per sweep
Nx = 8192
S = 32 * (Nx + 1) * N
with ParaDiagHandle(Nx, N) as h:
    x = torch.randn(h.size, dtype=torch.complex128, device=dev)
    w = torch.empty_like(x)
    y = torch.empty_like(x)
    ti = timeit(lambda: h.stage_fft(x, w, 2 * (Nx + 1), True))
    tf = timeit(lambda: h.stage_fft(w, y, 2 * (Nx + 1), False))
    ta = timeit(lambda: h.pc_apply(x, y))
    print(f"ifft {ti:.3f} ms = {2 * S / ti / 1e6:.0f} GB/s, fft {tf:.3f} ms = {2 * S / tf / 1e6:.0f} GB/s, "
          f"apply {ta:.3f} ms = {6 * S / ta / 1e6:.0f} GB/s algorithmic", flush=True)
