"""Parity of the CUDA path against the oracle AT THE FULL BASELINE SIZES (cfg5 4096 x 4096, cfg3 16384 x 4096,
cfg4 65536 x 16384) -- GPU box only, through the C ABI.

The apply is the composition  fft_t o (per-frequency stage) o ifft_t  (Control_Wave_PC.py:500-501, :445-540,
:547-548).  At these sizes the oracle cannot redo the whole vector in 80-bit arithmetic, but the stages decouple:

  * the time transforms act line by line      -> a sample of lines is compared with scipy.fft (what upstream calls);
  * the per-frequency stage acts column by column -> a sample of >= 64 frequencies, taken from the DEVICE's own
    ifft output, is solved by the oracle in 80-bit (np.longdouble) and in fp64 arithmetic.  Bar: the CUDA columns
    are within 1e-10 (the north-star tolerance) of the 80-bit truth;
  * the one-call apply (pd_pc_apply) must reproduce the staged composition bit for bit;
  * where the host can hold it (cfg5, cfg3) the WHOLE output vector is also compared with the fp64 oracle
    (scipy.fft + threaded C Thomas).  Two correct fp64 algorithms differ by cond * eps ~ 1e-9 here,
    so the tolerance is max(3 e_oracle, 1e-10) with e_oracle = the fp64 oracle's own distance from the 80-bit
    truth measured on the sampled columns.
"""
import numpy as np
import pytest
import scipy.fft as sfft

torch = pytest.importorskip("torch")
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402
from oracle.pc_fast import DiagFFTPCFast  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PC_TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def col_of(k, N_t):
    """Column that holds frequency k in the device's frequency layout ([k mod 4 = 0|1|2|3] at N_t = 16384)."""
    k = np.asarray(k)
    return (k % 4) * (N_t // 4) + k // 4 if N_t == 16384 else k


def sample_freqs(N_t, count=64, seed=5):
    special = [0, 1, 2, 3, N_t // 4 - 1, N_t // 4, N_t // 4 + 1, N_t // 2 - 1, N_t // 2, N_t // 2 + 1,
               3 * N_t // 4, N_t - 2, N_t - 1]
    rng = np.random.default_rng(seed)
    ks = set(special) | set(rng.choice(N_t, size=count, replace=False).tolist())
    return np.array(sorted(ks))


def staged_checks(h, x, N_x, N_t, gamma=1.0, tol=PC_TOL):
    """Runs ifft / stage / fft on the device (in place on one scratch vector) and checks every stage on samples.
    Returns (y as a device tensor, e_cuda, e_oracle)."""
    n = N_x + 1
    perm = np.concatenate([np.arange(q, N_t, 4) for q in range(4)]) if N_t == 16384 else np.arange(N_t)
    rng = np.random.default_rng(11)
    lines = np.unique(np.concatenate([[0, n - 1, n, 2 * n - 1], rng.choice(2 * n, size=44, replace=False)]))
    lt = torch.tensor(lines, device=DEV)
    w = torch.empty_like(x)

    # ---- :500-501 ifft along time, sampled lines against scipy
    h.stage_fft(x, w, 2 * n, True)
    X = x.view(2 * n, N_t)[lt].cpu().numpy()
    XH = w.view(2 * n, N_t)[lt].cpu().numpy()
    assert rel(XH, sfft.ifft(X, axis=1)[:, perm]) < 5e-15

    # ---- :445-540 per-frequency stage, sampled columns against the 80-bit and the fp64 oracle
    ks = sample_freqs(N_t)
    ct = torch.tensor(col_of(ks, N_t), device=DEV)
    xh_cols = w.view(2, n, N_t)[:, :, ct].cpu().numpy()
    h.stage_solve(w)
    w_cols = w.view(2, n, N_t)[:, :, ct].cpu().numpy()
    truth = DiagFFTPCFast(N_x, N_t, 2.0, gamma, dtype=np.longdouble).stage_columns(ks, xh_cols)
    f64 = DiagFFTPCFast(N_x, N_t, 2.0, gamma).stage_columns(ks, xh_cols)
    e_cuda, e_oracle = rel(w_cols, truth), rel(f64, truth)
    assert e_cuda < tol, (e_cuda, e_oracle)
    assert e_cuda < 0.05 * e_oracle or e_oracle < 1e-11, (e_cuda, e_oracle)   # and far closer than fp64 LU is
    # per column as well: no single frequency may hide behind the others
    percol = np.linalg.norm((w_cols - truth).reshape(-1, len(ks)), axis=0) / np.linalg.norm(
        truth.reshape(-1, len(ks)).astype(np.complex128), axis=0)
    assert float(percol.max()) < 10 * tol, (ks[int(percol.argmax())], float(percol.max()))
    assert np.abs(w_cols[:, [0, -1], :]).max() == 0.0            # Dirichlet rows exactly zero (:482)

    # ---- :547-548 fft along time, sampled lines against scipy
    W = w.view(2 * n, N_t)[lt].cpu().numpy()
    h.stage_fft(w, w, 2 * n, False)
    Y = w.view(2 * n, N_t)[lt].cpu().numpy()
    Wnat = np.empty_like(W)
    Wnat[:, perm] = W
    ref = sfft.fft(Wnat, axis=1)
    nz = np.linalg.norm(ref, axis=1) > 0
    assert rel(Y[nz], ref[nz]) < 5e-15
    assert np.abs(Y[~nz]).max(initial=0.0) == 0.0
    return w, e_cuda, e_oracle


def device_random(size, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.empty(size, dtype=torch.complex128, device=DEV)
    xr = torch.view_as_real(x)
    step = 1 << 27
    for o in range(0, size, step):
        xr[o:o + min(step, size - o)].normal_(generator=g)
    return x


@pytest.mark.parametrize("N_x,N_t", [(4096, 4096), (16384, 4096)])
def test_full_size_apply_against_fp64_and_80bit_oracle(N_x, N_t):
    """cfg5 and cfg3: whole vector against the fp64 oracle, sampled columns against the 80-bit oracle."""
    n = N_x + 1
    size = 2 * n * N_t
    rng = np.random.default_rng(0)
    xn = np.empty(size, dtype=np.complex128)
    step = 1 << 24
    for o in range(0, size, step):
        m = min(step, size - o)
        xn[o:o + m] = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    with ParaDiagHandle(N_x, N_t) as h:
        x = torch.tensor(xn, device=DEV)
        y_staged, e_cuda, e_oracle = staged_checks(h, x, N_x, N_t)
        y = h.pc_apply(x)
        assert torch.equal(y, y_staged)                            # the one-call apply = the staged composition
        del y_staged
        yh = h.pc_apply_host(xn)                                   # the drop-in entry point: host buffers
        assert np.array_equal(yh, y.cpu().numpy())
        del x, y
    torch.cuda.empty_cache()
    ref = DiagFFTPCFast(N_x, N_t).apply_threaded(xn)
    err = rel(yh, ref)
    assert err < max(3 * e_oracle, PC_TOL), (err, e_oracle, e_cuda)
    assert np.abs(yh.reshape(2, n, N_t)[:, [0, -1], :]).max() == 0.0


def test_cfg4_apply_sampled_against_80bit_oracle():
    """cfg4 (65536 x 16384, 34 GB vectors): device-generated input, every stage checked on samples, the one-call
    apply bit-identical to the staged composition.  Needs ~105 GB of device memory.

    Tolerance: at N_x = 65536 the per-frequency systems have cond ~ 12 / h^2 = 5e10; fp64 LU -- the oracle's Thomas
    and upstream's MUMPS alike -- is 2.6e-7 away from the 80-bit truth on these columns (measured, B200 run of this
    test), so "1e-10 against the reference" is not a meaningful bar here.  The CUDA path (detuning-form
    coefficients) measured 1.5e-10; asserted: < 5e-10 against the 80-bit truth and >= 20x closer than fp64 LU."""
    N_x, N_t = 65536, 16384
    free, _ = torch.cuda.mem_get_info()
    if free < 110 * (1 << 30):
        pytest.skip(f"needs 110 GB of free device memory, {free >> 30} GB available")
    with ParaDiagHandle(N_x, N_t) as h:
        x = device_random(h.size)
        y_staged, e_cuda, e_oracle = staged_checks(h, x, N_x, N_t, tol=5e-10)
        h.pc_apply(x, x)                                           # in place: no fourth 34 GB vector
        assert torch.equal(x, y_staged)
        Y = x.view(2, N_x + 1, N_t)
        assert float(Y[:, [0, -1], :].abs().max()) == 0.0
        del x, y_staged, Y
    torch.cuda.empty_cache()


@pytest.mark.parametrize("N_x,N_t", [(4096, 4096), (16384, 4096)])
def test_full_size_real_input_path_against_complex_path(N_x, N_t):
    """pd_pc_apply_real at cfg5 / cfg3 equals the real part of the (oracle-checked) complex apply."""
    with ParaDiagHandle(N_x, N_t) as h:
        g = torch.Generator(device=DEV).manual_seed(1)
        xr = torch.randn(h.size, dtype=torch.float64, device=DEV, generator=g)
        yr = h.pc_apply_real(xr)
        yc = h.pc_apply(xr.to(torch.complex128))
        err = float(torch.linalg.norm(yr - yc.real) / torch.linalg.norm(yc.real))
        assert err < 1e-12, err
        assert float(yc.imag.abs().max() / yc.real.abs().max()) < 1e-8
