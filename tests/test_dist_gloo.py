"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic: slab partition, pack/unpack and
the two all-to-all transposes of ``DistributedDiagFFTPC``.  The compute stages are stood in by the
oracle (tests only); on the GPU box the same class drives libparadiag (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from optimal_control_paradiag_b200.dist import DistributedDiagFFTPC, slab_bounds  # noqa: E402


class OracleStageBackend:
    """pd_stage_fft / pd_stage_solve semantics on CPU tensors, from the oracle's fast route."""
    launch_count = 0

    def __init__(self, N_x, N_t, T, gamma, k_begin=0, k_count=0, n_local=0, slab_rank=0, slab_count=0, alpha=1.0):
        from oracle.pc_fast import DiagFFTPCFast
        self.pc = DiagFFTPCFast(N_x, N_t, T, gamma, workers=1)
        self.cf = None                                   # alpha != 1 (extension): oracle/pc_alpha.py's closed form
        if alpha != 1.0:
            from oracle.pc_alpha import decoupled_coeffs
            self.cf = decoupled_coeffs(N_x, N_t, T, gamma, alpha)
        self.N_t, self.n = N_t, N_x + 1
        k_count = k_count or N_t
        self.ks = slice(k_begin, k_begin + k_count)
        self.k_count = k_count
        self.G, self.r = slab_count, slab_rank
        if slab_count > 1:
            counts, offs = slab_bounds(self.n, slab_count)
            self.n_r = counts[slab_rank]
            self.body = [c - 1 - (1 if s == slab_count - 1 else 0) for s, c in enumerate(counts)]

    # ---- slab mode, restated with dense little solves (SPIKE): independent of the CUDA algorithm
    def _rot_in(self, W, kidx=None):
        pc = self.pc
        kidx = np.arange(self.N_t) if kidx is None else kidx
        if self.cf is not None:
            cf = self.cf
            rp = cf["gp"][kidx] * W[0] + cf["e"][kidx] * W[1]
            rm = cf["gm"][kidx] * W[0] - cf["e"][kidx] * W[1]
            return rp, np.conj(rm)
        uz = W[0] * np.conj(pc.z[kidx])
        ip = (1j * pc.sigma[kidx]) * W[1]
        return (uz + ip) / 2, np.conj((uz - ip) / 2)          # slot +, conj slot -

    def _ab(self, k):
        return (self.cf["off"][k], self.cf["diag"][k]) if self.cf is not None else (self.pc.a[k], self.pc.b[k])

    def _T(self, m, k):
        a, b = self._ab(k)
        return (np.diag(np.full(m, b)) + np.diag(np.full(m - 1, a), 1) + np.diag(np.full(m - 1, a), -1))

    def slab_reduce(self, w, out, kidx=None):
        kidx = np.arange(self.N_t) if kidx is None else kidx       # column -> frequency
        K = len(kidx)
        W = w.numpy().reshape(2, self.n_r, K)
        rP, rM = self._rot_in(W, kidx)
        m = self.body[self.r]
        o = out.numpy().reshape(6, K)
        for col, k in enumerate(kidx):
            T = self._T(m, k)
            yP = np.linalg.solve(T, rP[1:1 + m, col])
            yM = np.linalg.solve(T, rM[1:1 + m, col])
            o[0, col], o[1, col], o[2, col], o[3, col] = yP[0], yM[0], yP[-1], yM[-1]
            o[4, col], o[5, col] = (rP[0, col], rM[0, col]) if self.r > 0 else (0, 0)

    # ---- real-input path: half spectrum, rows padded to Kp columns (padding columns hold zeros)
    @property
    def half_cols(self):
        return (self.N_t // 2 + 1 + 7) & ~7

    def _half_kidx(self):
        return np.minimum(np.arange(self.half_cols), self.N_t // 2)

    def stage_rfft_pair(self, src, dst, nnodes, to_freq):
        N, Kp, H = self.N_t, self.half_cols, self.N_t // 2
        if to_freq:
            x = src.numpy().reshape(2, nnodes, N)
            out = np.zeros((2, nnodes, Kp), complex)
            out[..., :H + 1] = np.fft.ifft(x, axis=2)[..., :H + 1]
            dst.copy_(torch.from_numpy(out.reshape(-1)))
        else:
            Y = src.numpy().reshape(2, nnodes, Kp)[..., :H + 1]
            y = N * np.fft.irfft(np.conj(Y), n=N, axis=2)            # fft of the Hermitian extension
            dst.copy_(torch.from_numpy(np.ascontiguousarray(y).reshape(-1)))

    def slab_reduce_half(self, w, out):
        self.slab_reduce(w, out, self._half_kidx())

    def slab_finish_half(self, w, gathered):
        self.slab_finish(w, gathered, self._half_kidx())

    # ---- Krylov pieces (numpy), same contracts as ParaDiagHandle
    def _op(self):
        if not hasattr(self, "_aao"):
            from oracle.operator import AllAtOnce
            self._aao = AllAtOnce(self.pc.N_x, self.N_t, self.pc.T, self.pc.gamma)
            _, self._offs = slab_bounds(self.n, self.G)
        return self._aao

    def build_rhs(self, b):
        op = self._op()
        j0 = self._offs[self.r]
        b.copy_(torch.from_numpy(op.rhs().reshape(2, self.n, self.N_t)[:, j0:j0 + self.n_r].reshape(-1) + 0j))
        return b

    def matvec_slab(self, x, lo, hi, y):
        op = self._op()
        j0 = self._offs[self.r]
        full = np.zeros((2, self.n, self.N_t), complex)      # only the slab and its halo rows matter
        full[:, j0:j0 + self.n_r] = x.numpy().reshape(2, self.n_r, self.N_t)
        if lo is not None:
            full[:, j0 - 1] = lo.numpy().reshape(2, self.N_t)
        if hi is not None:
            full[:, j0 + self.n_r] = hi.numpy().reshape(2, self.N_t)
        out = op.matvec(full.reshape(-1)).reshape(2, self.n, self.N_t)[:, j0:j0 + self.n_r]
        y.copy_(torch.from_numpy(np.ascontiguousarray(out).reshape(-1)))
        return y

    def mdot(self, V, w):
        return torch.from_numpy(V.numpy().conj() @ w.numpy())

    def maxpy(self, V, coef, sign, w, norm2_out=None):
        wn = w.numpy()
        wn += sign * (coef.numpy()[: V.shape[0]] @ V.numpy())
        if norm2_out is not None:
            norm2_out[0] = float(np.vdot(wn, wn).real)
        return w

    def slab_finish(self, w, gathered, kidx=None):
        pc, G, r = self.pc, self.G, self.r
        kidx = np.arange(self.N_t) if kidx is None else kidx
        K = len(kidx)
        W = w.numpy().reshape(2, self.n_r, K)
        g = gathered.numpy().reshape(G, 6, K)
        rP, rM = self._rot_in(W, kidx)
        m = self.body[r]
        for col, k in enumerate(kidx):
            a, b = self._ab(k)
            inv = [np.linalg.inv(self._T(ms, k)) for ms in self.body]
            A = np.zeros((G - 1, G - 1), complex)
            rhs = np.zeros((G - 1, 2), complex)
            for s in range(1, G):                         # separator s between slab s-1 and slab s
                A[s - 1, s - 1] = b - a * a * (inv[s - 1][-1, -1] + inv[s][0, 0])
                if s > 1:
                    A[s - 1, s - 2] = -a * a * inv[s - 1][-1, 0]
                if s < G - 1:
                    A[s - 1, s] = -a * a * inv[s][0, -1]
                rhs[s - 1, 0] = g[s, 4, col] - a * (g[s - 1, 2, col] + g[s, 0, col])
                rhs[s - 1, 1] = g[s, 5, col] - a * (g[s - 1, 3, col] + g[s, 1, col])
            zs = np.linalg.solve(A, rhs)
            zl = zs[r - 1] if r > 0 else np.zeros(2)
            zr = zs[r] if r < G - 1 else np.zeros(2)
            out = []
            for slot, rr in ((0, rP), (1, rM)):
                v = rr[1:1 + m, col].copy()
                v[0] -= a * zl[slot]
                v[-1] -= a * zr[slot]
                out.append(np.linalg.solve(self._T(m, k), v))
            zP = np.zeros(self.n_r, complex)
            zM = np.zeros(self.n_r, complex)
            zP[1:1 + m], zM[1:1 + m] = out[0], np.conj(out[1])
            zP[0], zM[0] = zl[0], np.conj(zl[1])
            if self.cf is not None:
                W[0, :, col] = self.cf["eic"][k] * (zP + zM)
                W[1, :, col] = 1j * (self.cf["bmd"][k] * zP + self.cf["bpd"][k] * zM)
            else:
                W[0, :, col] = zP + zM
                W[1, :, col] = (-1j * pc.sigma[k] * pc.z[k]) * (zP - zM)

    def stage_fft(self, src, dst, nlines, inverse):
        import scipy.fft as sfft
        a = src.numpy().reshape(nlines, self.N_t)
        out = sfft.ifft(a, axis=1) if inverse else sfft.fft(a, axis=1)
        dst.copy_(torch.from_numpy(out.reshape(-1)))

    def stage_solve(self, w):
        pc, ks = self.pc, self.ks
        W = w.numpy().reshape(2, self.n, self.k_count)
        z, sg = pc.z[ks], pc.sigma[ks]
        uz = W[0] * np.conj(z)
        ip = (1j * sg) * W[1]
        rp, rm = (uz + ip) / 2, (uz - ip) / 2
        zp, zm = np.zeros_like(rp), np.zeros_like(rm)
        zp[1:-1] = pc.solver(pc.a[ks], pc.b[ks], rp[1:-1])
        zm[1:-1] = np.conj(pc.solver(pc.a[ks], pc.b[ks], np.conj(rm[1:-1])))
        W[0] = zp + zm
        W[1] = (-1j * sg * z) * (zp - zm)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N_x, N_t, gamma, ret, mode="alltoall"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pc_fast import DiagFFTPCFast
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, gamma, **kw)
        dpc = DistributedDiagFFTPC(N_x, N_t, T=2.0, gamma=gamma, backend_factory=factory, mode=mode)
        rng = np.random.default_rng(0)
        size = 2 * (N_x + 1) * N_t
        xg = rng.standard_normal(size) + 1j * rng.standard_normal(size)
        x_local = dpc.scatter_from_global(torch.from_numpy(xg))
        y_local = dpc.apply(x_local)
        yg = dpc.gather_to_global(y_local).numpy()
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(xg)
        err = np.linalg.norm(yg - ref) / np.linalg.norm(ref)
        # host-buffer entry point (numpy blocks in, numpy blocks out): same result
        yh = np.empty(dpc.local_size, dtype=complex)
        dpc.apply_host(x_local.numpy().copy(), yh)
        err = max(err, float(np.abs(yh - y_local.numpy()).max()))
        d = dpc.describe()
        ok = err < 1e-11 and sum(d["node_slabs"]) == N_x + 1 and (mode == "slab" or sum(d["freq_slabs"]) == N_t)
        ret[rank] = (bool(ok), float(err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N_x,N_t", [(2, 16, 12), (2, 21, 9), (3, 20, 16)])
def test_distributed_apply_matches_single_process_oracle(world, N_x, N_t):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), N_x, N_t, 0.5, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, err = ret[r]
        assert ok, (r, err)


def _real_worker(rank, world, port, N_x, N_t, gamma, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pc_fast import DiagFFTPCFast
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, gamma, **kw)
        dpc = DistributedDiagFFTPC(N_x, N_t, T=2.0, gamma=gamma, backend_factory=factory, mode="slab")
        xg = np.random.default_rng(0).standard_normal(2 * (N_x + 1) * N_t)
        x_local = dpc.scatter_from_global(torch.from_numpy(xg))               # float64 block
        y_local = dpc.apply_real(x_local)
        assert y_local.dtype == torch.float64
        yg = dpc.gather_to_global(y_local.to(torch.complex128)).numpy()
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(xg + 0j)
        ret[rank] = float(np.linalg.norm(yg - ref.real) / np.linalg.norm(ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N_x,N_t", [(2, 16, 8), (3, 22, 16)])
def test_slab_mode_real_input_apply_matches_single_process_oracle(world, N_x, N_t):
    # host logic of DistributedDiagFFTPC.apply_real (half spectrum through the slab-distributed solve)
    ret = mp.Manager().dict()
    mp.spawn(_real_worker, args=(world, _free_port(), N_x, N_t, 0.5, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] < 1e-11, (r, ret[r])


def _alpha_worker(rank, world, port, N_x, N_t, alpha, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pc_alpha import DiagFFTPCAlpha
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, 1.0, alpha=alpha, **kw)
        dpc = DistributedDiagFFTPC(N_x, N_t, T=2.0, gamma=1.0, backend_factory=factory, mode="slab", alpha=alpha)
        rng = np.random.default_rng(2)
        size = 2 * (N_x + 1) * N_t
        xg = rng.standard_normal(size) + 1j * rng.standard_normal(size)
        ref = DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha).apply(xg)
        x_local = dpc.scatter_from_global(torch.from_numpy(xg))
        x_keep = x_local.clone()
        yg = dpc.gather_to_global(dpc.apply(x_local)).numpy()
        assert torch.equal(x_local, x_keep)                                   # the input block is not scaled in place
        xr = rng.standard_normal(size)
        refr = DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha).apply(xr + 0j).real
        yr = dpc.gather_to_global(dpc.apply_real(dpc.scatter_from_global(torch.from_numpy(xr))).to(torch.complex128))
        ret[rank] = (float(np.linalg.norm(yg - ref) / np.linalg.norm(ref)),
                     float(np.linalg.norm(yr.numpy().real - refr) / np.linalg.norm(refr)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N_x,N_t,alpha", [(2, 16, 8, 0.5), (3, 22, 16, 1e-2)])
def test_slab_mode_alpha_extension_matches_single_process_oracle(world, N_x, N_t, alpha):
    # alpha != 1 (no upstream counterpart) through the collective transport: Gamma / Gamma^-1 as elementwise products
    # around the slab-distributed solve, complex and real-input applies, against oracle/pc_alpha.py
    ret = mp.Manager().dict()
    mp.spawn(_alpha_worker, args=(world, _free_port(), N_x, N_t, alpha, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r][0] < 1e-10 and ret[r][1] < 1e-10, (r, ret[r])


@pytest.mark.parametrize("world,N_x,N_t", [(2, 16, 6), (3, 22, 5)])
def test_slab_mode_apply_matches_single_process_oracle(world, N_x, N_t):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), N_x, N_t, 0.5, ret, "slab"), nprocs=world, join=True)
    for r in range(world):
        ok, err = ret[r]
        assert ok, (r, err)


def _gmres_worker(rank, world, port, N_x, N_t, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.gmres import gmres as ogmres
        from oracle.operator import AllAtOnce
        from oracle.pc_fast import DiagFFTPCFast
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, 1.0, **kw)
        dpc = DistributedDiagFFTPC(N_x, N_t, T=2.0, gamma=1.0, backend_factory=factory, mode="slab")
        b = dpc.build_rhs()
        x, its, hist, reason = dpc.gmres(b, rtol=1e-7)
        xg = dpc.gather_to_global(x).numpy()
        op = AllAtOnce(N_x, N_t)
        xo, its_o, hist_o, _ = ogmres(op.matvec, DiagFFTPCFast(N_x, N_t).apply, op.rhs() + 0j, rtol=1e-7)
        # distributed matvec against the global one
        rng = np.random.default_rng(1)
        vg = rng.standard_normal(2 * (N_x + 1) * N_t) + 0j
        yl = dpc.matvec(dpc.scatter_from_global(torch.from_numpy(vg)))
        mv_err = np.linalg.norm(dpc.gather_to_global(yl).numpy() - op.matvec(vg)) / np.linalg.norm(op.matvec(vg))
        ret[rank] = (its, its_o, reason, float(np.linalg.norm(xg - xo) / np.linalg.norm(xo)), float(mv_err))
    finally:
        dist.destroy_process_group()


def test_distributed_gmres_matches_oracle():
    world, N_x, N_t = 2, 20, 12
    ret = mp.Manager().dict()
    mp.spawn(_gmres_worker, args=(world, _free_port(), N_x, N_t, ret), nprocs=world, join=True)
    for r in range(world):
        its, its_o, reason, err, mv_err = ret[r]
        assert reason == "CONVERGED_RTOL" and abs(its - its_o) <= 1 and err < 1e-6 and mv_err < 1e-13, ret[r]


def test_slab_bounds():
    assert slab_bounds(16385, 8) == ([2049] + [2048] * 7, [0, 2049, 4097, 6145, 8193, 10241, 12289, 14337, 16385])
    c, o = slab_bounds(10, 3)
    assert c == [4, 3, 3] and o == [0, 4, 7, 10]
    c, o = slab_bounds(4096, 8)
    assert c == [512] * 8 and o[-1] == 4096


# ----------------------------------------------------- DiagFFTPC itself on a parallel communicator
def _pc_worker(rank, world, port, N_x, N_t, gamma, ret, how):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optimal_control_paradiag_b200 import DiagFFTPC, petsc_shim
        from oracle.pc_fast import DiagFFTPCFast
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, gamma, **kw)
        opts = petsc_shim.Options()
        if how == "option":          # <prefix>diagfft_distributed, like any other PC-local option
            opts["fd_diagfft_distributed"] = 1
            DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=gamma, backend_factory=factory)
            pc = petsc_shim.PC(prefix="fd_", options=opts)
        else:                        # auto: the PC's communicator and torch.distributed agree on a size > 1
            DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=gamma, backend_factory=factory)
            pc = petsc_shim.PC(comm=petsc_shim.Comm(world, rank))
        pc.setPythonType("optimal_control_paradiag_b200.DiagFFTPC")
        pc.setUp()
        ctx = pc.getPythonContext()
        assert ctx.dpc is not None and ctx.dpc.world == world and ctx.dpc.mode == "slab"
        rng = np.random.default_rng(0)
        size = 2 * (N_x + 1) * N_t
        xg = rng.standard_normal(size) + 1j * rng.standard_normal(size)
        j0, j1 = ctx.dpc.noff[rank], ctx.dpc.noff[rank + 1]
        xl = np.ascontiguousarray(xg.reshape(2, N_x + 1, N_t)[:, j0:j1, :]).reshape(-1)   # this rank's local Vec
        xv, yv = petsc_shim.Vec(xl), petsc_shim.Vec.zeros(xl.size)
        pc.apply(xv, yv)                                                                   # the reference's call
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(xg).reshape(2, N_x + 1, N_t)[:, j0:j1, :].reshape(-1)
        err = float(np.linalg.norm(yv.getArray() - ref) / np.linalg.norm(ref))
        # a Vec of the wrong (global) size is an error, as a PETSc size mismatch would be
        bad = False
        try:
            pc.apply(petsc_shim.Vec(xg), petsc_shim.Vec.zeros(size))
        except ValueError:
            bad = True
        ret[rank] = (err, bad)
        DiagFFTPC._defaults = {}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("how", ["option", "comm"])
def test_diagfftpc_selects_the_distributed_backend_on_a_parallel_communicator(how):
    world, N_x, N_t = 2, 21, 12
    ret = mp.Manager().dict()
    mp.spawn(_pc_worker, args=(world, _free_port(), N_x, N_t, 0.5, ret, how), nprocs=world, join=True)
    for r in range(world):
        err, bad = ret[r]
        assert err < 1e-11 and bad, ret[r]
