"""Multi-GPU apply of the ParaDiag preconditioner: one process per GPU (torch.distributed).

The reference has no parallel decomposition at all (its "parallel-in-time" is algebraic:
all frequencies go into one monolithic MUMPS factorisation, Control_Wave_PC.py:481-484).
The path shards naturally with ONE exchange step each way:

    stage 1  time-axis inverse FFT        sharded by NODE slab   (rank r: nodes [j0_r, j1_r), all N_t)
    ---- all-to-all: (2, n_r, N_t) -> (2, n, k_r) ----
    stage 2  per-frequency solves          sharded by FREQUENCY   (rank r: k in [k0_r, k1_r), all nodes)
    ---- all-to-all: (2, n, k_r) -> (2, n_r, N_t) ----
    stage 3  time-axis forward FFT         sharded by node slab

Vectors are distributed by node slab -- the layout a PETSc/Firedrake spatial decomposition of
the reference would give: rank r holds x[:, j0_r:j1_r, :] as a contiguous (2, n_r, N_t) block.

The compute backend is a ``ParaDiagHandle`` created with this rank's frequency slab
(``pd_stage_fft`` / ``pd_stage_solve`` of include/paradiag.h).  The transposes use
``all_to_all_single`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) with uneven splits,
so neither n = N_x + 1 nor N_t has to be divisible by the world size.

``mode="slab"`` keeps the NODE-slab sharding through stage 2 instead: the partition method that
already solves each x-line in chunks is extended across ranks -- every rank eliminates its slab
(``pd_slab_reduce``), the first/last entries of the slab-local solves (6 N_t complex numbers per rank)
are all-gathered, every rank solves the (G-1)-row separator system of each frequency redundantly
and back-substitutes its slab (``pd_slab_finish``).  Communication drops from 2 x (G-1)/G x S/G bytes
per rank (NVLink-bound) to 96 N_t bytes per rank; no transposes, no pack/unpack.
"""
import math
import os

import numpy as np


def slab_bounds(total, parts):
    """Balanced contiguous split: the first ``total % parts`` slabs get one extra item."""
    base, extra = divmod(total, parts)
    counts = [base + (1 if r < extra else 0) for r in range(parts)]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    return counts, offs


class DistributedDiagFFTPC:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, device=0, group=None, backend_factory=None,
                 mode="alltoall", transport=None, alpha=1.0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.N_x, self.N_t, self.n = int(N_x), int(N_t), int(N_x) + 1
        self.ncount, self.noff = slab_bounds(self.n, self.world)
        self.kcount, self.koff = slab_bounds(self.N_t, self.world)
        if min(self.kcount) < 1 or min(self.ncount) < 1:
            raise ValueError(f"world size {self.world} exceeds N_t = {N_t} or the node count {self.n}")
        self.n_r, self.k_r = self.ncount[self.rank], self.kcount[self.rank]
        if backend_factory is None:
            from .handle import ParaDiagHandle

            def backend_factory(**kw):
                return ParaDiagHandle(N_x, N_t, T=T, gamma=gamma, alpha=alpha, device=device, **kw)
            self.device = torch.device(f"cuda:{device}")
        else:
            self.device = torch.device("cpu")
        if mode not in ("alltoall", "slab"):
            raise ValueError(f"unknown mode {mode!r}")
        self.alpha = float(alpha)
        if self.alpha != 1.0 and mode != "slab":
            raise NotImplementedError("alpha != 1 (an extension) is available in slab mode only")
        self.mode = mode
        c128 = torch.complex128
        self.local_size = 2 * self.n_r * self.N_t
        if mode == "slab":
            self.backend = backend_factory(slab_rank=self.rank, slab_count=self.world)
            self.comm_bytes_per_apply = 16 * 6 * self.N_t
            # transport of the slab functionals: "peer" = stores into every rank's IPC-mapped exchange buffer
            # from inside the producing kernel (pd_slab_apply, no collective call on the data path);
            # "nccl" = one all_gather_into_tensor between pd_slab_reduce and pd_slab_finish (also what the
            # CPU test backends use)
            self.transport = "nccl"
            want = transport or os.environ.get("PD_SLAB_TRANSPORT", "peer")
            if want == "peer" and hasattr(self.backend, "slab_comm_create") and self.device.type == "cuda":
                self.transport = self._connect_peers()
            if self.transport == "nccl":
                self._alloc_nccl_buffers()
            return
        self.backend = backend_factory(k_begin=self.koff[self.rank], k_count=self.k_r, n_local=self.n_r)
        self.freq_size = 2 * self.n * self.k_r
        # work buffers: time-domain slab, frequency-domain slab, send / receive staging
        self.w_time = torch.empty(self.local_size, dtype=c128, device=self.device)
        self.w_freq = torch.empty(self.freq_size, dtype=c128, device=self.device)
        self.sendbuf = torch.empty(max(self.local_size, self.freq_size), dtype=c128, device=self.device)
        self.recvbuf = torch.empty(max(self.local_size, self.freq_size), dtype=c128, device=self.device)
        # split sizes in complex elements
        self.a_send = [2 * self.n_r * k for k in self.kcount]      # to rank s: (2, n_r, k_s)
        self.a_recv = [2 * ns * self.k_r for ns in self.ncount]    # from rank s: (2, n_s, k_r)
        self.comm_bytes_per_apply = 16 * 2 * (sum(self.a_send) - self.a_send[self.rank])

    def _alloc_nccl_buffers(self):
        t, c128 = self.torch, self.torch.complex128
        self.w_time = t.empty(self.local_size, dtype=c128, device=self.device)
        self.fl_out = t.empty(6 * self.N_t, dtype=c128, device=self.device)
        self.gathered = t.empty(self.world * 6 * self.N_t, dtype=c128, device=self.device)

    def _connect_peers(self):
        """Exchange the IPC handles of the ranks' exchange buffers and map them; "peer" on success on EVERY rank,
        else "nccl" everywhere (the decision is collective)."""
        ok, handles = 1, None
        try:
            mine, _ = self.backend.slab_comm_create()
            handles = [None] * self.world
            self.dist.all_gather_object(handles, mine, group=self.group)
            self.backend.slab_comm_connect_ipc(handles)
        except Exception as ex:  # no IPC in this environment: every rank falls back together
            ok, self.transport_error = 0, str(ex)
        flag = self.torch.tensor([ok], dtype=self.torch.int32, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        return "peer" if int(flag.item()) == 1 else "nccl"

    def check_exchange(self):
        """Raises if a bounded wait of the peer-store exchange expired (a rank never delivered)."""
        if getattr(self, "transport", None) == "peer":
            timed_out, epoch = self.backend.slab_comm_status()
            if timed_out:
                raise RuntimeError(f"rank {self.rank}: the slab exchange timed out waiting for a peer (epoch {epoch})")

    def apply_profile(self, x_local, y_local):
        """Per-stage device times of one distributed apply (peer transport): dict of milliseconds."""
        if getattr(self, "transport", None) != "peer":
            raise NotImplementedError("per-stage timing needs the peer-store transport")
        return self.backend.slab_apply_profile(x_local.reshape(-1), y_local.reshape(-1))

    # -------------------------------------------------------------------------------
    @property
    def launch_count(self):
        return getattr(self.backend, "launch_count", 0)

    def describe(self):
        if self.mode == "slab":
            return {"world": self.world, "mode": "slab", "node_slabs": self.ncount, "transport": self.transport,
                    "exchange_bytes_sent_per_rank_per_apply": self.comm_bytes_per_apply * (
                        self.world - 1 if self.transport == "peer" else 1),
                    "collective": ("none on the data path: the functionals kernel stores 6 N_t complex values into "
                                   "every peer's IPC-mapped buffer and the separator kernel waits on per-block flags")
                    if self.transport == "peer" else
                    "one all_gather of 6 N_t complex values per rank (slab functionals)"}
        return {"world": self.world, "mode": "alltoall", "node_slabs": self.ncount, "freq_slabs": self.kcount,
                "alltoall_bytes_sent_per_rank_per_apply": self.comm_bytes_per_apply,
                "collective": "all_to_all_single x2 (uneven splits), pack/unpack by strided copies"}

    def random_local(self, seed=0):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(self.local_size) + 1j * rng.standard_normal(self.local_size)
        return self.torch.tensor(x, device=self.device)

    def scatter_from_global(self, x_global):
        """This rank's (2, n_r, N_t) block of a replicated global vector (tests / set-up only)."""
        xg = x_global.reshape(2, self.n, self.N_t)
        j0, j1 = self.noff[self.rank], self.noff[self.rank + 1]
        return xg[:, j0:j1, :].reshape(-1).contiguous().to(self.device)

    def _a2a(self, out, inp, out_split, in_split):
        # complex128 travels as pairs of float64 (NCCL has no complex type)
        t = self.torch
        self.dist.all_to_all_single(t.view_as_real(out).reshape(-1), t.view_as_real(inp).reshape(-1),
                                    [2 * s for s in out_split], [2 * s for s in in_split], group=self.group)

    # -------------------------------------------------------------------------------
    def apply(self, x_local, y_local=None):
        """y = P^-1 x for node-slab distributed vectors (DiagFFTPC.apply, Control_Wave_PC.py:491-553)."""
        t = self.torch
        if y_local is None:
            y_local = t.empty_like(x_local)
        if self.mode == "slab":
            return self._apply_slab(x_local, y_local)
        n_r, k_r, n, N_t, G = self.n_r, self.k_r, self.n, self.N_t, self.world
        # stage 1: inverse FFT along time of this rank's 2 n_r lines (:500-501)
        self.backend.stage_fft(x_local.reshape(-1), self.w_time, 2 * n_r, True)
        wt = self.w_time.view(2, n_r, N_t)
        # pack: block for rank s = (2, n_r, k_s)
        off = 0
        for s in range(G):
            k0, k1 = self.koff[s], self.koff[s + 1]
            self.sendbuf[off:off + self.a_send[s]].view(2, n_r, k1 - k0).copy_(wt[:, :, k0:k1])
            off += self.a_send[s]
        self._a2a(self.recvbuf[:self.freq_size], self.sendbuf[:self.local_size], self.a_recv, self.a_send)
        # unpack: block from rank s = (2, n_s, k_r) -> rows [j0_s, j1_s) of (2, n, k_r)
        wf = self.w_freq.view(2, n, k_r)
        off = 0
        for s in range(G):
            j0, j1 = self.noff[s], self.noff[s + 1]
            wf[:, j0:j1, :].copy_(self.recvbuf[off:off + self.a_recv[s]].view(2, j1 - j0, k_r))
            off += self.a_recv[s]
        # stage 2: this rank's frequencies, all nodes (:445-540)
        self.backend.stage_solve(self.w_freq)
        # pack back: block for rank s = (2, n_s, k_r)
        off = 0
        for s in range(G):
            j0, j1 = self.noff[s], self.noff[s + 1]
            self.sendbuf[off:off + self.a_recv[s]].view(2, j1 - j0, k_r).copy_(wf[:, j0:j1, :])
            off += self.a_recv[s]
        self._a2a(self.recvbuf[:self.local_size], self.sendbuf[:self.freq_size], self.a_send, self.a_recv)
        off = 0
        for s in range(G):
            k0, k1 = self.koff[s], self.koff[s + 1]
            wt[:, :, k0:k1].copy_(self.recvbuf[off:off + self.a_send[s]].view(2, n_r, k1 - k0))
            off += self.a_send[s]
        # stage 3: forward FFT along time (:547-548)
        self.backend.stage_fft(self.w_time, y_local.reshape(-1), 2 * n_r, False)
        return y_local

    def _apply_slab(self, x_local, y_local):
        t = self.torch
        if self.transport == "peer":
            return self.backend.slab_apply(x_local.reshape(-1), y_local.reshape(-1))          # :491-553
        if self.alpha != 1.0:
            # (collective transport + the alpha extension: Gamma / Gamma^-1 as plain elementwise products here; the
            # peer-store transport has them inside the transforms)
            x_local = (x_local.reshape(-1, self.N_t) * self._gamma(False)).reshape(-1)
        self.backend.stage_fft(x_local.reshape(-1), self.w_time, 2 * self.n_r, True)      # :500-501
        self.backend.slab_reduce(self.w_time, self.fl_out)
        self.dist.all_gather_into_tensor(t.view_as_real(self.gathered).reshape(-1),
                                         t.view_as_real(self.fl_out).reshape(-1), group=self.group)
        self.backend.slab_finish(self.w_time, self.gathered)                               # :445-540
        self.backend.stage_fft(self.w_time, y_local.reshape(-1), 2 * self.n_r, False)     # :547-548
        if self.alpha != 1.0:
            y_local.reshape(-1, self.N_t).mul_(self._gamma(True))
        return y_local

    def _gamma(self, inverse):
        """Gamma_alpha time weights a^j (inverse: a^-j), a = alpha^(1/N_t), as a float64 device vector."""
        t = self.torch
        if getattr(self, "_gam", None) is None:
            j = t.arange(self.N_t, dtype=t.float64, device=self.device) / self.N_t
            lna = float(np.log(self.alpha))
            self._gam = (t.exp(lna * j), t.exp(-lna * j))
        return self._gam[1 if inverse else 0]

    def apply_real(self, x_local, y_local=None):
        """The same apply for REAL node-slab blocks (float64, (2, n_r, N_t)): what GMRES feeds the PC in this
        real problem.  Half spectrum (k <= N_t/2) in every stage: half the bytes of ``apply``, and the
        all-gather shrinks to 6 (N_t/2 + 1) values per rank.  Slab mode, N_t >= 8."""
        t = self.torch
        if self.mode != "slab":
            raise NotImplementedError("the real-input distributed apply uses the slab decomposition")
        if y_local is None:
            y_local = t.empty_like(x_local)
        if self.transport == "peer":
            return self.backend.slab_apply(x_local.reshape(-1), y_local.reshape(-1), real=True)
        if getattr(self, "_w_half", None) is None:
            if not hasattr(self, "w_time"):
                self._alloc_nccl_buffers()
            Kp = self.backend.half_cols
            c128 = t.complex128
            self._w_half = t.empty(2 * self.n_r * Kp, dtype=c128, device=self.device)
            self._fl_half = t.empty(6 * Kp, dtype=c128, device=self.device)
            self._gath_half = t.empty(self.world * 6 * Kp, dtype=c128, device=self.device)
        if self.alpha != 1.0:
            x_local = (x_local.reshape(-1, self.N_t) * self._gamma(False)).reshape(-1)
        self.backend.stage_rfft_pair(x_local.reshape(-1), self._w_half, self.n_r, True)
        self.backend.slab_reduce_half(self._w_half, self._fl_half)
        self.dist.all_gather_into_tensor(t.view_as_real(self._gath_half).reshape(-1),
                                         t.view_as_real(self._fl_half).reshape(-1), group=self.group)
        self.backend.slab_finish_half(self._w_half, self._gath_half)
        self.backend.stage_rfft_pair(self._w_half, y_local.reshape(-1), self.n_r, False)
        if self.alpha != 1.0:
            y_local.reshape(-1, self.N_t).mul_(self._gamma(True))
        return y_local

    def apply_host(self, x_host, y_host):
        """The same apply for node-slab blocks that live in HOST memory (what a PETSc ``Vec`` of a spatial
        decomposition hands the PC, :493-497 / :552-553): H2D of this rank's block, apply, D2H.
        ``x_host`` / ``y_host``: complex128 torch CPU tensors (pinned for full PCIe rate) or numpy arrays."""
        t = self.torch
        if isinstance(x_host, np.ndarray):
            x_host = t.from_numpy(x_host)
        if isinstance(y_host, np.ndarray):
            y_host = t.from_numpy(y_host)
        if getattr(self, "_xdev", None) is None:
            self._xdev = t.empty(self.local_size, dtype=t.complex128, device=self.device)
            self._ydev = t.empty(self.local_size, dtype=t.complex128, device=self.device)
        self._xdev.copy_(x_host.reshape(-1), non_blocking=True)
        self.apply(self._xdev, self._ydev)
        y_host.reshape(-1).copy_(self._ydev, non_blocking=True)
        if self.device.type == "cuda":
            t.cuda.current_stream(self.device).synchronize()
        return y_host

    # ------------------------------------------------------------------ distributed Krylov solve
    def build_rhs(self, real=False):
        """This rank's block of the manufactured right-hand side (Build_f/g/IC, :48-83); ``real``: float64."""
        if real:
            b = self.torch.empty(self.local_size, dtype=self.torch.float64, device=self.device)
            return self.backend.build_rhs_real(b)
        b = self.torch.empty(self.local_size, dtype=self.torch.complex128, device=self.device)
        return self.backend.build_rhs(b)

    def matvec(self, x_local, y_local=None):
        """y = A x (Build_L, :86-179) on node-slab distributed vectors: one small all-gather brings the
        neighbours' edge rows (2 x 2 x N_t complex per rank), then the local stencil kernel runs."""
        t = self.torch
        if self.mode != "slab":
            raise NotImplementedError("the distributed matvec / GMRES use the slab decomposition")
        if y_local is None:
            y_local = t.empty_like(x_local)
        X = x_local.view(2, self.n_r, self.N_t)
        edges = t.stack([X[:, 0, :], X[:, -1, :]]).contiguous()            # (edge, field, N_t)
        alle = t.empty((self.world,) + tuple(edges.shape), dtype=edges.dtype, device=self.device)
        self.dist.all_gather_into_tensor(t.view_as_real(alle).reshape(-1), t.view_as_real(edges).reshape(-1),
                                         group=self.group)
        lo = alle[self.rank - 1, 1].reshape(-1) if self.rank > 0 else None          # last row of the left slab
        hi = alle[self.rank + 1, 0].reshape(-1) if self.rank < self.world - 1 else None
        self._halo_keepalive = alle
        return self.backend.matvec_slab(x_local.reshape(-1), lo, hi, y_local.reshape(-1))

    def matvec_real(self, x_local, y_local=None):
        """``matvec`` for float64 node-slab blocks (the real problem): same halo exchange on float64 rows."""
        t = self.torch
        if self.mode != "slab":
            raise NotImplementedError("the distributed matvec / GMRES use the slab decomposition")
        if y_local is None:
            y_local = t.empty_like(x_local)
        X = x_local.view(2, self.n_r, self.N_t)
        edges = t.stack([X[:, 0, :], X[:, -1, :]]).contiguous()
        alle = t.empty((self.world,) + tuple(edges.shape), dtype=edges.dtype, device=self.device)
        self.dist.all_gather_into_tensor(alle.reshape(-1), edges.reshape(-1), group=self.group)
        lo = alle[self.rank - 1, 1].reshape(-1) if self.rank > 0 else None
        hi = alle[self.rank + 1, 0].reshape(-1) if self.rank < self.world - 1 else None
        self._halo_keepalive = alle
        return self.backend.matvec_slab_real(x_local.reshape(-1), lo, hi, y_local.reshape(-1))

    def _allreduce(self, v):
        self.dist.all_reduce(self.torch.view_as_real(v), group=self.group)
        return v

    def gmres(self, b_local, rtol=1e-7, atol=1e-50, restart=300, max_it=1000, real=None):
        """Left-preconditioned GMRES with the options of Control_Wave_PC.py:347-359 (classical
        Gram-Schmidt, zero initial guess, preconditioned-residual test) on slab-distributed vectors.
        Same arithmetic as pd_gmres; inner products are local pd_mdot + one all-reduce.
        ``real`` (default: by the dtype of ``b_local``): float64 vectors, the half-spectrum preconditioner
        ``apply_real`` and the float64 matvec -- half the bytes in every sweep (the problem is real).
        Returns (x_local, iterations, history, reason)."""
        t, be = self.torch, self.backend
        if real is None:
            real = b_local.dtype == t.float64
        c128, ln = t.complex128, self.local_size
        vdt = t.float64 if real else c128
        esz = 8 if real else 16
        x = t.zeros(ln, dtype=vdt, device=self.device)
        tmp = t.empty(ln, dtype=vdt, device=self.device)
        apply_fn = self.apply_real if real else self.apply
        matvec_fn = self.matvec_real if real else self.matvec
        # the BLAS-1 kernels see float64 vectors as complex pairs (only the real part of the sums is kept)
        cv = (lambda v: t.view_as_complex(v.reshape(*v.shape[:-1], v.shape[-1] // 2, 2))) if real else (lambda v: v)
        if real and hasattr(be, "set_option"):
            be.set_option("krylov_real_vectors", 1)
        # Krylov basis in blocks of BS vectors, allocated as the iteration proceeds (a cfg3 vector is
        # 2.1 GB / G per rank: the restart length of 300 cannot be pre-allocated)
        BS = max(1, min(8, (2 << 30) // (esz * ln)))      # ~2 GB per block
        blocks = [t.empty((BS, ln), dtype=vdt, device=self.device)]

        def vec(j):
            while j // BS >= len(blocks):
                blocks.append(t.empty((BS, ln), dtype=vdt, device=self.device))
            return blocks[j // BS][j % BS]

        def mdot_all(nv, w):
            return t.cat([be.mdot(cv(blocks[b][: min(BS, nv - b * BS)]), cv(w)) for b in range((nv + BS - 1) // BS)])

        def maxpy_all(nv, coef, sign, w, norm2_out=None):
            nb = (nv + BS - 1) // BS
            for b in range(nb):
                cnt = min(BS, nv - b * BS)
                be.maxpy(cv(blocks[b][:cnt]), coef[b * BS: b * BS + cnt], sign, cv(w), norm2_out if b == nb - 1 else None)
        try:
            return self._gmres_loop(b_local, x, tmp, vec, mdot_all, maxpy_all, apply_fn, matvec_fn, rtol, atol, restart,
                                    max_it)
        finally:
            if real and hasattr(be, "set_option"):
                be.set_option("krylov_real_vectors", 0)

    def _gmres_loop(self, b_local, x, tmp, vec, mdot_all, maxpy_all, apply_fn, matvec_fn, rtol, atol, restart, max_it):
        t, c128 = self.torch, self.torch.complex128
        hist, its, reason, first, converged = [], 0, "DIVERGED_ITS", True, False
        beta0 = target = 0.0
        hbuf = t.zeros(restart + 2, dtype=c128, device=self.device)
        from ._lib import Hessenberg
        hess = Hessenberg(restart)
        while not converged and (its < max_it or first):
            v0 = vec(0)
            if first:
                apply_fn(b_local, v0)
            else:
                matvec_fn(x, tmp)
                tmp.mul_(-1).add_(b_local)
                apply_fn(tmp, v0)
            beta = math.sqrt(self._allreduce(mdot_all(1, v0))[0].real.item())
            if first:
                beta0, first = beta, False
                target = max(rtol * beta0, atol)
                hist.append(beta0)
                if beta0 <= target or beta0 == 0.0:
                    converged, reason = True, ("CONVERGED_ATOL" if beta0 <= atol else "CONVERGED_RTOL")
                    break
                if max_it == 0:
                    break
            v0.mul_(1.0 / beta)
            m = restart
            hess.start(beta)
            jdone = 0
            for j in range(m):
                w = vec(j + 1)
                matvec_fn(vec(j), tmp)
                apply_fn(tmp, w)
                hd = self._allreduce(mdot_all(j + 1, w))                     # classical Gram-Schmidt
                hbuf[: j + 1] = hd
                maxpy_all(j + 1, hbuf, -1.0, w, hbuf[j + 1: j + 2])
                self._allreduce(hbuf[j + 1: j + 2])
                # Hessenberg column, Givens rotations, residual estimate: the library's one implementation
                rn, hn = hess.push(hbuf[: j + 2].cpu().numpy())
                its += 1
                jdone = j + 1
                hist.append(rn)
                if rn <= target:
                    converged, reason = True, ("CONVERGED_RTOL" if rn > atol else "CONVERGED_ATOL")
                    break
                if its >= max_it or hn == 0.0:
                    break
                w.mul_(1.0 / hn)
            if jdone:
                coef = t.tensor(hess.solve(), dtype=c128, device=self.device)
                maxpy_all(jdone, coef, 1.0, x)
            if its >= max_it:
                break
        return x, its, hist, reason

    def gather_to_global(self, y_local):
        """All ranks receive the full (2, n, N_t) vector (tests only)."""
        t = self.torch
        parts = [t.empty(2 * c * self.N_t, dtype=t.complex128, device=self.device) for c in self.ncount]
        reals = [t.view_as_real(p) for p in parts]
        self.dist.all_gather(reals, t.view_as_real(y_local.reshape(-1).contiguous()), group=self.group) \
            if len(set(self.ncount)) == 1 else self._uneven_gather(reals, y_local)
        full = t.empty(2, self.n, self.N_t, dtype=t.complex128, device=self.device)
        for s, p in enumerate(parts):
            full[:, self.noff[s]:self.noff[s + 1], :] = p.view(2, self.ncount[s], self.N_t)
        return full.reshape(-1)

    def _uneven_gather(self, reals, y_local):
        t = self.torch
        mine = t.view_as_real(y_local.reshape(-1).contiguous())
        for s in range(self.world):
            if s == self.rank:
                reals[s].copy_(mine)
            self.dist.broadcast(reals[s], src=self.dist.get_global_rank(self.group, s) if self.group else s,
                                group=self.group)


class LocalSlabGroup:
    """``G`` x-slab handles driven by ONE process: the peer-store exchange of ``pd_slab_apply`` with plain device
    pointers instead of IPC handles.  All slabs may sit on one GPU (how the distributed path is exercised where
    only one GPU is leased: every first half ``pd_slab_apply_begin`` is issued before any second half, so no
    kernel ever waits for one queued behind it) or on several GPUs with peer access (one per slab)."""

    def __init__(self, N_x, N_t, G, T=2.0, gamma=1.0, devices=None, split_on_one_stream=False, alpha=1.0):
        import torch

        from .handle import ParaDiagHandle
        self.torch = torch
        self.N_x, self.N_t, self.n, self.G = int(N_x), int(N_t), int(N_x) + 1, int(G)
        self.devices = [0] * self.G if devices is None else [int(d) for d in devices]
        self.ncount, self.noff = slab_bounds(self.n, self.G)
        self.handles = [ParaDiagHandle(N_x, N_t, T=T, gamma=gamma, alpha=alpha, device=self.devices[r], slab_rank=r,
                                       slab_count=G)
                        for r in range(self.G)]
        bases = [h.slab_comm_create()[1] for h in self.handles]
        for h in self.handles:
            h.slab_comm_connect_local(bases, self.devices)
            # several ranks on one GPU: everything stays on ONE stream, so that every first half really runs before
            # any second half (a second stream per handle would let a waiting kernel overtake its producer)
            if len(set(self.devices)) < self.G:
                # (2: still split into the two frequency halves, one after the other -- covers the range kernels)
                h.set_option("slab_overlap", 2 if split_on_one_stream else 0)

    def close(self):
        for h in self.handles:
            h.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def scatter(self, x_global):
        xg = x_global.reshape(2, self.n, self.N_t)
        return [xg[:, self.noff[r]:self.noff[r + 1], :].reshape(-1).contiguous().to(f"cuda:{self.devices[r]}")
                for r in range(self.G)]

    def apply_blocks(self, xs, real=False):
        """One distributed apply on the ranks' blocks; returns the output blocks (same devices)."""
        t = self.torch
        ys = [t.empty_like(x) for x in xs]
        if len(set(self.devices)) > 1:                      # the producers' inputs must be complete device-side
            for d in set(self.devices):
                t.cuda.synchronize(d)
        for h, x in zip(self.handles, xs):
            h.slab_apply_begin(x, real=real)
        for h, y in zip(self.handles, ys):
            h.slab_apply_end(y, real=real)
        return ys

    def apply(self, x_global, real=False):
        ys = self.apply_blocks(self.scatter(x_global), real=real)
        dev = x_global.device
        parts = [y.to(dev).view(2, self.ncount[r], self.N_t) for r, y in enumerate(ys)]
        return self.torch.cat(parts, dim=1).reshape(-1)

    def status(self):
        return [h.slab_comm_status() for h in self.handles]
