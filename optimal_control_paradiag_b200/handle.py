"""Python wrapper of one ``pd_handle`` (a problem size bound to one CUDA device).

Device vectors are torch tensors (complex128, contiguous, on the handle's device); the
wrapper only extracts ``data_ptr()`` and the current CUDA stream -- torch is plumbing
for memory and streams, all arithmetic happens inside libparadiag.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, pd_config

CONVERGED_REASONS = {2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", -3: "DIVERGED_ITS"}


def _torch():
    import torch
    return torch


class ParaDiagHandle:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, alpha=1.0, bug138=True, device=0,
                 k_begin=0, k_count=0, n_local=0, slab_rank=0, slab_count=0):
        self._h = C.c_void_p()
        self.lib = _lib.load_library()
        self.N_x, self.N_t, self.T, self.gamma = int(N_x), int(N_t), float(T), float(gamma)
        self.n = self.N_x + 1
        self.device = int(device)
        self.alpha = float(alpha)
        self.size = 2 * self.n * self.N_t
        self.k_count = int(k_count) if k_count else self.N_t
        self.k_begin = int(k_begin) if k_count else 0
        self.n_local = int(n_local) if n_local else self.n
        # local node rows of an x-slab handle (balanced split, the first n % G slabs one node longer)
        if slab_count and int(slab_count) > 1:
            G, r = int(slab_count), int(slab_rank)
            self.n_slab = self.n // G + (1 if r < self.n % G else 0)
        else:
            self.n_slab = self.n
        cfg = pd_config(abi_version=_lib.PD_ABI_VERSION, N_x=self.N_x, N_t=self.N_t, bug138=int(bool(bug138)),
                        T=self.T, gamma=self.gamma, alpha=float(alpha), device=self.device,
                        k_begin=int(k_begin), k_count=int(k_count), n_local=int(n_local),
                        slab_rank=int(slab_rank), slab_count=int(slab_count))
        check(self.lib.pd_create(C.byref(cfg), C.byref(self._h)))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def workspace_bytes(self):
        return int(self.lib.pd_workspace_bytes(self._h))

    @property
    def real_path_supported(self):
        """True when pd_pc_apply_real exists for this handle: every N_t >= 8 (register pipelines for the powers of
        two in [128, 16384], the shared-memory pair kernel for the rest -- the upstream default N_t = 81 included)."""
        return self.N_t >= 8

    @property
    def launch_count(self):
        return int(self.lib.pd_launch_count(self._h))

    # ------------------------------------------------------------------- helpers
    def _ptr(self, t, numel=None, name="vector"):
        torch = _torch()
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch tensor, got {type(t)}")
        if t.dtype != torch.complex128 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name}: need a contiguous complex128 CUDA tensor")
        if t.device.index != self.device:
            raise ValueError(f"{name}: tensor on cuda:{t.device.index}, handle on cuda:{self.device}")
        if numel is not None and t.numel() != numel:
            raise ValueError(f"{name}: expected {numel} entries, got {t.numel()}")
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def empty(self):
        torch = _torch()
        return torch.empty(self.size, dtype=torch.complex128, device=f"cuda:{self.device}")

    # -------------------------------------------------------------- entry points
    def pc_apply(self, x, y=None):
        """y = P^-1 x on device tensors (DiagFFTPC.apply, Control_Wave_PC.py:491-553)."""
        if y is None:
            y = self.empty()
        check(self.lib.pd_pc_apply(self._h, self._ptr(x, self.size, "x"), self._ptr(y, self.size, "y"),
                                   self._stream()))
        return y

    def pc_apply_real(self, x, y=None):
        """Real-input fast path (pd_pc_apply_real): x, y float64 CUDA tensors of 2 n N_t entries."""
        torch = _torch()
        if y is None:
            y = torch.empty_like(x)
        for name, v in (("x", x), ("y", y)):
            if v.dtype != torch.float64 or not v.is_cuda or not v.is_contiguous() or v.numel() != self.size:
                raise ValueError(f"{name}: need a contiguous float64 CUDA tensor with {self.size} entries")
        check(self.lib.pd_pc_apply_real(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), self._stream()))
        return y

    def stage_rfft(self, src, dst, nlines, to_freq):
        check(self.lib.pd_stage_rfft(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), int(nlines),
                                     int(bool(to_freq)), self._stream()))
        return dst

    def stage_solve_half(self, w):
        check(self.lib.pd_stage_solve_half(self._h, C.c_void_p(w.data_ptr()), self._stream()))
        return w

    def pc_apply_profile(self, x, y):
        """One apply with per-kernel CUDA-event timing: dict of milliseconds."""
        ms = (C.c_float * 5)()
        check(self.lib.pd_pc_apply_profile(self._h, self._ptr(x, self.size, "x"), self._ptr(y, self.size, "y"),
                                           self._stream(), ms, 5))
        return dict(zip(("ifft", "passA", "interface", "passB", "fft"), [float(v) for v in ms]))

    def pc_apply_host(self, x, y=None):
        """Same through host buffers (numpy complex128): H2D, apply, D2H."""
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        if x.size != self.size:
            raise ValueError(f"x: expected {self.size} entries, got {x.size}")
        if y is None:
            y = np.empty(self.size, dtype=np.complex128)
        if y.dtype != np.complex128 or not y.flags.c_contiguous or y.size != self.size:
            raise ValueError("y: need a contiguous complex128 array of the vector size")
        check(self.lib.pd_pc_apply_host(self._h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
        return y

    def pc_apply_real_host(self, x, y=None):
        """Real-input path through host buffers (numpy float64): H2D, pd_pc_apply_real, D2H -- half the bytes."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if x.size != self.size:
            raise ValueError(f"x: expected {self.size} entries, got {x.size}")
        if y is None:
            y = np.empty(self.size, dtype=np.float64)
        if y.dtype != np.float64 or not y.flags.c_contiguous or y.size != self.size:
            raise ValueError("y: need a contiguous float64 array of the vector size")
        check(self.lib.pd_pc_apply_real_host(self._h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
        return y

    def host_unregister_all(self):
        """Drop the page-lock registrations pc_apply_host made for host buffers (pd_host_unregister_all)."""
        check(self.lib.pd_host_unregister_all(self._h))

    def pc_apply_transpose(self, x, y):
        st = self.lib.pd_pc_apply_transpose(self._h, None, None, None)
        if st == _lib.PD_ERR_UNSUPPORTED:
            raise NotImplementedError
        check(st)

    def stage_fft(self, src, dst, nlines, inverse):
        check(self.lib.pd_stage_fft(self._h, self._ptr(src, nlines * self.N_t, "src"),
                                    self._ptr(dst, nlines * self.N_t, "dst"), int(nlines), int(bool(inverse)),
                                    self._stream()))
        return dst

    def stage_gamma(self, src, dst, nlines, inverse):
        """Gamma_alpha (inverse False) / Gamma_alpha^-1 (inverse True) time weights; alpha != 1 only."""
        check(self.lib.pd_stage_gamma(self._h, self._ptr(src, nlines * self.N_t, "src"),
                                      self._ptr(dst, nlines * self.N_t, "dst"), int(nlines), int(bool(inverse)),
                                      self._stream()))
        return dst

    def stage_solve(self, w):
        check(self.lib.pd_stage_solve(self._h, self._ptr(w, 2 * self.n * self.k_count, "w"), self._stream()))
        return w

    def slab_reduce(self, w, out):
        """Slab mode, first half (pd_slab_reduce): out (6, N_t) <- slab functionals."""
        check(self.lib.pd_slab_reduce(self._h, self._ptr(w, None, "w"), self._ptr(out, 6 * self.N_t, "out"),
                                      self._stream()))
        return out

    def slab_finish(self, w, gathered):
        """Slab mode, second half (pd_slab_finish): separator solve + back-substitution in place."""
        check(self.lib.pd_slab_finish(self._h, self._ptr(w, None, "w"), self._ptr(gathered, None, "gathered"),
                                      self._stream()))
        return w

    # ---- slab mode with the peer-store exchange (the whole distributed apply behind one call)
    def slab_comm_create(self):
        """Allocate this rank's exchange buffer; returns (ipc_handle: 64 bytes, device address)."""
        buf = (C.c_ubyte * 64)()
        base = C.c_void_p()
        check(self.lib.pd_slab_comm_create(self._h, buf, C.byref(base)))
        return bytes(buf), int(base.value)

    def slab_comm_connect_ipc(self, handles):
        """``handles``: the 64-byte IPC handles of every rank (rank order), from other processes."""
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self.lib.pd_slab_comm_connect(self._h, arr, 0, None))

    def slab_comm_connect_local(self, bases, devices=None):
        """``bases``: device addresses of every rank's buffer, valid in this process (one process, many handles)."""
        ptrs = (C.c_void_p * len(bases))(*bases)
        devs = (C.c_int * len(bases))(*devices) if devices is not None else None
        check(self.lib.pd_slab_comm_connect(self._h, ptrs, 1, devs))

    def slab_comm_status(self):
        """(timed_out, epoch): a bounded wait expired since the last call / applies completed."""
        to, ep = C.c_int(0), C.c_uint64(0)
        check(self.lib.pd_slab_comm_status(self._h, C.byref(to), C.byref(ep)))
        return bool(to.value), int(ep.value)

    def _slab_ptr(self, t, real):
        torch = _torch()
        want = torch.float64 if real else torch.complex128
        if t.dtype != want or not t.is_cuda or not t.is_contiguous() or t.numel() != 2 * self.n_slab * self.N_t:
            raise ValueError(f"slab apply: need a contiguous {want} CUDA tensor with {2 * self.n_slab * self.N_t} entries")
        return C.c_void_p(t.data_ptr())

    def slab_apply(self, x, y, real=False):
        fn = self.lib.pd_slab_apply_real if real else self.lib.pd_slab_apply
        check(fn(self._h, self._slab_ptr(x, real), self._slab_ptr(y, real), self._stream()))
        return y

    def slab_apply_begin(self, x, real=False):
        check(self.lib.pd_slab_apply_begin(self._h, self._slab_ptr(x, real), self._stream(), int(real)))

    def slab_apply_end(self, y, real=False):
        check(self.lib.pd_slab_apply_end(self._h, self._slab_ptr(y, real), self._stream(), int(real)))
        return y

    def slab_apply_profile(self, x, y):
        ms = (C.c_float * 7)()
        check(self.lib.pd_slab_apply_profile(self._h, self._slab_ptr(x, False), self._slab_ptr(y, False),
                                             self._stream(), ms, 7))
        return dict(zip(("ifft", "passA", "interface", "functionals_push", "wait_separators", "passB", "fft"),
                        [float(v) for v in ms]))

    # ---- slab mode on the half spectrum of the real-input path
    @property
    def half_cols(self):
        """Row length Kp of a half spectrum: N_t/2 + 1 rounded up to a multiple of 8 complex numbers."""
        return (self.N_t // 2 + 1 + 7) & ~7

    def stage_rfft_pair(self, src, dst, nnodes, to_freq):
        """(2, nnodes, N_t) float64 <-> (2, nnodes, Kp) complex half spectra (pd_stage_rfft_pair)."""
        check(self.lib.pd_stage_rfft_pair(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                          int(nnodes), int(bool(to_freq)), self._stream()))
        return dst

    def slab_reduce_half(self, w, out):
        check(self.lib.pd_slab_reduce_half(self._h, self._ptr(w, None, "w"), self._ptr(out, 6 * self.half_cols, "out"),
                                           self._stream()))
        return out

    def slab_finish_half(self, w, gathered):
        check(self.lib.pd_slab_finish_half(self._h, self._ptr(w, None, "w"), self._ptr(gathered, None, "gathered"),
                                           self._stream()))
        return w

    def matvec(self, x, y=None):
        """y = A x, the Jacobian action of Build_L (Control_Wave_PC.py:86-179)."""
        if y is None:
            y = self.empty()
        check(self.lib.pd_matvec(self._h, self._ptr(x, self.size, "x"), self._ptr(y, self.size, "y"),
                                 self._stream()))
        return y

    def matvec_slab(self, x, halo_lo, halo_hi, y):
        """Slab-mode matvec (pd_matvec_slab): halos are (2, N_t) tensors or None at the domain ends."""
        lo = self._ptr(halo_lo, 2 * self.N_t, "halo_lo") if halo_lo is not None else None
        hi = self._ptr(halo_hi, 2 * self.N_t, "halo_hi") if halo_hi is not None else None
        check(self.lib.pd_matvec_slab(self._h, self._ptr(x, None, "x"), lo, hi, self._ptr(y, None, "y"),
                                      self._stream()))
        return y

    def matvec_slab_real(self, x, halo_lo, halo_hi, y):
        """pd_matvec_slab_real: float64 blocks, halos (2, N_t) float64 or None at the domain ends."""
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        check(self.lib.pd_matvec_slab_real(self._h, p(x), p(halo_lo), p(halo_hi), p(y), self._stream()))
        return y

    def maxpy(self, V, coef, sign, w, norm2_out=None):
        """w += sign * sum_i coef[i] V[i] (pd_maxpy); V is (nv, len), coef a device tensor."""
        nv, ln = V.shape
        no = self._ptr(norm2_out, 1, "norm2_out") if norm2_out is not None else None
        check(self.lib.pd_maxpy(self._h, self._ptr(V, None, "V"), int(V.stride(0)), int(nv),
                                self._ptr(coef, None, "coef"), float(sign), self._ptr(w, ln, "w"), int(ln), no,
                                self._stream()))
        return w

    def pc_matvec(self, x, y=None):
        """y = P x, the block-circulant matrix whose inverse ``pc_apply`` applies."""
        if y is None:
            y = self.empty()
        check(self.lib.pd_pc_matvec(self._h, self._ptr(x, self.size, "x"), self._ptr(y, self.size, "y"),
                                    self._stream()))
        return y

    def build_rhs(self, b=None):
        if b is None:
            b = self.empty()
        check(self.lib.pd_build_rhs(self._h, self._ptr(b, None, "b"), self._stream()))
        return b

    def set_option(self, name, value):
        check(self.lib.pd_set_option(self._h, name.encode(), float(value)))

    def delta(self, x, d=None):
        """d = (A - P) x, the operator of the residual-correction mode (pd_delta); complex128 or float64."""
        torch = _torch()
        real = x.dtype == torch.float64
        if d is None:
            d = torch.zeros_like(x)
        check(self.lib.pd_delta(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(d.data_ptr()), int(real), self._stream()))
        return d

    def gmres(self, b, x=None, rtol=1e-7, atol=1e-50, restart=300, max_it=1000, correction=False):
        """Left-preconditioned GMRES (options of Control_Wave_PC.py:347-359).  ``correction=True`` forms the
        preconditioned operator as v + P^-1 (A - P) v (pd_set_option "gmres_residual_correction").

        Returns (x, iterations, residual_history, reason)."""
        self.set_option("gmres_residual_correction", 1 if correction else 0)
        if x is None:
            x = self.empty()
        its, reason = C.c_int(0), C.c_int(0)
        hist = (C.c_double * (max_it + 1))()
        st = self.lib.pd_gmres(self._h, self._ptr(b, self.size, "b"), self._ptr(x, self.size, "x"),
                               float(rtol), float(atol), int(restart), int(max_it), C.byref(its), hist,
                               C.byref(reason), self._stream())
        check(st, allow=(_lib.PD_ERR_NOT_CONVERGED,))
        return x, its.value, list(hist[: its.value + 1]), CONVERGED_REASONS.get(reason.value, str(reason.value))

    # ---- float64 variants (real problem): half the bytes everywhere
    def _rptr(self, t, name):
        torch = _torch()
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or t.numel() != self.size:
            raise ValueError(f"{name}: need a contiguous float64 CUDA tensor with {self.size} entries")
        return C.c_void_p(t.data_ptr())

    def empty_real(self):
        torch = _torch()
        return torch.empty(self.size, dtype=torch.float64, device=f"cuda:{self.device}")

    def matvec_real(self, x, y=None):
        if y is None:
            y = self.empty_real()
        check(self.lib.pd_matvec_real(self._h, self._rptr(x, "x"), self._rptr(y, "y"), self._stream()))
        return y

    def build_rhs_real(self, b=None):
        """float64 right-hand side; on an x-slab handle ``b`` is this rank's (2, n_r, N_t) block."""
        if b is None:
            b = self.empty_real()
        torch = _torch()
        want = 2 * self.n_slab * self.N_t
        if b.dtype != torch.float64 or not b.is_cuda or not b.is_contiguous() or b.numel() != want:
            raise ValueError(f"b: need a contiguous float64 CUDA tensor with {want} entries")
        check(self.lib.pd_build_rhs_real(self._h, C.c_void_p(b.data_ptr()), self._stream()))
        return b

    def gmres_real(self, b, x=None, rtol=1e-7, atol=1e-50, restart=300, max_it=1000, correction=False):
        """pd_gmres_real: the same Krylov solve on float64 vectors with the half-spectrum preconditioner."""
        self.set_option("gmres_residual_correction", 1 if correction else 0)
        if x is None:
            x = self.empty_real()
        its, reason = C.c_int(0), C.c_int(0)
        hist = (C.c_double * (max_it + 1))()
        st = self.lib.pd_gmres_real(self._h, self._rptr(b, "b"), self._rptr(x, "x"), float(rtol), float(atol),
                                    int(restart), int(max_it), C.byref(its), hist, C.byref(reason), self._stream())
        check(st, allow=(_lib.PD_ERR_NOT_CONVERGED,))
        return x, its.value, list(hist[: its.value + 1]), CONVERGED_REASONS.get(reason.value, str(reason.value))

    def mdot(self, V, w):
        """V^H w for a (nv, len) basis tensor: PETSc VecMDot order."""
        torch = _torch()
        nv, ln = V.shape
        out = torch.empty(nv, dtype=torch.complex128, device=V.device)
        check(self.lib.pd_mdot(self._h, self._ptr(V, None, "V"), int(V.stride(0)), int(nv),
                               self._ptr(w, ln, "w"), int(ln), self._ptr(out, nv, "out"), self._stream()))
        return out
