/*
 * paradiag.h -- C ABI of libparadiag.so (B200 / sm_100a).
 *
 * Drop-in boundary for ONE hot path of Molin-Han/Optimal_Control_ParaDiag: the
 * ParaDiag block-circulant preconditioner `DiagFFTPC` that
 * Code/Control_Wave_PC.py applies inside the all-at-once GMRES solve of the 1-D
 * wave-equation optimal-control system, plus the matvec / right-hand side /
 * Krylov loop either side of it.  Every entry point names the upstream
 * interface it replaces (file:line into the upstream checkout).
 *
 * Conventions
 *   - plain C: pointers, sizes, doubles; no torch / PETSc types.
 *   - every function returns 0 on success, a negative pd_status otherwise;
 *     pd_last_error() returns a thread-local message.  The library never
 *     exits or aborts the process.
 *   - vectors are complex128 (interleaved re, im) in the PETSc layout of the
 *     reference: index(field f, node j, time i) = (f*n + j)*N_t + i with
 *     f in {0 = state u, 1 = adjoint p}, n = N_x + 1 nodes, time fastest
 *     (Control_Wave_PC.py:496-501: dat.data has shape (n, N_t), FFT on axis=1).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); work is
 *     enqueued on it, the *_host variants synchronise before returning.
 *   - the caller owns all vectors (torch tensors, PETSc Vecs); the library owns
 *     only its workspace (pd_workspace_bytes).
 */
#ifndef PARADIAG_H
#define PARADIAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_ABI_VERSION 1

typedef enum pd_status {
  PD_OK = 0,
  PD_ERR_INVALID = -1,      /* bad argument / unsupported size              */
  PD_ERR_CUDA = -2,         /* CUDA runtime error (message has the detail)  */
  PD_ERR_NOMEM = -3,        /* workspace allocation failed                  */
  PD_ERR_UNSUPPORTED = -4,  /* e.g. applyTranspose                          */
  PD_ERR_NOT_CONVERGED = -5 /* pd_gmres hit max_it (x still holds iterate)  */
} pd_status;

/* Problem description.  Replaces the module globals the upstream PC reads
 * (Control_Wave_PC.py:335-339 T, N_t, N_x, gamma; :362-368 dt, W, bcs).       */
typedef struct pd_config {
  int32_t abi_version; /* = PD_ABI_VERSION                                       */
  int32_t N_x;         /* cells of UnitIntervalMesh (:17, :337); n = N_x+1 nodes */
  int32_t N_t;         /* time levels (:336)                                     */
  int32_t bug138;      /* 1: keep the sqrt(gamma) factor on the last state row's
                          stiffness term exactly as upstream :138 (default);
                          0: drop it.  Only affects pd_matvec.                   */
  double T;            /* final time (:335)                                      */
  double gamma;        /* regulariser (:339); the sqrt(gamma) scaling of the
                          pc=True formulation is built in (:56-57, :78-80, :87)  */
  double alpha;        /* alpha-circulant weight in (0, 1]; 1.0 = the upstream operator.
                          Other values are an EXTENSION with no upstream pin (the
                          reference has no alpha): Gamma_alpha time weights around
                          the FFTs and alpha-shifted symbols in the per-frequency
                          stage, defined in oracle/pc_alpha.py.  Available on the
                          single-GPU and slab-mode applies (complex and real-input)
                          and GMRES; frequency-sharded stage handles (k_count /
                          n_local) and pd_pc_matvec return PD_ERR_UNSUPPORTED.    */
  int32_t device;      /* CUDA device ordinal                                    */
  /* Frequency shard solved by this handle in pd_stage_solve (multi-GPU):
   * global frequencies [k_begin, k_begin + k_count).  k_count = 0 means all.   */
  int32_t k_begin;
  int32_t k_count;
  /* Node slab transformed by this handle in pd_stage_fft (multi-GPU):
   * n_local lines per field.  0 means all n = N_x + 1 nodes.                    */
  int32_t n_local;
  /* Slab mode (multi-GPU without transposes): this handle owns x-slab `slab_rank` of `slab_count`
   * (balanced contiguous split of the n nodes, the first n % slab_count slabs one node longer) for ALL
   * frequencies; see pd_slab_reduce / pd_slab_finish.  slab_count <= 1: off.                        */
  int32_t slab_rank;
  int32_t slab_count;
  int32_t reserved[3];
} pd_config;

typedef struct pd_handle pd_handle;

/* Lifetime.  pd_create replaces DiagFFTPC.initialize (:380-484): instead of the
 * per-k numpy eig/inv loop (:415-436) and the MUMPS factorisation (:481-484) it
 * allocates the workspace and the time-twiddle table; all per-frequency
 * coefficients are regenerated inside the kernels.                              */
int pd_create(const pd_config* cfg, pd_handle** out);
int pd_destroy(pd_handle* h);
const char* pd_last_error(void);
int pd_abi_version(void);
size_t pd_workspace_bytes(const pd_handle* h);
/* number of kernels launched through this handle since creation                */
int64_t pd_launch_count(const pd_handle* h);

/* DiagFFTPC.apply (:491-553): y = P^-1 x, device pointers.  x and y may alias.  */
int pd_pc_apply(pd_handle* h, const void* x_dev, void* y_dev, void* stream);
/* Same with host buffers (PETSc Vec arrays, :493-497 / :552-553): H2D, apply, D2H, synchronises.  With the
 * option "host_register" (pd_set_option, or PD_HOST_REGISTER=1) pageable buffers are page-locked once per
 * (pointer, size) with cudaHostRegister and remembered (up to 16 buffers), so that long-lived KSP work vectors
 * travel at the full PCIe rate.  Opt-in: the caller must keep such buffers alive as long as the handle, or call
 * pd_host_unregister_all before freeing them.                                                        */
int pd_pc_apply_host(pd_handle* h, const void* x_host, void* y_host);
int pd_host_unregister_all(pd_handle* h);
/* Real-input fast path.  The vectors GMRES feeds the PC in this (real) problem are real: x_dev and
 * y_dev are float64 arrays in the same (field, node, time) layout, 2 n N_t doubles.  Their time spectra
 * are Hermitian, so only the frequencies 0..N_t/2 are transformed and solved: half the bytes of
 * pd_pc_apply in every stage.  Result = real part of pd_pc_apply on (x + 0i) (the imaginary part of that
 * is rounding noise).  Needs N_t >= 8 (register pipelines for the powers of two in [128, 16384], a
 * shared-memory pair kernel for every other length, N_t = 81 included); PD_ERR_UNSUPPORTED otherwise. */
int pd_pc_apply_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream);
/* pd_pc_apply_real through HOST buffers of 2 n N_t doubles (a real-scalar PETSc Vec, a numpy float64 array): H2D,
 * real-input apply, D2H -- half the PCIe bytes of pd_pc_apply_host.  Same :493-497 / :552-553 as that one.   */
int pd_pc_apply_real_host(pd_handle* h, const void* x_host, void* y_host);
/* The stages of the real-input path: real lines of N_t samples <-> half spectra of N_t/2 + 1 complex
 * numbers (to_freq != 0: scipy ifft restricted to k <= N_t/2; to_freq == 0: scipy fft of the Hermitian
 * extension, real output), and the per-frequency stage on w = (2, n, Kp), in place; rows of a half
 * spectrum are padded to Kp = (N_t/2 + 1 rounded up to a multiple of 8) complex numbers.  pd_stage_rfft (one
 * packed real line at a time) exists for the powers of two in [128, 16384] only; pd_stage_rfft_pair (below)
 * covers every N_t >= 8.                                                                              */
int pd_stage_rfft(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines, int to_freq, void* stream);
int pd_stage_solve_half(pd_handle* h, void* w_dev, void* stream);
/* One apply with CUDA events recorded on `stream` between its kernels; ms[0..4] receive
 * the device durations (milliseconds) of {inverse FFT, solve pass A, interface solve, solve
 * pass B, forward FFT}.  Synchronises the stream.  Measurement aid for bench.py.           */
int pd_pc_apply_profile(pd_handle* h, const void* x_dev, void* y_dev, void* stream, float* ms,
                        int nms);
/* DiagFFTPC.applyTranspose (:557-558): upstream raises NotImplementedError;
 * this returns PD_ERR_UNSUPPORTED.                                              */
int pd_pc_apply_transpose(pd_handle* h, const void* x_dev, void* y_dev, void* stream);

/* The three stages of the apply, exposed for the multi-GPU path (spatial slabs
 * for the FFTs, frequency slabs for the solves, an all-to-all in between).
 *   pd_stage_fft   : batched time-axis DFT of `nlines` contiguous lines of
 *                    length N_t.  inverse != 0: scipy ifft (:500-501, 1/N_t,
 *                    e^{+i..}); inverse == 0: scipy fft (:547-548).  in == out ok.
 *   pd_stage_solve : in place on w = [u-hat ; p-hat], shape (2, n, k_count),
 *                    frequency fastest: S^-1 rotation (:445-457), the 2*k_count
 *                    shifted tridiagonal solves with Dirichlet rows (:460-484,
 *                    :512), S rotation (:516-529) and 1/lambda_2 (:532-540).     */
int pd_stage_fft(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines,
                 int inverse, void* stream);
/* alpha != 1 only: the Gamma_alpha time-weight scaling of `nlines` lines, out[l][j] = in[l][j] a^(+-j)
 * with a = alpha^(1/N_t) (inverse == 0: Gamma before the inverse FFT; inverse != 0: Gamma^-1 after the
 * forward FFT).  in == out ok.  With alpha = 1 it copies nothing and returns PD_OK.                   */
int pd_stage_gamma(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines, int inverse, void* stream);
int pd_stage_solve(pd_handle* h, void* w_dev, void* stream);

/* Slab mode: the per-frequency solves with the x-direction distributed over `slab_count` ranks
 * and NO transposes.  w = this rank's (2, n_r, N_t) block after pd_stage_fft(inverse).
 *   pd_slab_reduce : local partial elimination; out_dev[6][N_t] receives, per frequency, the first
 *                    and last entry of the slab-local solve (two right-hand sides each) and the
 *                    right-hand side of the separator row this slab owns.
 *   (caller all-gathers the 6*N_t values of every rank, slab order, into gathered_dev[G][6][N_t])
 *   pd_slab_finish : every rank solves the (G-1)-row separator system of each frequency redundantly,
 *                    then back-substitutes its slab in place (rotation, Dirichlet rows included).
 * Together they replace pd_stage_solve (i.e. :445-540) for a distributed x-axis.                  */
int pd_slab_reduce(pd_handle* h, void* w_dev, void* out_dev, void* stream);
int pd_slab_finish(pd_handle* h, void* w_dev, const void* gathered_dev, void* stream);
/* Slab mode on the real-input path (float64 vectors, half spectrum k = 0..N_t/2, rows padded to Kp complex
 * numbers): pd_stage_rfft_pair transforms both fields of `nnodes` node lines, (2, nnodes, N_t) float64 <->
 * (2, nnodes, Kp) complex; pd_slab_reduce_half / pd_slab_finish_half are pd_slab_reduce / pd_slab_finish on
 * w = (2, n_r, Kp) with out (6, Kp) and gathered (G, 6, Kp).  No upstream counterpart (the reference has no
 * parallel decomposition).  N_t >= 8; PD_ERR_UNSUPPORTED otherwise.                                     */
int pd_stage_rfft_pair(pd_handle* h, const void* in_dev, void* out_dev, int64_t nnodes, int to_freq, void* stream);
int pd_slab_reduce_half(pd_handle* h, void* w_dev, void* out_dev, void* stream);
int pd_slab_finish_half(pd_handle* h, void* w_dev, const void* gathered_dev, void* stream);

/* Slab mode with the PEER-STORE EXCHANGE: the whole distributed apply behind one call, no host-launched
 * collective on the data path.  Every rank owns a small "symmetric" buffer that all ranks map (cudaIpc between
 * processes, plain peer access inside one process).  The kernel that produces a slab's functionals stores them
 * straight into slot [rank] of every rank's buffer over NVLink and publishes a per-frequency-block flag; the
 * separator-solve kernel waits (bounded) for the flags of its own frequency block.  Epochs and parities live in
 * device memory, so the sequence is capturable in a CUDA graph.  Replaces, for a distributed x-axis, the same
 * upstream lines as pd_pc_apply (:491-553); the reference has no parallel decomposition to mirror.
 *   pd_slab_comm_create  : allocate this rank's buffer; ipc_handle_out (64 bytes, may be NULL) receives its
 *                          cudaIpcMemHandle_t, *base_out (may be NULL) its device address.
 *   pd_slab_comm_connect : mode 0: `peers` = slab_count x 64-byte IPC handles in rank order (other processes);
 *                          mode 1: `peers` = void*[slab_count] device addresses valid in THIS process
 *                          (peer_devices[r] = CUDA ordinal of rank r's buffer, NULL if all on this device).
 *   pd_slab_apply        : y_local = P^-1 x on this rank's (2, n_r, N_t) complex128 block.  Collective in the
 *                          sense that every rank must call it the same number of times.
 *   pd_slab_apply_real   : the same on float64 blocks (half spectrum), N_t >= 8.
 *   pd_slab_apply_begin / _end : the two halves (up to the peer stores / from the wait on), so that one process
 *                          driving several ranks can issue all first halves before any second half (kernels that
 *                          wait on one another must never be queued on ONE GPU in the wrong order).
 *   pd_slab_apply_profile: one apply with CUDA events between its stages, ms[0..6] = {inverse FFT, pass A,
 *                          interface levels, functionals + peer stores, wait + separator solve, pass B, FFT}.
 *   pd_slab_comm_status  : *timed_out != 0 if a bounded wait expired since the last call (a peer never
 *                          delivered); *epoch = applies completed.  Synchronises the device.               */
int pd_slab_comm_create(pd_handle* h, void* ipc_handle_out, void** base_out);
int pd_slab_comm_connect(pd_handle* h, const void* peers, int mode, const int* peer_devices);
int pd_slab_comm_status(pd_handle* h, int* timed_out, uint64_t* epoch);
int pd_slab_apply(pd_handle* h, const void* x_dev, void* y_dev, void* stream);
int pd_slab_apply_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream);
int pd_slab_apply_begin(pd_handle* h, const void* x_dev, void* stream, int real_input);
int pd_slab_apply_end(pd_handle* h, void* y_dev, void* stream, int real_input);
int pd_slab_apply_profile(pd_handle* h, const void* x_dev, void* y_dev, void* stream, float* ms, int nms);

/* Matrix-free action of the Jacobian of Build_L (:86-179, pc=True branches):
 * y = A x with Dirichlet rows as identity.  x and y must not alias.             */
int pd_matvec(pd_handle* h, const void* x_dev, void* y_dev, void* stream);

/* Slab-mode version of pd_matvec: x, y are this rank's (2, n_r, N_t) blocks; halo_lo / halo_hi are the
 * (2, N_t) node rows just below / above the slab (from the neighbouring ranks; NULL at the domain ends). */
int pd_matvec_slab(pd_handle* h, const void* x_dev, const void* halo_lo_dev, const void* halo_hi_dev,
                   void* y_dev, void* stream);

/* The same for float64 blocks (2 n_r N_t doubles; halos 2 N_t doubles) -- the float64 distributed Krylov loop.       */
int pd_matvec_slab_real(pd_handle* h, const void* x_dev, const void* halo_lo_dev, const void* halo_hi_dev,
                        void* y_dev, void* stream);

/* y = P x for the block-circulant matrix P that DiagFFTPC inverts (the operator above
 * with the time stencils of :121, :137 made periodic -- C1, C2 of mat_test.ipynb cells
 * 8-9 -- and the half weights :117, :143 and the :138 factor replaced by 1).  Lets a
 * caller verify an apply at any size: P (P^-1 x) = x on interior rows.               */
int pd_pc_matvec(pd_handle* h, const void* x_dev, void* y_dev, void* stream);

/* Right-hand side of the manufactured problem, Build_f / Build_g /
 * Build_Initial_Condition (:48-83) folded through the residual (:118, :139,
 * :144, :93-95): b such that the ksponly solve is A U = b.                      */
int pd_build_rhs(pd_handle* h, void* b_dev, void* stream);

/* KSP GMRES as configured at :347-359 (left PC, classical Gram-Schmidt, zero
 * initial guess, preconditioned-residual test against rtol*||P^-1 b||).
 *   hist      : optional, length max_it+1, receives the residual-norm history
 *               (what -ksp_monitor prints); hist[0] = ||P^-1 b||.
 *   its       : Krylov iterations taken.
 *   reason    : PETSc-style converged reason (2 = RTOL, 3 = ATOL, -3 = ITS).
 * Returns PD_OK when converged, PD_ERR_NOT_CONVERGED at max_it.                 */
int pd_gmres(pd_handle* h, const void* b_dev, void* x_dev, double rtol, double atol,
             int restart, int max_it, int* its, double* hist, int* reason,
             void* stream);

/* Run-time options of a handle.  "host_register" (default 0): see pd_pc_apply_host.
 * "krylov_real_vectors" (default 0): pd_mdot / pd_maxpy are handed float64 vectors viewed as complex pairs.
 * "slab_overlap" (default 1): pd_slab_apply runs the per-frequency stage as two frequency halves on two streams so
 * that the latency-bound interface / separator kernels of one half overlap the streaming passes of the other; 0
 * keeps everything on the caller's stream (needed when one process drives several ranks on one GPU).
 * "gmres_residual_correction" (default 0): with 1, pd_gmres / pd_gmres_real form
 * the preconditioned operator as  v + P^-1 ((A - P) v)  instead of  P^-1 (A v)  -- the same operator in exact
 * arithmetic (Krylov vectors have zero Dirichlet rows), but (A - P) v only touches the wrap-around time levels
 * and the half-weight rows (:117, :143, :93-110, :138), so the cancelling second differences of A v are never
 * formed and the rounding noise that the ill-conditioned P^-1 amplifies is gone.  Opt-in: the default keeps KSP's
 * order of operations (:347-359).                                                                          */
int pd_set_option(pd_handle* h, const char* name, double value);
/* d = (A - P) x (complex128, or float64 when real_vectors != 0).  d must be zero on entry outside the at most three
 * time levels per field that A - P touches; only those entries are written.                                */
int pd_delta(pd_handle* h, const void* x_dev, void* d_dev, int real_vectors, void* stream);

/* float64 variants for the real problem (vectors of 2 n N_t doubles, same layout): the matvec, the
 * right-hand side and the whole GMRES solve with the half-spectrum preconditioner pd_pc_apply_real.
 * Same iteration as pd_gmres at half the memory traffic.                                          */
int pd_matvec_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream);
int pd_build_rhs_real(pd_handle* h, void* b_dev, void* stream);
int pd_gmres_real(pd_handle* h, const void* b_dev, void* x_dev, double rtol, double atol, int restart,
                  int max_it, int* its, double* hist, int* reason, void* stream);

/* Batched reductions used by the multi-GPU Krylov loop (device results):
 * out[i] = sum_j conj(V[i*ld + j]) * w[j], i < nv  (PETSc VecMDot order).       */
int pd_mdot(pd_handle* h, const void* V_dev, int64_t ld, int nv, const void* w_dev,
            int64_t len, void* out_dev, void* stream);

/* w += sign * sum_{i<nv} coef[i] V[i*ld + .] with device-resident coefficients (PETSc VecMAXPY);
 * norm2_out_dev (optional, one complex) receives ||w_new||^2 of the local part.                     */
int pd_maxpy(pd_handle* h, const void* V_dev, int64_t ld, int nv, const void* coef_dev, double sign,
             void* w_dev, int64_t len, void* norm2_out_dev, void* stream);

/* The host-side algebra of one restarted-GMRES cycle as KSPGMRES keeps it (the solver :347-359 selects): Hessenberg
 * column update with Givens rotations, residual-norm estimate, back substitution.  Host memory only, no CUDA calls.
 * pd_gmres / pd_gmres_real use exactly this object internally; a distributed Krylov loop (dist.py) that owns its
 * own collectives calls it through these entry points, so there is one implementation of the recurrence.
 *   pd_hess_create : state for cycles of up to `restart` columns.
 *   pd_hess_start  : begin a cycle with ||r0|| = beta.
 *   pd_hess_push   : append column j (the j-th call of the cycle): hcol = j + 2 complex numbers (re, im pairs), the
 *                    classical Gram-Schmidt coefficients h_0..h_j followed by the SQUARED norm of the orthogonalised
 *                    vector (real part).  *resnorm_out = |g_{j+1}|, *hnorm_out (optional) = h_{j+1,j}.
 *   pd_hess_solve  : y = H^-1 g for the columns pushed in this cycle (y_out: that many complex numbers).       */
typedef struct pd_hessenberg pd_hessenberg;
int pd_hess_create(int restart, pd_hessenberg** out);
int pd_hess_destroy(pd_hessenberg* q);
int pd_hess_start(pd_hessenberg* q, double beta);
int pd_hess_push(pd_hessenberg* q, const void* hcol, double* resnorm_out, double* hnorm_out);
int pd_hess_solve(pd_hessenberg* q, void* y_out, int* ncol_out);

#ifdef __cplusplus
}
#endif
#endif /* PARADIAG_H */
