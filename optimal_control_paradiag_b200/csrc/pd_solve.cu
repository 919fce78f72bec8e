// Frequency-domain stage of the DiagFFTPC apply (complex128, sm_100a).
//
// Replaces, per frequency k (Control_Wave_PC.py line numbers):
//   :445-457  right-hand side  S_k^-1 [u-hat; p-hat]
//   :460-484, :512  the two shifted solves (Sigma_i(k) M + dt^2/2 K) w = rhs with
//             homogeneous Dirichlet rows (MUMPS LU of the monolithic D upstream)
//   :516-529  multiplication by S_k
//   :532-540  division by lambda_2(k) / conj(lambda_2(k))
//
// Nothing per-frequency is stored: with theta = 2 pi k / N_t, z = e^{i theta},
// sigma = sign(cos theta), kappa = dt^2 cos(theta), c = dt^2/sqrt(gamma) the
// closed forms (pre_cond.py:32-38, mat_test.ipynb cell 1) give the division-free
// form
//   rho_+- = (u-hat / z +- i sigma p-hat) / 2
//   Tt zeta_+ = rho_+ ,  conj(Tt) zeta_- = rho_-     (interior nodes)
//   Tt = tridiag(a, b, a),  a = s h/6 - kappa/h,  b = 2 s h/3 + 2 kappa/h,
//   s = -4 sin^2(theta/2) + i c sigma
//   w_u = zeta_+ + zeta_- ,  w_p = -i sigma z (zeta_+ - zeta_-)
// so each k is ONE complex-symmetric Toeplitz tridiagonal matrix with two
// right-hand sides (rho_+ and conj(rho_-)).
//
// Algorithm.  Layout [field][node][k], k fastest, so a warp touches 512
// contiguous bytes per node row and every access below is coalesced.  The
// systems are solved by recursive partitioning:
//   level 0  the m interior rows are cut into chunks of PD_L = 16 rows with one
//            separator row between chunks.  pass A (1 read sweep of the big
//            array) runs a register-resident forward recurrence per (k, chunk),
//            emits the first entry f_c of the local solve and folds the last
//            entry l_c into the right-hand side of the separator that follows.
//   level l  the separators of level l-1 form a tridiagonal interface system
//            (Toeplitz except its last diagonal entry -- a structure the reduction
//            preserves, so its coefficients too are regenerated, never stored).
//            While it has more than PD_PCR_MAX rows it is reduced again the same way
//            with chunks of PD_LG = 8 rows (generic kernels, ~6 % of the data).
//   top      the last interface system (<= 32 rows per k) is solved by parallel
//            cyclic reduction held in shared memory, up to 32 frequencies per CTA.
//   back     the generic levels are back-substituted, then pass B (1 read + 1 write
//            sweep of the big array) runs Thomas per (k, chunk) with the now-known
//            separator values, rotates back and stores in place.
// The level-0 pivots m_i depend on (k, i) only; each CTA regenerates them once into a
// per-thread shared-memory column and reuses them for all the chunks it visits.
#include <stdlib.h>
#include <string.h>

#include "pd_common.cuh"


#include "pd_solve_dev.cuh"

// ------------------------------------------------------------------- pass A
template <bool AL>
__global__ void __launch_bounds__(PD_KB, 4)
pd_solve_passA_kernel(const cplx* __restrict__ w, cplx* __restrict__ F0, cplx* __restrict__ R1,
                      SolveParams sp, cplx* __restrict__ lastl) {
  pd_pdl_enter((sp.pdl_early & 2) != 0);  // programmatic dependent launch: see pd_common.cuh
  __shared__ cplx mtab[PD_L][PD_KB];
  const int tid = threadIdx.x;
  const int kk = sp.koff + blockIdx.x * PD_KB + tid;
  const bool valid = kk < sp.kend;
  const int kc_idx = valid ? kk : sp.kend - 1;
  const KCoef kc = make_coef<AL>(freq_of(sp, kc_idx), sp);
  fill_pivots<PD_KB>(kc, mtab, tid);
  const cplx* wu = w + kc_idx;
  const cplx* wp = w + sp.plane + kc_idx;
  for (int c = sp.c0 + blockIdx.y; c < sp.c1; c += gridDim.y)
    passA_chunk<AL, PD_KB, false>(wu, wp, F0, R1, sp, lastl, kc, mtab, tid, kk, valid, c);
}

// ------------------------------------------------ generic level: reduce (1 <= lev < top)
// thread = (k, chunk c of level lev).  Finalises and stores the rhs of its rows, runs the forward
// recurrences, emits f_c and the partial rhs of its trailing separator for level lev + 1.
__global__ void __launch_bounds__(PD_KB)
pd_solve_level_reduce_kernel(Levels lv, SolveParams sp, int lev) {
  pd_pdl_enter((sp.pdl_early & 4) != 0);  // programmatic dependent launch: see pd_common.cuh
  const int kk = sp.koff + blockIdx.x * PD_KB + threadIdx.x;
  if (kk >= sp.kend) return;
  const KCoef kc = make_coef(freq_of(sp, kk), sp);
  const Sys below = level_sys(kc, sp, lev - 1);
  const Sys s = reduce_sys(below, chunk_len(lev - 1));
  const int64_t K = sp.K;
  const int P = sp.rows[lev + 1], Llast = s.n - P * (PD_LG + 1);
  cplx* R = lv.R[lev];
  const cplx* Fb = lv.F[lev - 1];
  for (int c = blockIdx.y; c <= P; c += gridDim.y) {
    const int Lc = c < P ? PD_LG : Llast;
    const int q0 = c * (PD_LG + 1);
    cplx dP = cmake(0, 0), dM = cmake(0, 0), fP = cmake(0, 0), fM = cmake(0, 0);
    cplx pi = cmake(1, 0), m = cmake(0, 0);
    PivotGen pg;
    pg.init(s);
#pragma unroll 4
    for (int i = 0; i < Lc; ++i) {
      const int64_t q = q0 + i;
      const cplx rP = cfms(below.off, Fb[((q + 1) * 2) * K + kk], R[(q * 2) * K + kk]);
      const cplx rM = cfms(below.off, Fb[((q + 1) * 2 + 1) * K + kk], R[(q * 2 + 1) * K + kk]);
      R[(q * 2) * K + kk] = rP;
      R[(q * 2 + 1) * K + kk] = rM;
      if (i > 0) pi = cneg(cmul(pi, cmul(s.off, m)));
      m = pg.next();
      if (q == s.n - 1) m = last_row_pivot(m, s.glast);
      dP = cmul(cfms(s.off, dP, rP), m);
      dM = cmul(cfms(s.off, dM, rM), m);
      fP = cfma(pi, dP, fP);
      fM = cfma(pi, dM, fM);
    }
    lv.F[lev][((int64_t)c * 2) * K + kk] = fP;
    lv.F[lev][((int64_t)c * 2 + 1) * K + kk] = fM;
    if (c < P) {
      const int64_t q = q0 + PD_LG;
      const cplx rP = cfms(below.off, Fb[((q + 1) * 2) * K + kk], R[(q * 2) * K + kk]);
      const cplx rM = cfms(below.off, Fb[((q + 1) * 2 + 1) * K + kk], R[(q * 2 + 1) * K + kk]);
      lv.R[lev + 1][((int64_t)c * 2) * K + kk] = cfms(s.off, dP, rP);
      lv.R[lev + 1][((int64_t)c * 2 + 1) * K + kk] = cfms(s.off, dM, rM);
    }
  }
}

// ------------------------------------------- generic level: back substitution (1 <= lev < top)
// thread = (k, chunk c).  Separator solutions live in R[lev+1]; the chunk rows (and a copy of the
// trailing separator) are overwritten by the solution in R[lev].
__global__ void __launch_bounds__(PD_KB)
pd_solve_level_back_kernel(Levels lv, SolveParams sp, int lev) {
  pd_pdl_enter((sp.pdl_early & 4) != 0);  // programmatic dependent launch: see pd_common.cuh
  __shared__ cplx mtab[PD_LG][PD_KB];  // per-thread column of chunk pivots
  const int kk = sp.koff + blockIdx.x * PD_KB + threadIdx.x;
  if (kk >= sp.kend) return;
  const KCoef kc = make_coef(freq_of(sp, kk), sp);
  const Sys s = level_sys(kc, sp, lev);
  const int64_t K = sp.K;
  const int P = sp.rows[lev + 1], Llast = s.n - P * (PD_LG + 1);
  cplx* R = lv.R[lev];
  const cplx* Z = lv.R[lev + 1];
  const cplx zero = cmake(0, 0);
  for (int c = blockIdx.y; c <= P; c += gridDim.y) {
    const int Lc = c < P ? PD_LG : Llast;
    const int q0 = c * (PD_LG + 1);
    cplx zlP = zero, zlM = zero, zrP = zero, zrM = zero;
    if (c > 0) {
      zlP = Z[((int64_t)(c - 1) * 2) * K + kk];
      zlM = Z[((int64_t)(c - 1) * 2 + 1) * K + kk];
    }
    if (c < P) {
      zrP = Z[((int64_t)c * 2) * K + kk];
      zrM = Z[((int64_t)c * 2 + 1) * K + kk];
      R[((int64_t)(q0 + PD_LG) * 2) * K + kk] = zrP;
      R[((int64_t)(q0 + PD_LG) * 2 + 1) * K + kk] = zrM;
    }
    // all rows of the chunk are loaded up front (independent 128-bit loads), then Thomas in registers
    cplx dP[PD_LG], dM[PD_LG];
#pragma unroll
    for (int i = 0; i < PD_LG; ++i) {
      if (i < Lc) {
        dP[i] = R[((int64_t)(q0 + i) * 2) * K + kk];
        dM[i] = R[((int64_t)(q0 + i) * 2 + 1) * K + kk];
      }
    }
    cplx pP = zlP, pM = zlM;
    PivotGen pg;
    pg.init(s);
#pragma unroll
    for (int i = 0; i < PD_LG; ++i) {
      if (i < Lc) {
        cplx rP = dP[i], rM = dM[i];
        if (i == Lc - 1) {
          rP = cfms(s.off, zrP, rP);
          rM = cfms(s.off, zrM, rM);
        }
        cplx m = pg.next();
        if (q0 + i == s.n - 1) m = last_row_pivot(m, s.glast);
        mtab[i][threadIdx.x] = m;
        pP = cmul(cfms(s.off, pP, rP), m);
        pM = cmul(cfms(s.off, pM, rM), m);
        dP[i] = pP;
        dM[i] = pM;
      }
    }
    cplx nP = zero, nM = zero;
#pragma unroll
    for (int i = PD_LG - 1; i >= 0; --i) {
      if (i < Lc) {
        if (i < Lc - 1) {
          const cplx cp = cmul(s.off, mtab[i][threadIdx.x]);
          nP = cfms(cp, nP, dP[i]);
          nM = cfms(cp, nM, dM[i]);
        } else {
          nP = dP[i];
          nM = dM[i];
        }
        R[((int64_t)(q0 + i) * 2) * K + kk] = nP;
        R[((int64_t)(q0 + i) * 2 + 1) * K + kk] = nM;
      }
    }
  }
}

// ---- peer-store exchange of the slab functionals (the fused compute + collective step of slab mode)
// Every rank owns one "symmetric" buffer, mapped into all ranks (cudaIpc between processes, plain peer access
// inside one process):   gathered[2 parities][G ranks][6][kmax] complex  +  flags[2][G][nflag] (uint64 epochs).
// The kernel that produces a slab's six functionals per frequency STORES them straight into slot [rank] of
// every rank's buffer (NVLink peer stores, fire and forget), fences, and publishes one flag per (rank, block of
// PD_KB frequencies) = the apply's epoch.  The kernel that consumes them (separator solve) waits at its head for
// the G flags of ITS frequency block only.  No host-launched collective, no extra kernel, no barrier:
//   * parity = epoch & 1 double-buffers the slots: a rank can be at most one apply ahead of a peer (its
//     epoch e+1 separator solve needs the peer's e+1 functionals, which the peer issues after finishing e);
//   * epochs live in device memory (bumped by pass B), so the whole apply is capturable in a CUDA graph;
//   * the wait is bounded (PD_SLAB_SPIN_LIMIT clock ticks): on expiry the kernel raises err[0] and goes on,
//     the host reports it (pd_slab_comm_status) -- a dead peer can never hang the GPU.
#define PD_SLAB_SPIN_LIMIT (8ll * 1000 * 1000 * 1000)
struct SlabCommDev {
  cplx* peer_gath[PD_MAX_SLABS];                  // every rank's gathered buffer (own included)
  unsigned long long* peer_flag[PD_MAX_SLABS];    // every rank's flag array
  const unsigned long long* epoch;                // completed applies of THIS rank
  int* err;
  int64_t kmax;
  int nflag, G, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// frequencies per flag of the exchange (every producer covers whole groups of PD_FKB columns: the sequential interface
// kernel PD_ITK = 32 per CTA, the functionals kernel PD_KB = 128, the PCR kernel when it holds 32 per CTA).  One flag
// per producing CTA and peer keeps the number of system-scope release stores -- each a fence -- at G per CTA.
#define PD_FKB 32

// slot [parity][rank] of EVERY rank's buffer (peer stores over NVLink; the own copy is a local store)
__device__ __forceinline__ void slab_push(const SlabCommDev& cm, unsigned long long ep, int kk, cplx fP, cplx fM,
                                          cplx lP, cplx lM, cplx sP, cplx sM) {
  const int64_t slot = (((int64_t)(ep & 1ull) * cm.G + cm.rank) * 6) * cm.kmax + kk;
  for (int p = 0; p < cm.G; ++p) {
    cplx* g = cm.peer_gath[p] + slot;
    g[0] = fP; g[cm.kmax] = fM; g[2 * cm.kmax] = lP; g[3 * cm.kmax] = lM;
    g[4 * cm.kmax] = sP; g[5 * cm.kmax] = sM;
  }
}
// Publish the columns [k0, k0 + ncols) of this CTA to every rank: called by ALL threads of the CTA after their
// slab_push calls.  The block barrier orders every thread's (weak) data stores before the flag writers' release
// stores, and a release at system scope is cumulative over what its thread has observed through the barrier -- the
// "all store, barrier, one thread releases" pattern.  (A __threadfence_system() by every thread in front of the
// barrier was measured as 19 % of the interface kernel's time in slab mode, ncu: ERRBAR.)
__device__ __forceinline__ void slab_publish(const SlabCommDev& cm, unsigned long long ep, int k0, int ncols) {
  __syncthreads();
  const int f0 = k0 / PD_FKB, nf = (ncols + PD_FKB - 1) / PD_FKB;
  for (int i = threadIdx.x; i < nf * cm.G; i += blockDim.x) {
    const int p = i / nf, f = f0 + (i - p * nf);
    st_release_sys(cm.peer_flag[p] + ((int64_t)(ep & 1ull) * cm.G + cm.rank) * cm.nflag + f, ep);
  }
}
// Wait (bounded) until the columns [k0, k0 + ncols) have arrived from every rank; all threads of the CTA call it.
__device__ __forceinline__ void slab_wait(const SlabCommDev& cm, unsigned long long ep, int k0, int ncols) {
  const int f0 = k0 / PD_FKB, nf = (ncols + PD_FKB - 1) / PD_FKB;
  for (int i = threadIdx.x; i < nf * cm.G; i += blockDim.x) {
    const int src = i / nf, f = f0 + (i - src * nf);
    const unsigned long long* fl = cm.peer_flag[cm.rank] + ((int64_t)(ep & 1ull) * cm.G + src) * cm.nflag + f;
    const long long t0 = clock64();
    while (ld_acquire_sys(fl) < ep) {
      if (clock64() - t0 > PD_SLAB_SPIN_LIMIT) {
        atomicExch(cm.err, 1);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// The six functionals of this slab for one frequency: first / last entry of the slab-local solve (two right-hand
// sides each) and the rotated right-hand side of the separator row this slab owns.
//   f0P/f0M : F[0] of chunk 0 (zero without interface), lastP/lastM : last entry of the last chunk (pass A),
//   z0*, ze* : interface solution at the first / last level-1 row (unused when P == 0)
__device__ __forceinline__ void slab_functionals(const KCoef& kc, const SolveParams& sp, const cplx* __restrict__ w,
                                                 int kk, cplx f0P, cplx f0M, cplx lastP, cplx lastM, cplx z0P,
                                                 cplx z0M, cplx zeP, cplx zeM, cplx& fP, cplx& fM, cplx& lP, cplx& lM,
                                                 cplx& sP, cplx& sM) {
  const int P = sp.rows[1], Llast = sp.m - P * (PD_L + 1);
  fP = f0P; fM = f0M;
  lP = cmake(0, 0); lM = cmake(0, 0);
  if (Llast > 0) { lP = lastP; lM = lastM; }
  if (P > 0) {
    // first entry of chunk 0 and last entry of the last chunk, given the interface solution
    VRec v;
    v.init(kc.a, kc.sh, cmake(0, 0));
    cplx rvL = cmake(0, 0), rvLl = cmake(0, 0);
    if (!v.diag) {
      for (int i = 1; i <= PD_L; ++i) {
        v.step();
        if (i == Llast) rvLl = cscale(crcp(v.V), v.one);
      }
      rvL = cscale(crcp(v.V), v.one);
    }
    fP = cfma(z0P, rvL, fP);  // f_0 - a z_sep0 (T_L^-1)_{1L} = f_0 + z_sep0 / V_L
    fM = cfma(z0M, rvL, fM);
    if (Llast > 0) {
      lP = cfma(zeP, rvLl, lP);
      lM = cfma(zeM, rvLl, lM);
    } else {
      lP = zeP;  // the last body row is the last separator itself
      lM = zeM;
    }
  }
  sP = cmake(0, 0); sM = cmake(0, 0);
  if (!sp.first_dirichlet) {
    if (sp.al) rotate_in<true>(kc, w[kk], w[sp.plane + kk], sP, sM);
    else rotate_in<false>(kc, w[kk], w[sp.plane + kk], sP, sM);
  }
}

// ------------------------------------------------ top interface system (PCR in smem)
// One CTA solves the top-level systems of `kpb` consecutive frequencies (n <= PD_PCR_MAX rows
// each, two right-hand sides) by parallel cyclic reduction.  Rows are kept normalised (unit
// diagonal): (lo, 1, up | rP, rM).  smem slot = ks * n + q.
// PUSH (slab mode, lev == 1 == top): the level-1 solution is still in shared memory when the kernel ends, so the
// slab functionals are formed and stored into every rank's exchange buffer right here (no separate launch).
template <bool PUSH>
__global__ void __launch_bounds__(PD_PCR_THREADS)
pd_solve_pcr_kernel(Levels lv, SolveParams sp, int lev, int kpb, const cplx* __restrict__ w, SlabPtrs sl,
                    SlabCommDev cm) {
  pd_pdl_enter((sp.pdl_early & 4) != 0);  // programmatic dependent launch: see pd_common.cuh
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  __shared__ cplx c_off[32], c_dmain[32], c_dlast[32], c_offb[32];
  const int n = sp.rows[lev];
  const int rows = n * kpb;
  cplx* s_lo = reinterpret_cast<cplx*>(pd_smem_raw);
  cplx* s_up = s_lo + rows;
  cplx* s_rp = s_up + rows;
  cplx* s_rm = s_rp + rows;
  const int tid = threadIdx.x;
  const int kk0 = sp.koff + blockIdx.x * kpb;
  const int64_t K = sp.K;

  // coefficients of this CTA's frequencies, regenerated once
  if (tid < kpb) {
    int kk = kk0 + tid;
    if (kk >= sp.kend) kk = sp.kend - 1;
    const KCoef kc = make_coef(freq_of(sp, kk), sp);
    const Sys below = level_sys(kc, sp, lev - 1);
    const Sys s = reduce_sys(below, chunk_len(lev - 1));
    // at the top level the detuning has been amplified by (L+1)^2 per level: plain sums are safe here
    const cplx dm = sys_dmain(s);
    c_off[tid] = s.off; c_dmain[tid] = dm; c_dlast[tid] = cadd(dm, s.glast); c_offb[tid] = below.off;
  }
  __syncthreads();

  // build: idx -> (q = idx / kpb, ks = idx % kpb) so that consecutive threads read consecutive k
  const cplx* R = lv.R[lev];
  const cplx* Fb = lv.F[lev - 1];
  for (int idx = tid; idx < rows; idx += PD_PCR_THREADS) {
    const int q = idx / kpb, ks = idx - q * kpb;
    int kk = kk0 + ks;
    if (kk >= sp.kend) kk = sp.kend - 1;
    const cplx rP = cfms(c_offb[ks], Fb[((int64_t)(q + 1) * 2) * K + kk], R[((int64_t)q * 2) * K + kk]);
    const cplx rM = cfms(c_offb[ks], Fb[((int64_t)(q + 1) * 2 + 1) * K + kk], R[((int64_t)q * 2 + 1) * K + kk]);
    const cplx dinv = crcp(q == n - 1 ? c_dlast[ks] : c_dmain[ks]);
    const cplx od = cmul(c_off[ks], dinv);
    const int slot = ks * n + q;
    s_lo[slot] = q > 0 ? od : cmake(0, 0);
    s_up[slot] = q + 1 < n ? od : cmake(0, 0);
    s_rp[slot] = cmul(rP, dinv);
    s_rm[slot] = cmul(rM, dinv);
  }
  __syncthreads();

  for (int delta = 1; delta < n; delta <<= 1) {
    cplx nlo[PD_PCR_MAXROWS], nup[PD_PCR_MAXROWS], nrp[PD_PCR_MAXROWS], nrm[PD_PCR_MAXROWS];
#pragma unroll
    for (int it = 0; it < PD_PCR_MAXROWS; ++it) {
      const int idx = tid + it * PD_PCR_THREADS;
      if (idx < rows) {
        const int q = idx / kpb, ks = idx - q * kpb;
        const int slot = ks * n + q;
        const cplx l = s_lo[slot], u = s_up[slot];
        cplx diag = cmake(1, 0), rp = s_rp[slot], rm = s_rm[slot];
        cplx l2 = cmake(0, 0), u2 = cmake(0, 0);
        if (q - delta >= 0) {
          const int sl = slot - delta;
          diag = cfms(l, s_up[sl], diag);
          rp = cfms(l, s_rp[sl], rp);
          rm = cfms(l, s_rm[sl], rm);
          l2 = cneg(cmul(l, s_lo[sl]));
        }
        if (q + delta < n) {
          const int sl = slot + delta;
          diag = cfms(u, s_lo[sl], diag);
          rp = cfms(u, s_rp[sl], rp);
          rm = cfms(u, s_rm[sl], rm);
          u2 = cneg(cmul(u, s_up[sl]));
        }
        const cplx dinv = crcp(diag);
        nlo[it] = cmul(l2, dinv);
        nup[it] = cmul(u2, dinv);
        nrp[it] = cmul(rp, dinv);
        nrm[it] = cmul(rm, dinv);
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < PD_PCR_MAXROWS; ++it) {
      const int idx = tid + it * PD_PCR_THREADS;
      if (idx < rows) {
        const int q = idx / kpb, ks = idx - q * kpb;
        const int slot = ks * n + q;
        s_lo[slot] = nlo[it];
        s_up[slot] = nup[it];
        s_rp[slot] = nrp[it];
        s_rm[slot] = nrm[it];
      }
    }
    __syncthreads();
  }

  cplx* Rw = lv.R[lev];
  for (int idx = tid; idx < rows; idx += PD_PCR_THREADS) {
    const int q = idx / kpb, ks = idx - q * kpb;
    const int kk = kk0 + ks;
    if (kk >= sp.kend) continue;
    const int slot = ks * n + q;
    Rw[((int64_t)q * 2) * K + kk] = s_rp[slot];
    Rw[((int64_t)q * 2 + 1) * K + kk] = s_rm[slot];
  }
  if (PUSH) {
    const unsigned long long ep = *cm.epoch + 1ull;
    const int kk = kk0 + tid;
    if (tid < kpb && kk < sp.kend) {
      const KCoef kc = make_coef(freq_of(sp, kk), sp);
      const int b0 = tid * n, b1 = tid * n + n - 1;
      cplx fP, fM, lP, lM, sP, sM;
      slab_functionals(kc, sp, w, kk, lv.F[0][kk], lv.F[0][K + kk], sl.lastl[kk], sl.lastl[K + kk], s_rp[b0], s_rm[b0],
                       s_rp[b1], s_rm[b1], fP, fM, lP, lM, sP, sM);
      slab_push(cm, ep, kk, fP, fM, lP, lM, sP, sM);
    }
    slab_publish(cm, ep, kk0, min(kpb, sp.kend - kk0));
  }
}

// --------------------------------------- the whole level-1 interface system in ONE launch (sequential LU)
// thread = one frequency (both right-hand sides).  The level-1 system tridiag(off, dmain, off) with its modified
// last row is factorised by the same cancellation-free pivot generator as the chunk-local systems
// (m_q = -V_{q-1} / (off V_q), PivotGen) and swept once forward, once backward, straight out of global memory:
// every row access of a warp is one contiguous 512-byte segment, the loads of the next PD_IT rows are issued
// before the dependent recurrences of the current ones.
// Replaces the reduce / PCR / back chain (3-7 short launches) wherever the interface is short enough for the
// sequential latency (rows[1] <= PD_ITHOMAS_MAX): small N_x and, above all, the x-slabs of a multi-GPU run, where
// that chain of launches was the part of the apply that did not shrink with the number of GPUs.
// PUSH (slab mode): the thread ends with the first and last interface values in registers, forms the slab
// functionals and stores them into every rank's exchange buffer (no separate launch).
#define PD_IT 4
// The interface pivots m_q(k) depend on (k, q) only.  Generating them inside the sweep costs one double-precision
// division per row on the critical path of a kernel that runs ONE warp per scheduler (measured 0.46 us per row);
// the plan therefore factorises the interface once (this kernel) and keeps the pivots, [rows[1]][K] complex,
// 3 % of a vector -- the per-frequency coefficients of the big level-0 systems are still regenerated in-kernel.
__global__ void __launch_bounds__(PD_KB)
pd_iface_pivots_kernel(SolveParams sp, cplx* __restrict__ piv) {
  // TWISTED factorisation: rows 0 .. mid-1 are eliminated from the top, rows P-1 .. mid from the bottom (whose first
  // row carries the modified diagonal dmain + glast), the two sweeps meet between rows mid-1 and mid.  Two threads
  // per right-hand side then run the two sweeps of the solve concurrently: half the critical path of a kernel whose
  // run time IS its critical path.
  const int kk = blockIdx.x * PD_KB + threadIdx.x;
  if (kk >= sp.K) return;
  const KCoef kc = make_coef(freq_of(sp, kk), sp);
  const Sys s = reduce_sys(level_sys(kc, sp, 0), PD_L);
  const int P = sp.rows[1], mid = P / 2;
  PivotGen pg;
  pg.init(s);
  for (int q = 0; q < mid; ++q) piv[(int64_t)q * sp.K + kk] = pg.next();
  // from the bottom: the same generator with glast / off added to eta in its first step (cf. reduce_sys)
  PivotGen pb;
  pb.v.init(s.off, s.det, cmake(0, 0));  // (sets `diag`: decoupled systems have no off-diagonal to divide by)
  if (!pb.v.diag) pb.v.init(s.off, s.det, cmul(s.glast, crcp(s.off)));
  pb.roff = pb.v.diag ? cmake(0, 0) : crcp(s.off);
  pb.mdiag = pb.v.diag ? crcp(sys_dmain(s)) : cmake(0, 0);
  for (int q = P - 1; q >= mid; --q) {
    cplx m = pb.next();
    if (pb.v.diag && q == P - 1) m = crcp(cadd(sys_dmain(s), s.glast));
    piv[(int64_t)q * sp.K + kk] = m;
  }
}

// thread = (frequency, right-hand side, direction): a half-warp owns 16 consecutive frequencies of one right-hand
// side and one sweep direction, so every row access is a contiguous 256-byte segment.  The kernel runs one warp per
// scheduler and nothing but its own prefetching hides the latency of its loads: measured 0.42 us per row with one
// batch of 4 rows in flight in registers, 0.30 us with two.  The rows therefore stream through a PER-THREAD RING IN
// SHARED MEMORY filled by cp.async (LDGSTS): PD_IRING batches of PD_IT rows are in flight ahead of the elimination,
// every thread consumes only what it copied itself (cp.async.wait_group); the only block barrier is the meeting of
// the top-down and the bottom-up sweep of the twisted factorisation.
#define PD_ITK 32    // frequencies per CTA of the sequential interface kernel (4 threads each)
// ring depth (batches of PD_IT rows) = template parameter: 8 when the grid fits the GPU with one CTA per SM (196 KB of
// ring each), 4 / 2 when there are more CTAs than SMs (N_t = 16384: 512 CTAs) -- then several CTAs per SM hide each
// other's latency and one resident wave beats a deep ring in 3.5 waves (cfg4: 1.9 ms -> see DESIGN.md)
#define PD_ISMEM(RING) ((RING) * PD_IT * 3 * 4 * PD_ITK * 16)
__device__ __forceinline__ void cp_async16(cplx* smem_dst, const cplx* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool PUSH, int PD_IRING>
__global__ void __launch_bounds__(4 * PD_ITK)
pd_solve_iface_thomas_kernel(Levels lv, SolveParams sp, const cplx* __restrict__ piv, const cplx* __restrict__ w,
                             SlabPtrs sl, SlabCommDev cm) {
  pd_pdl_enter((sp.pdl_early & 4) != 0);  // programmatic dependent launch: see pd_common.cuh
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  __shared__ cplx xch[4 * PD_ITK][2];                                // meeting point: (last value, off * pivot)
  constexpr int NT = 4 * PD_ITK;
  cplx* ring = reinterpret_cast<cplx*>(pd_smem_raw) + threadIdx.x;   // slot (batch, item) of this thread: + (batch * 12 + item) * NT
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int rhs = lane >> 4;                                         // 0: the + system, 1: the (conjugated) - system
  const bool up = (wrp & 2) != 0;                                    // warps 2, 3: the bottom-up sweep
  const int k0 = sp.koff + blockIdx.x * PD_ITK;
  const int kk = k0 + (wrp & 1) * 16 + (lane & 15);
  const bool active = kk < sp.kend;
  const int kc_idx = active ? kk : sp.kend - 1;                      // idle threads still meet at the barrier
  const unsigned long long ep = PUSH ? *cm.epoch + 1ull : 0ull;
  const KCoef kc = make_coef(freq_of(sp, kc_idx), sp);
  const Sys below = level_sys(kc, sp, 0);
  const Sys s = reduce_sys(below, PD_L);
  const int64_t K = sp.K;
  const int P = sp.rows[1], mid = P / 2;
  cplx* R = lv.R[1] + (int64_t)rhs * K + kc_idx;
  const cplx* F = lv.F[0] + (int64_t)rhs * K + kc_idx;
  const cplx* M = piv + kc_idx;
  // this thread's rows in sweep order: top-down 0 .. mid-1, bottom-up P-1 .. mid
  const int nrow = up ? P - mid : mid;
  auto row = [&](int i) -> int64_t { return up ? (int64_t)P - 1 - i : (int64_t)i; };
  const int nb = (nrow + PD_IT - 1) / PD_IT;
  // ---- elimination towards the middle: d_q = (rhs_q - off d_prev) m_q,  rhs_q = R[q] - off_below F[q+1]
  auto issue_fwd = [&](int b) {
    if (b < nb && active) {
      cplx* slot = ring + (int64_t)((b % PD_IRING) * (PD_IT * 3)) * NT;
#pragma unroll
      for (int i = 0; i < PD_IT; ++i) {
        const int64_t q = row(min(b * PD_IT + i, nrow - 1));
        cp_async16(slot + (i * 3 + 0) * NT, R + (q * 2) * K);
        cp_async16(slot + (i * 3 + 1) * NT, F + ((q + 1) * 2) * K);
        cp_async16(slot + (i * 3 + 2) * NT, M + q * K);
      }
    }
    cp_async_commit();  // (an empty group keeps the group count in step)
  };
  cplx d = cmake(0, 0), clast = cmake(0, 0);
  for (int b = 0; b < PD_IRING - 1; ++b) issue_fwd(b);
  for (int b = 0; b < nb; ++b) {
    issue_fwd(b + PD_IRING - 1);
    cp_async_wait<PD_IRING - 1>();
    const cplx* slot = ring + (int64_t)((b % PD_IRING) * (PD_IT * 3)) * NT;
    cplx g[PD_IT], m[PD_IT];
#pragma unroll
    for (int i = 0; i < PD_IT; ++i) {
      g[i] = cfms(below.off, slot[(i * 3 + 1) * NT], slot[(i * 3 + 0) * NT]);
      m[i] = slot[(i * 3 + 2) * NT];
    }
#pragma unroll
    for (int i = 0; i < PD_IT; ++i) {
      const int r = b * PD_IT + i;
      if (r < nrow && active) {
        d = cmul(cfms(s.off, d, g[i]), m[i]);
        clast = cmul(s.off, m[i]);
        R[(row(r) * 2) * K] = d;
      }
    }
  }
  cp_async_wait<0>();
  // ---- the meeting: x_a = d_a - c_a x_b (a = mid-1, top sweep),  x_b = e_b - c_b x_a (b = mid, bottom sweep)
  xch[threadIdx.x][0] = d;       // (a sweep without rows contributes d = 0, c = 0)
  xch[threadIdx.x][1] = clast;
  __threadfence_block();
  __syncthreads();
  const cplx od = xch[threadIdx.x ^ 64][0], oc = xch[threadIdx.x ^ 64][1];
  cplx z;
  {
    const cplx da = up ? od : d, ca = up ? oc : clast, eb = up ? d : od, cb = up ? clast : oc;
    const cplx xa = cmul(cfms(ca, eb, da), crcp(cfms(ca, cb, cmake(1, 0))));   // (d_a - c_a e_b) / (1 - c_a c_b)
    z = up ? cfms(cb, xa, eb) : xa;
  }
  // ---- substitution away from the middle: x_q = d_q - c_q x_next, this sweep's rows in reverse order.  The values
  // d_q just stored are read back through L2 by this same thread.
  __threadfence();
  if (active && nrow > 0) R[(row(nrow - 1) * 2) * K] = z;
  const int nrb = nrow - 1;  // rows still to substitute: sweep indices nrow-2 .. 0
  const int nbb = (nrb + PD_IT - 1) / PD_IT;
  auto issue_bwd = [&](int b) {
    if (b < nbb && active) {
      cplx* slot = ring + (int64_t)((b % PD_IRING) * (PD_IT * 3)) * NT;
#pragma unroll
      for (int i = 0; i < PD_IT; ++i) {
        const int64_t q = row(max(nrow - 2 - (b * PD_IT + i), 0));
        cp_async16(slot + (i * 3 + 0) * NT, R + (q * 2) * K);
        cp_async16(slot + (i * 3 + 2) * NT, M + q * K);
      }
    }
    cp_async_commit();
  };
  for (int b = 0; b < PD_IRING - 1; ++b) issue_bwd(b);
  for (int b = 0; b < nbb; ++b) {
    issue_bwd(b + PD_IRING - 1);
    cp_async_wait<PD_IRING - 1>();
    const cplx* slot = ring + (int64_t)((b % PD_IRING) * (PD_IT * 3)) * NT;
#pragma unroll
    for (int i = 0; i < PD_IT; ++i) {
      const int r = nrow - 2 - (b * PD_IT + i);
      if (r >= 0 && active) {
        z = cfms(cmul(s.off, slot[(i * 3 + 2) * NT]), z, slot[(i * 3 + 0) * NT]);
        R[(row(r) * 2) * K] = z;
      }
    }
  }
  cp_async_wait<0>();
  if (PUSH) {
    // the slab functionals (see slab_functionals), split between the two sweeps: the top-down thread ends with the
    // FIRST interface value (first entry of the slab-local solve, rotated right-hand side of the separator row), the
    // bottom-up thread with the LAST one (last entry of the slab-local solve)
    if (P == 1) {  // a single interface row belongs to the bottom-up sweep: hand it to the top-down thread
      xch[threadIdx.x][0] = z;
      __syncthreads();
      if (!up) z = xch[threadIdx.x ^ 64][0];
    }
    if (active) {
      const int Llast = sp.m - P * (PD_L + 1);
      VRec v;
      v.init(kc.a, kc.sh, cmake(0, 0));
      cplx rvL = cmake(0, 0), rvLl = cmake(0, 0);
      if (!v.diag) {
        for (int i = 1; i <= PD_L; ++i) {
          v.step();
          if (i == Llast) rvLl = cscale(crcp(v.V), v.one);
        }
        rvL = cscale(crcp(v.V), v.one);
      }
      const int64_t slot = (((int64_t)(ep & 1ull) * cm.G + cm.rank) * 6 + rhs) * cm.kmax + kk;
      if (!up) {
        const cplx fv = cfma(z, rvL, lv.F[0][(int64_t)rhs * K + kk]);
        cplx sP = cmake(0, 0), sM = cmake(0, 0);
        if (!sp.first_dirichlet) {
          if (sp.al) rotate_in<true>(kc, w[kk], w[sp.plane + kk], sP, sM);
          else rotate_in<false>(kc, w[kk], w[sp.plane + kk], sP, sM);
        }
        for (int p = 0; p < cm.G; ++p) {
          cplx* g = cm.peer_gath[p] + slot;
          g[0] = fv; g[4 * cm.kmax] = rhs ? sM : sP;
        }
      } else {
        const cplx lvv = Llast > 0 ? cfma(z, rvLl, sl.lastl[(int64_t)rhs * K + kk]) : z;
        for (int p = 0; p < cm.G; ++p) cm.peer_gath[p][slot + 2 * cm.kmax] = lvv;
      }
    }
    slab_publish(cm, ep, k0, min(PD_ITK, sp.kend - k0));
  }
}

// ------------------------------------------------------------------- pass B
template <bool SLAB, bool AL>
__global__ void __launch_bounds__(PD_KB)
pd_solve_passB_kernel(cplx* __restrict__ w, const cplx* __restrict__ zsep, SolveParams sp, SlabPtrs sl) {
  pd_pdl_enter((sp.pdl_early & 8) != 0);  // programmatic dependent launch: see pd_common.cuh
  __shared__ cplx mtab[PD_L][PD_KB];
  const int tid = threadIdx.x;
  const int kk = sp.koff + blockIdx.x * PD_KB + tid;
  const bool valid = kk < sp.kend;
  const int kc_idx = valid ? kk : sp.kend - 1;
  const KCoef kc = make_coef<AL>(freq_of(sp, kc_idx), sp);
  fill_pivots<PD_KB>(kc, mtab, tid);
  cplx* wu = w + kc_idx;
  cplx* wp = w + sp.plane + kc_idx;
  const cplx zero = cmake(0, 0);
  const int P = sp.rows[1], Llast = sp.m - P * (PD_L + 1);
  // slab mode: the outer separator values (left / right neighbour slab).  Their effect on the INTERIOR separators
  // of this slab has already been folded into zsep by pd_slab_global_kernel (interface Green's vectors), so the
  // chunk loop below is the single-GPU one.
  cplx oLP = zero, oLM = zero, oRP = zero, oRM = zero;
  if (SLAB) {
    oLP = sl.zout[kc_idx]; oLM = sl.zout[sp.K + kc_idx];
    oRP = sl.zout[2 * (int64_t)sp.K + kc_idx]; oRM = sl.zout[3 * (int64_t)sp.K + kc_idx];
  }
  for (int c = sp.c0 + blockIdx.y; c < sp.c1; c += gridDim.y) {
    const int Lc = c < P ? PD_L : Llast;
    const int j0 = c * (PD_L + 1) + 1;
    cplx dP[PD_L], dM[PD_L];
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        dP[i] = wu[(int64_t)(j0 + i) * sp.K];
        dM[i] = wp[(int64_t)(j0 + i) * sp.K];
      }
    }
    cplx zlP = oLP, zlM = oLM, zrP = oRP, zrM = oRM;  // zero in single-GPU mode
    if (c > 0) {
      const int64_t o = ((int64_t)(c - 1) * 2) * sp.K + kc_idx;
      zlP = zsep[o];
      zlM = zsep[o + sp.K];
    }
    if (c < P) {
      const int64_t o = ((int64_t)c * 2) * sp.K + kc_idx;
      zrP = zsep[o];
      zrM = zsep[o + sp.K];
    }
    // forward elimination (in place: d_i overwrites rho_i)
    cplx pP = zlP, pM = zlM;  // "d_{-1}" = known left neighbour value
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        cplx rP, rM;
        rotate_in<AL>(kc, dP[i], dM[i], rP, rM);
        if (i == Lc - 1) {
          rP = cfms(kc.a, zrP, rP);
          rM = cfms(kc.a, zrM, rM);
        }
        const cplx mi = mtab[i][tid];
        pP = cmul(cfms(kc.a, pP, rP), mi);
        pM = cmul(cfms(kc.a, pM, rM), mi);
        dP[i] = pP;
        dM[i] = pM;
      }
    }
    // back substitution, rotation and store
    cplx nP = zero, nM = zero;
#pragma unroll
    for (int i = PD_L - 1; i >= 0; --i) {
      if (i < Lc) {
        if (i < Lc - 1) {
          const cplx cp = cmul(kc.a, mtab[i][tid]);
          nP = cfms(cp, nP, dP[i]);
          nM = cfms(cp, nM, dM[i]);
        } else {
          nP = dP[i];
          nM = dM[i];
        }
        cplx ou, op;
        rotate_out<AL>(kc, nP, nM, ou, op);
        if (valid) {
          wu[(int64_t)(j0 + i) * sp.K] = ou;
          wp[(int64_t)(j0 + i) * sp.K] = op;
        }
      }
    }
    if (valid && c < P) {  // the separator row that follows this chunk
      cplx ou, op;
      rotate_out<AL>(kc, zrP, zrM, ou, op);
      wu[(int64_t)(j0 + PD_L) * sp.K] = ou;
      wp[(int64_t)(j0 + PD_L) * sp.K] = op;
    }
    // Dirichlet rows: output exactly 0 (:482, bcs :44-45); in slab mode row 0 may be the
    // inter-slab separator this rank owns
    if (valid && c == 0) {
      cplx ou = zero, op = zero;
      if (SLAB && !sp.first_dirichlet) rotate_out<AL>(kc, oLP, oLM, ou, op);
      wu[0] = ou;
      wp[0] = op;
    }
    if (valid && c == P && sp.last_dirichlet) {
      wu[(int64_t)(sp.m + 1) * sp.K] = zero;
      wp[(int64_t)(sp.m + 1) * sp.K] = zero;
    }
  }
  // peer-store exchange, single-stream apply: this apply's functionals and separator kernels are complete (stream
  // order), the next apply's have not started -> the apply is marked complete here, no extra launch
  if (SLAB && sl.epoch_bump && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *sl.epoch_bump += 1ull;
}

// The same as a launch of its own, for the two-stream variant of the apply (after the join of both halves).
__global__ void pd_slab_epoch_bump_kernel(unsigned long long* epoch, int pdl_early) {
  pd_pdl_enter((pdl_early & 16) != 0);  // programmatic dependent launch: see pd_common.cuh
  *epoch += 1ull;
}

// ------------------------------------------------------------ slab-mode kernels
// Slab mode = x-slab sharding kept through the solve (no transposes): rank r owns a contiguous node
// slab for ALL frequencies; the first node of every slab r > 0 is a global separator.  Each slab body is
// a pure Toeplitz block T_{m_r}; the G-1 separators of one frequency form a tiny tridiagonal system
//   a rv_{r-1} z_{r-1} + a (eta - dvv_{r-1} - dvv_r) z_r + a rv_r z_{r+1} = rho_r - a (l_{r-1} + f_r)
// with rv_s = 1/V_{m_s}, dvv_s = (V_{m_s} - V_{m_s - 1}) / V_{m_s} (cancellation-free, see Sys) and
// f_s / l_s the first / last entry of the slab-local solve with zero neighbours.

// out[6][K] = (f+, f-, l+, l-, rho_sep+, rho_sep-) of this slab
template <bool PUSH>
__global__ void __launch_bounds__(PD_KB)
pd_slab_functionals_kernel(const cplx* __restrict__ w, Levels lv, SolveParams sp, SlabPtrs sl,
                           cplx* __restrict__ out, SlabCommDev cm) {
  pd_pdl_enter((sp.pdl_early & 4) != 0);  // programmatic dependent launch: see pd_common.cuh
  const int kk = sp.koff + blockIdx.x * PD_KB + threadIdx.x;
  const unsigned long long ep = PUSH ? *cm.epoch + 1ull : 0ull;
  if (kk < sp.kend) {
    const int64_t K = sp.K;
    const KCoef kc = make_coef(freq_of(sp, kk), sp);
    const int P = sp.rows[1];
    const cplx zero = cmake(0, 0);
    cplx z0P = zero, z0M = zero, zeP = zero, zeM = zero;
    if (P > 0) {
      z0P = lv.R[1][kk]; z0M = lv.R[1][K + kk];
      zeP = lv.R[1][((int64_t)(P - 1) * 2) * K + kk]; zeM = lv.R[1][((int64_t)(P - 1) * 2 + 1) * K + kk];
    }
    cplx fP, fM, lP, lM, sP, sM;
    slab_functionals(kc, sp, w, kk, lv.F[0] ? lv.F[0][kk] : zero, lv.F[0] ? lv.F[0][K + kk] : zero, sl.lastl[kk],
                     sl.lastl[K + kk], z0P, z0M, zeP, zeM, fP, fM, lP, lM, sP, sM);
    if (!PUSH) {
      out[kk] = fP; out[K + kk] = fM; out[2 * K + kk] = lP; out[3 * K + kk] = lM;
      out[4 * K + kk] = sP; out[5 * K + kk] = sM;
    } else {
      slab_push(cm, ep, kk, fP, fM, lP, lM, sP, sM);
    }
  }
  if (PUSH) slab_publish(cm, ep, sp.koff + blockIdx.x * PD_KB, min(PD_KB, sp.kend - sp.koff - blockIdx.x * PD_KB));
}

struct SlabGeom {
  int G, rank;
  int body[PD_MAX_SLABS];  // m_s of every slab
};

// Per-slab Green's coefficients rv_s = 1/V_{m_s}, dvv_s = (V_{m_s} - V_{m_s-1})/V_{m_s}: a recurrence of
// m_s steps per frequency, independent of the right-hand side -> computed once when the plan is built
// (coef[G][2][K], a few hundred KB), not per apply.  One thread per (k, slab).
__global__ void __launch_bounds__(PD_KB)
pd_slab_coef_kernel(SolveParams sp, SlabGeom sg, cplx* __restrict__ coef) {
  const int kk = blockIdx.x * PD_KB + threadIdx.x;
  const int s = blockIdx.y;
  if (kk >= sp.K) return;
  const KCoef kc = make_coef(freq_of(sp, kk), sp);
  VRec v;
  v.init(kc.a, kc.sh, cmake(0, 0));
  cplx rv = cmake(0, 0), dvv = cmake(1, 0);  // decoupled: V_{m-1}/V_m -> 0
  if (!v.diag) {
    for (int i = 1; i <= sg.body[s]; ++i) v.step();
    const cplx r = crcp(v.V);
    rv = cscale(r, v.one);
    dvv = cmul(cmake(v.one + v.E.x, v.E.y), r);
  }
  coef[((int64_t)s * 2) * sp.K + kk] = rv;
  coef[((int64_t)s * 2 + 1) * sp.K + kk] = dvv;
}

// gathered[G][6][K] -> zout[4][K] (left+, left-, right+, right-) of this rank
// grid (frequency blocks, y): every CTA solves the tiny separator system of its frequencies (redundantly in y),
// y = 0 stores the outer separator values zout, and all CTAs together add their effect on this slab's INTERIOR
// separators, zsep[c] += cL g0[c] + cR g1[c] (g0, g1: interface Green's vectors of the plan; cL/cR = -(a / V_L) z_left,
// -(a / V_Llast) z_right), so that pass B needs no slab-specific arithmetic in its chunk loop.
// GT: the slab count as a compile-time constant (2, 4, 8: the tiny Thomas arrays live in registers) or 0 (any G <= 16)
template <bool WAIT, int GT>
__global__ void __launch_bounds__(PD_KB)
pd_slab_global_kernel(const cplx* __restrict__ gathered, int64_t gstride, SolveParams sp, SlabGeom sg,
                      const cplx* __restrict__ coef, cplx* __restrict__ zout, SlabCommDev cm,
                      cplx* __restrict__ zsep, const cplx* __restrict__ green) {
  pd_pdl_enter((sp.pdl_early & 16) != 0);  // programmatic dependent launch: see pd_common.cuh
  // The correction sweep at the end touches rows of zsep / green that depend on nothing the peers send: its first
  // batch of rows is loaded HERE, so that the loads are in flight under the flag wait and the tiny separator solve
  // (the sweep used to be half of this kernel's time, one exposed DRAM round trip per row).
  constexpr int UNR = 4;
  const int kk = sp.koff + blockIdx.x * PD_KB + threadIdx.x;
  const int64_t K = sp.K;
  const int P = sp.rows[1];
  const bool sweep = P > 0 && zsep != nullptr && kk < sp.kend;
  cplx bg0[UNR], bg1[UNR], bzP[UNR], bzM[UNR];
  if (sweep) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int c = blockIdx.y + u * gridDim.y;
      if (c < P) {
        const int64_t o = ((int64_t)c * 2) * K + kk;
        bg0[u] = green[o]; bg1[u] = green[o + K];
        bzP[u] = zsep[o]; bzM[u] = zsep[o + K];
      }
    }
  }
  if (WAIT) {
    // wait for the functionals of THIS frequency block from every rank (bounded spin, see SlabCommDev)
    const unsigned long long ep = *cm.epoch + 1ull;
    slab_wait(cm, ep, sp.koff + blockIdx.x * PD_KB, min(PD_KB, sp.kend - sp.koff - blockIdx.x * PD_KB));
    gathered = cm.peer_gath[cm.rank] + (int64_t)(ep & 1ull) * cm.G * 6 * cm.kmax;
  }
  if (kk >= sp.kend) return;
  const int64_t GS = gstride;
  const KCoef kc = make_coef(freq_of(sp, kk), sp);
  constexpr int GA = GT ? GT : PD_MAX_SLABS;
  const int G = GT ? GT : sg.G;
  cplx rv[GA], dvv[GA];
#pragma unroll
  for (int s = 0; s < G; ++s) {
    rv[s] = coef[((int64_t)s * 2) * K + kk];
    dvv[s] = coef[((int64_t)s * 2 + 1) * K + kk];
  }
  // separators 1..G-1 -> rows 0..G-2; Thomas with two right-hand sides
  const cplx roff = crcp(kc.a);
  VRec t;
  t.init(kc.a, kc.sh, cmake(0, 0));
  const cplx eta = t.diag ? cmake(0, 0) : t.eta;
  cplx cp[GA], dP[GA], dM[GA];
  cplx prevc = cmake(0, 0), pP = cmake(0, 0), pM = cmake(0, 0);
#pragma unroll
  for (int r = 1; r < G; ++r) {
    const cplx* gl = gathered + ((int64_t)(r - 1) * 6) * GS + kk;  // slab left of separator r
    const cplx* gr = gathered + ((int64_t)r * 6) * GS + kk;        // slab right of it (owns the separator)
    cplx di, lo, up;
    if (t.diag) {
      di = cmake(kc.sh.x - 2.0 * kc.a.x, kc.sh.y - 2.0 * kc.a.y);
      lo = up = cmake(0, 0);
    } else {
      di = cmul(kc.a, csub(csub(eta, dvv[r - 1]), dvv[r]));
      lo = r > 1 ? cmul(kc.a, rv[r - 1]) : cmake(0, 0);
      up = r < G - 1 ? cmul(kc.a, rv[r]) : cmake(0, 0);
    }
    // (__ldcg: the peers' stores land in L2; never read these through L1)
    cplx rP = cfms(kc.a, cadd(__ldcg(gl + 2 * GS), __ldcg(gr)), __ldcg(gr + 4 * GS));
    cplx rM = cfms(kc.a, cadd(__ldcg(gl + 3 * GS), __ldcg(gr + GS)), __ldcg(gr + 5 * GS));
    const cplx inv = crcp(cfms(lo, prevc, di));
    prevc = cmul(up, inv);
    pP = cmul(cfms(lo, pP, rP), inv);
    pM = cmul(cfms(lo, pM, rM), inv);
    cp[r] = prevc; dP[r] = pP; dM[r] = pM;
  }
  (void)roff;
  cplx zP = cmake(0, 0), zM = cmake(0, 0);
  cplx leftP = cmake(0, 0), leftM = cmake(0, 0), rightP = cmake(0, 0), rightM = cmake(0, 0);
#pragma unroll
  for (int r = G - 1; r >= 1; --r) {
    zP = cfms(cp[r], zP, dP[r]);
    zM = cfms(cp[r], zM, dM[r]);
    if (r == sg.rank) { leftP = zP; leftM = zM; }
    if (r == sg.rank + 1) { rightP = zP; rightM = zM; }
  }
  if (blockIdx.y == 0) {
    zout[kk] = leftP; zout[K + kk] = leftM; zout[2 * K + kk] = rightP; zout[3 * K + kk] = rightM;
  }
  if (sweep) {
    const int Llast = sp.m - P * (PD_L + 1);
    VRec v;
    v.init(kc.a, kc.sh, cmake(0, 0));
    cplx gl = cmake(0, 0), gr = kc.a;  // a / V_L and a / V_Llast (V_0 = 1)
    if (!v.diag) {
      for (int i = 1; i <= PD_L; ++i) {
        v.step();
        if (i == Llast) gr = cmul(kc.a, cscale(crcp(v.V), v.one));
      }
      gl = cmul(kc.a, cscale(crcp(v.V), v.one));
    } else if (Llast > 0) {
      gr = cmake(0, 0);
    }
    const cplx cLP = cneg(cmul(gl, leftP)), cLM = cneg(cmul(gl, leftM));
    const cplx cRP = cneg(cmul(gr, rightP)), cRM = cneg(cmul(gr, rightM));
    // batches of UNR rows: all loads of a batch before its first store (the first batch is already in registers)
    for (int cb = blockIdx.y; cb < P; cb += gridDim.y * UNR) {
      if (cb != (int)blockIdx.y) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int c = cb + u * gridDim.y;
          if (c < P) {
            const int64_t o = ((int64_t)c * 2) * K + kk;
            bg0[u] = green[o]; bg1[u] = green[o + K];
            bzP[u] = zsep[o]; bzM[u] = zsep[o + K];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int c = cb + u * gridDim.y;
        if (c < P) {
          const int64_t o = ((int64_t)c * 2) * K + kk;
          zsep[o] = cfma(cLP, bg0[u], cfma(cRP, bg1[u], bzP[u]));
          zsep[o + K] = cfma(cLM, bg0[u], cfma(cRM, bg1[u], bzM[u]));
        }
      }
    }
  }
}

// R1 <- e_0 in slot 0, e_{P-1} in slot 1 (right-hand sides of the two interface Green's vectors)
__global__ void pd_slab_unit_rhs_kernel(cplx* __restrict__ R1, int P, int64_t K) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)P * 2 * K) return;
  const int64_t q = e / (2 * K);
  const int slot = (int)((e / K) & 1);
  R1[e] = ((slot == 0 && q == 0) || (slot == 1 && q == P - 1)) ? cmake(1, 0) : cmake(0, 0);
}

// --------------------------------------------------------------- host side
struct SolvePlan {
  int nlev;
  int rows[PD_MAX_LEVELS];
  cplx* R[PD_MAX_LEVELS];
  cplx* F[PD_MAX_LEVELS];
  // slab mode
  cplx* lastl;
  cplx* green;
  cplx* zout;
  cplx* slabcoef;
  cplx* green_h;     // the same two plan-time tables for the half spectrum of the real-input path
  cplx* slabcoef_h;  //   (columns 0 .. N_t/2 in natural order, row stride Kp)
  SlabGeom sg;
  cplx* ipiv[2];     // interface pivots [rows[1]][K] for the full ([0]) and the half spectrum ([1]); null: multi-level
  // peer-store exchange (pd_slab_comm_*): null / 0 until created
  void* comm_base;                    // this rank's symmetric buffer (gathered + flags), one allocation
  size_t comm_bytes;
  int64_t comm_kmax;
  int comm_nflag;
  void* comm_peer[PD_MAX_SLABS];      // mapped base of every rank's buffer
  bool comm_ipc[PD_MAX_SLABS];        // opened with cudaIpcOpenMemHandle (to be closed)
  bool comm_connected;
  unsigned long long* comm_epoch;     // device: completed applies
  int* comm_err;                      // device: a bounded wait expired
};

static SolvePlan* plan_of(pd_handle* h) { return reinterpret_cast<SolvePlan*>(h->solve_plan); }

void pd_solve_free(pd_handle* h) {
  SolvePlan* pl = plan_of(h);
  if (!pl) return;
  for (int l = 0; l < PD_MAX_LEVELS; ++l) {
    if (pl->R[l]) cudaFree(pl->R[l]);
    if (pl->F[l]) cudaFree(pl->F[l]);
  }
  if (pl->ipiv[0]) cudaFree(pl->ipiv[0]);
  if (pl->ipiv[1]) cudaFree(pl->ipiv[1]);
  if (pl->green_h) cudaFree(pl->green_h);
  if (pl->slabcoef_h) cudaFree(pl->slabcoef_h);
  if (pl->lastl) cudaFree(pl->lastl);
  if (pl->green) cudaFree(pl->green);
  if (pl->zout) cudaFree(pl->zout);
  if (pl->slabcoef) cudaFree(pl->slabcoef);
  for (int r = 0; r < PD_MAX_SLABS; ++r)
    if (pl->comm_ipc[r] && pl->comm_peer[r]) cudaIpcCloseMemHandle(pl->comm_peer[r]);
  if (pl->comm_base) cudaFree(pl->comm_base);
  if (pl->comm_epoch) cudaFree(pl->comm_epoch);
  if (pl->comm_err) cudaFree(pl->comm_err);
  delete pl;
  h->solve_plan = nullptr;
}

// cps: CTAs of the kernel that are resident per SM (pass A 4, pass B 2, the small level kernels 8)
static dim3 stream_grid(const pd_handle* h, int K, int nchunks, int cps = 8) {
  const int kblocks = (K + PD_KB - 1) / PD_KB;
  // enough CTAs for ~8 resident per SM; more chunks than that are looped over
  int ny = (h->num_sms * 8 + kblocks - 1) / kblocks;
  // Short x-ranges (the x-slab of a multi-GPU run: 121 chunks at cfg3 on 8 GPUs): with that grid a CTA would see 3
  // chunks, and its prologue (the pivot recurrences, ~a chunk's worth of time) plus a ragged last wave cost 25 %.
  // Use whole waves of resident CTAs with at least ~8 chunks each instead.
  if (nchunks < ny * 8) {
    int per_wave = h->num_sms * cps / kblocks;
    if (per_wave < 1) per_wave = 1;
    int waves = nchunks / 8 / per_wave;
    if (waves < 1) waves = 1;
    if (per_wave * waves < ny) ny = per_wave * waves;
  }
  if (ny > nchunks) ny = nchunks;
  if (ny < 1) ny = 1;
  if (ny > 65535) ny = 65535;
  return dim3(kblocks, ny);
}

static void fill_params(pd_handle* h, SolveParams& sp, Levels& lv, SlabPtrs& sl, int half_spectrum = 0) {
  SolvePlan* pl = plan_of(h);
  memset(&sp, 0, sizeof(sp));
  sp.n = h->n; sp.m = h->m; sp.K = h->kcount; sp.kbegin = h->kbegin; sp.N_t = h->cfg.N_t;
  if (half_spectrum) {  // real-input path: frequencies 0 .. N_t/2 in natural order
    sp.K = (h->cfg.N_t / 2 + 1 + 7) & ~7;  // padded to 128-byte rows; the padding columns hold zeros
    sp.kbegin = 0;
  }
  sp.h = h->h; sp.dt2 = h->dt * h->dt; sp.c = h->c;
  sp.plane = (int64_t)h->n * sp.K;
  sp.koff = 0;
  sp.kend = sp.K;
  sp.c0 = 0;
  sp.c1 = pl->rows[1] + 1;
  sp.nlev = pl->nlev;
  sp.first_dirichlet = h->slab_count <= 1 || h->slab_rank == 0;
  sp.last_dirichlet = h->slab_count <= 1 || h->slab_rank == h->slab_count - 1;
  sp.freq_perm = h->cfg.N_t == 16384 && !half_spectrum;
  sp.al = h->cfg.alpha != 1.0;
  sp.lna = log(h->cfg.alpha) / (double)h->cfg.N_t;
  sp.pdl_early = h->pdl ? h->pdl_early : 0;
  for (int l = 0; l < PD_MAX_LEVELS; ++l) sp.rows[l] = pl->rows[l];
  for (int l = 0; l < PD_MAX_LEVELS; ++l) { lv.R[l] = pl->R[l]; lv.F[l] = pl->F[l]; }
  sl.lastl = pl->lastl; sl.green = half_spectrum ? pl->green_h : pl->green; sl.zout = pl->zout;
  sl.epoch_bump = nullptr;
}

void pd_solve_fill_params(pd_handle* h, SolveParams& sp, Levels& lv, SlabPtrs& sl, int half_spectrum) {
  fill_params(h, sp, lv, sl, half_spectrum);
}

// levels 1..top: reduce, PCR on the top system, back-substitute; leaves the level-1 solution in R[1]
struct PushCtx {
  const cplx* w;
  SlabPtrs sl;
  SlabCommDev cm;
};
// push != nullptr and a single-level interface: the PCR kernel also forms and pushes the slab functionals
// (returns with *pushed = true); otherwise the caller launches pd_slab_functionals_kernel
static int run_interface(pd_handle* h, const SolveParams& sp, const Levels& lv, cudaStream_t st,
                         const PushCtx* push = nullptr, bool* pushed = nullptr) {
  const int top = sp.nlev;
  const int ncol = sp.kend - sp.koff;
  const cplx* piv = plan_of(h)->ipiv[sp.K == h->kcount ? 0 : 1];
  if (piv) {
    const int nblk = (ncol + PD_ITK - 1) / PD_ITK;
    SlabPtrs nosl;
    SlabCommDev nocm;
    memset(&nosl, 0, sizeof(nosl));
    memset(&nocm, 0, sizeof(nocm));
    const int ring = nblk <= h->num_sms ? 8 : (nblk <= 2 * h->num_sms ? 4 : 2);
#define PD_IFACE_LAUNCH(RING)                                                                                       \
  do {                                                                                                              \
    if (push)                                                                                                       \
      PD_KLAUNCH((pd_solve_iface_thomas_kernel<true, RING>), nblk, 4 * PD_ITK, PD_ISMEM(RING), st, lv, sp, piv, push->w,      \
                                                                                       push->sl, push->cm);         \
    else                                                                                                            \
      PD_KLAUNCH((pd_solve_iface_thomas_kernel<false, RING>), nblk, 4 * PD_ITK, PD_ISMEM(RING), st, lv, sp, piv, nullptr,     \
                                                                                        nosl, nocm);                \
  } while (0)
    if (ring == 8) PD_IFACE_LAUNCH(8);
    else if (ring == 4) PD_IFACE_LAUNCH(4);
    else PD_IFACE_LAUNCH(2);
    if (push && pushed) *pushed = true;
    PD_CHECK_LAUNCH();
    h->launches++;
    return PD_OK;
  }
  for (int lev = 1; lev < top; ++lev) {
    PD_KLAUNCH(pd_solve_level_reduce_kernel, stream_grid(h, ncol, sp.rows[lev + 1] + 1), PD_KB, 0, st, lv, sp, lev);
    PD_CHECK_LAUNCH();
    h->launches++;
  }
  const int n = sp.rows[top];
  int kpb = (PD_PCR_THREADS * PD_PCR_MAXROWS) / n;
  if (kpb > 32) kpb = 32;
  // keep at least ~2 CTAs per SM when the frequency count allows it
  while (kpb > 4 && (ncol + kpb - 1) / kpb < 2 * h->num_sms) kpb >>= 1;
  if (kpb < 1) kpb = 1;
  const size_t smem = (size_t)n * kpb * 64;
  const int nblk = (ncol + kpb - 1) / kpb;
  if (push && top == 1 && kpb % PD_FKB == 0) {
    PD_KLAUNCH((pd_solve_pcr_kernel<true>), nblk, PD_PCR_THREADS, smem, st, lv, sp, top, kpb, push->w, push->sl, push->cm);
    if (pushed) *pushed = true;
  } else {
    SlabPtrs nosl;
    SlabCommDev nocm;
    memset(&nosl, 0, sizeof(nosl));
    memset(&nocm, 0, sizeof(nocm));
    PD_KLAUNCH((pd_solve_pcr_kernel<false>), nblk, PD_PCR_THREADS, smem, st, lv, sp, top, kpb, nullptr, nosl, nocm);
  }
  PD_CHECK_LAUNCH();
  h->launches++;
  for (int lev = top - 1; lev >= 1; --lev) {
    PD_KLAUNCH(pd_solve_level_back_kernel, stream_grid(h, ncol, sp.rows[lev + 1] + 1), PD_KB, 0, st, lv, sp, lev);
    PD_CHECK_LAUNCH();
    h->launches++;
  }
  return PD_OK;
}

int pd_solve_plan(pd_handle* h) {
  // largest dynamic shared memory the PCR kernel is ever launched with (set once, not per launch)
  PD_CUDA(cudaFuncSetAttribute(pd_solve_pcr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               PD_PCR_THREADS * PD_PCR_MAXROWS * 64));
  PD_CUDA(cudaFuncSetAttribute(pd_solve_pcr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               PD_PCR_THREADS * PD_PCR_MAXROWS * 64));
#define PD_IFACE_ATTR(RING)                                                                                         \
  PD_CUDA(cudaFuncSetAttribute(pd_solve_iface_thomas_kernel<false, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               PD_ISMEM(RING)));                                                                     \
  PD_CUDA(cudaFuncSetAttribute(pd_solve_iface_thomas_kernel<true, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                               PD_ISMEM(RING)))
  PD_IFACE_ATTR(8);
  PD_IFACE_ATTR(4);
  PD_IFACE_ATTR(2);
  SolvePlan* pl = new SolvePlan();
  memset(pl, 0, sizeof(*pl));
  h->solve_plan = pl;
  h->L = PD_L;
  const size_t K = (size_t)h->kcount;
  // rows[0] = m; reduce while the interface is too large for the PCR kernel (few frequencies: up to 128 rows go
  // to PCR and save launches).  PD_PCR_MAX overrides (experiments; PCR of 120 rows at K = 4096 measured 0.23 ms
  // against 0.05 ms for reduce / PCR(13) / back, and is less accurate).
  int pcr_max = K <= 2048 ? PD_PCR_MAX_SMALLK : PD_PCR_MAX;
  if (const char* e = getenv("PD_PCR_MAX")) {
    const int v = atoi(e);
    if (v >= 1 && v <= PD_PCR_MAX_SMALLK) pcr_max = v;
  }
  // interface systems of up to this many rows go to the one-launch sequential kernel (pd_solve_iface_thomas_kernel);
  // PD_ITHOMAS_MAX overrides (0 = never)
  // Measured on B200 (tools/scale_probe.py, K = 4096, interface rows 120 / 240 / 481 / 963 = N_x 2048 ... 16384):
  //   reduce / PCR / back chain (3-5 launches)   52 /  83 / 137 / 205 us   (+ 24 us for the slab functionals launch)
  //   sequential kernel, cp.async ring           30 /  50 /  98 / 169 us   (functionals and peer stores included)
  //   (its first versions: one register batch in flight 45 / 88 / 194 / 412 us, two batches 39 / 73 / 150 / 308 us;
  //    a single PCR launch on 120 rows: 233 us)
  // so the sequential kernel is the default whenever the interface has more than one PCR launch's worth of rows;
  // PD_ITHOMAS_MAX overrides (0 = never).
  h->iface_thomas_max = 8192;
  if (const char* e = getenv("PD_ITHOMAS_MAX")) h->iface_thomas_max = atoi(e);
  if (h->kcount != h->cfg.N_t) h->iface_thomas_max = 0;  // frequency-sharded handles keep the multi-level path
  pl->rows[0] = h->m;
  int l = 0;
  while (true) {
    if (l + 1 >= PD_MAX_LEVELS) {
      pd_set_error("N_x = %d needs more than %d partition levels", h->cfg.N_x, PD_MAX_LEVELS);
      return PD_ERR_INVALID;
    }
    const int next = pl->rows[l] / (chunk_len(l) + 1);
    pl->rows[l + 1] = next;
    ++l;
    if (next <= pcr_max) break;
  }
  // l is the top level: solved by PCR when it has rows, absent when rows == 0
  pl->nlev = pl->rows[l] > 0 ? l : l - 1;
  h->P = pl->rows[1];
  h->Llast = h->m - h->P * (PD_L + 1);
  for (int lev = 0; lev <= pl->nlev; ++lev) {
    if (lev >= 1) {
      size_t bytes = sizeof(cplx) * (size_t)pl->rows[lev] * 2 * K;
      PD_CUDA(cudaMalloc(&pl->R[lev], bytes));
      h->ws_bytes += bytes;
    }
    if (lev < pl->nlev) {
      size_t bytes = sizeof(cplx) * (size_t)(pl->rows[lev + 1] + 1) * 2 * K;
      PD_CUDA(cudaMalloc(&pl->F[lev], bytes));
      h->ws_bytes += bytes;
    }
  }
  // one-launch sequential interface (pd_solve_iface_thomas_kernel): factorise the level-1 system once
  if (pl->nlev >= 1 && h->iface_thomas_max > 0 && pl->rows[1] <= h->iface_thomas_max &&
      (pl->rows[1] > PD_PCR_MAX || h->slab_count > 1 || getenv("PD_ITHOMAS_MAX"))) {
    const bool want_half = pd_rfft_supported(h);
    for (int half = 0; half <= (want_half ? 1 : 0); ++half) {
      SolveParams sp; Levels lv; SlabPtrs sl;
      fill_params(h, sp, lv, sl, half);
      if (half == 1 && sp.K == (int)K) break;  // (cannot tell the two apart by width: keep the multi-level path)
      const size_t bytes = sizeof(cplx) * (size_t)pl->rows[1] * (size_t)sp.K;
      PD_CUDA(cudaMalloc(&pl->ipiv[half], bytes));
      h->ws_bytes += bytes;
      pd_iface_pivots_kernel<<<(sp.K + PD_KB - 1) / PD_KB, PD_KB>>>(sp, pl->ipiv[half]);
      PD_CHECK_LAUNCH();
    }
  }
  if (h->slab_count > 1) {
    // geometry of every slab (balanced split of the n nodes, first `extra` slabs one node longer)
    const int G = h->slab_count, ntot = h->cfg.N_x + 1;
    if (G > PD_MAX_SLABS) {
      pd_set_error("slab_count %d exceeds %d", G, PD_MAX_SLABS);
      return PD_ERR_INVALID;
    }
    pl->sg.G = G;
    pl->sg.rank = h->slab_rank;
    for (int s = 0; s < G; ++s) {
      const int cnt = ntot / G + (s < ntot % G ? 1 : 0);
      pl->sg.body[s] = cnt - 1 - (s == G - 1 ? 1 : 0);
      if (pl->sg.body[s] < 1) {
        pd_set_error("slab %d of %d has no interior rows (N_x = %d)", s, G, h->cfg.N_x);
        return PD_ERR_INVALID;
      }
    }
    PD_CUDA(cudaMalloc(&pl->lastl, sizeof(cplx) * 2 * K));
    PD_CUDA(cudaMalloc(&pl->zout, sizeof(cplx) * 4 * K));
    PD_CUDA(cudaMemset(pl->lastl, 0, sizeof(cplx) * 2 * K));
    PD_CUDA(cudaMemset(pl->zout, 0, sizeof(cplx) * 4 * K));
    h->ws_bytes += sizeof(cplx) * 6 * K;
    // plan-time tables (per-slab Green's coefficients, interface Green's vectors), for the full spectrum and,
    // when the real-input path exists for this N_t, for its half spectrum
    const bool want_half = pd_rfft_supported(h);
    for (int half = 0; half <= (want_half ? 1 : 0); ++half) {
      SolveParams sp; Levels lv; SlabPtrs sl;
      fill_params(h, sp, lv, sl, half);
      const size_t Kt = (size_t)sp.K;
      cplx** coef = half ? &pl->slabcoef_h : &pl->slabcoef;
      cplx** green = half ? &pl->green_h : &pl->green;
      PD_CUDA(cudaMalloc(coef, sizeof(cplx) * (size_t)G * 2 * Kt));
      h->ws_bytes += sizeof(cplx) * (size_t)G * 2 * Kt;
      pd_slab_coef_kernel<<<dim3((unsigned)((Kt + PD_KB - 1) / PD_KB), G), PD_KB>>>(sp, pl->sg, *coef);
      PD_CHECK_LAUNCH();
      if (pl->nlev >= 1) {
        // interface Green's vectors: one run of the interface levels on unit right-hand sides
        const int P = pl->rows[1];
        const size_t bytes = sizeof(cplx) * (size_t)P * 2 * Kt;
        PD_CUDA(cudaMalloc(green, bytes));
        h->ws_bytes += bytes;
        PD_CUDA(cudaMemset(pl->F[0], 0, sizeof(cplx) * (size_t)(P + 1) * 2 * Kt));
        const int64_t tot = (int64_t)P * 2 * Kt;
        pd_slab_unit_rhs_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(pl->R[1], P, (int64_t)Kt);
        PD_CHECK_LAUNCH();
        int rc = run_interface(h, sp, lv, 0);
        if (rc) return rc;
        PD_CUDA(cudaMemcpy(*green, pl->R[1], bytes, cudaMemcpyDeviceToDevice));
      }
    }
  }
  return PD_OK;
}

// one column range [koff, kend) through pass A, the interface levels and pass B on stream st
static int solve_range(pd_handle* h, cplx* w, SolveParams sp, const Levels& lv, const SlabPtrs& sl, int koff, int kend,
                       cudaStream_t st, cudaEvent_t* ev, cudaEvent_t after_passA = nullptr) {
  sp.koff = koff;
  sp.kend = kend;
  const dim3 gridA = stream_grid(h, kend - koff, sp.rows[1] + 1, 4);
  const dim3 grid0 = stream_grid(h, kend - koff, sp.rows[1] + 1, 2);
  if (sp.nlev >= 1) {
    if (sp.al)
      PD_KLAUNCH((pd_solve_passA_kernel<true>), gridA, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, nullptr);
    else
      PD_KLAUNCH((pd_solve_passA_kernel<false>), gridA, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, nullptr);
    PD_CHECK_LAUNCH();
    h->launches++;
    if (ev) cudaEventRecord(ev[0], st);
    if (after_passA) cudaEventRecord(after_passA, st);
    int rc = run_interface(h, sp, lv, st);
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[1], st);
  } else if (ev) {
    cudaEventRecord(ev[0], st);
    cudaEventRecord(ev[1], st);
  }
  if (sp.al)
    PD_KLAUNCH((pd_solve_passB_kernel<false, true>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  else
    PD_KLAUNCH((pd_solve_passB_kernel<false, false>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

int pd_solve_launch(pd_handle* h, cplx* w, cudaStream_t st, cudaEvent_t* ev, int half_spectrum) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, half_spectrum);
  // (Splitting the columns over two streams to hide the latency-bound interface kernels under the
  // streaming passes of the other half was tried and does not help: pass A / pass B occupy the whole
  // register file of every SM, so nothing else becomes resident until they drain.)
  return solve_range(h, w, sp, lv, sl, 0, sp.K, st, ev);
}

// ---- the three parts of the solve stage as separate launches over level-0 chunk ranges (the node-slab
// interleaved schedule of pd_pc_apply: pass A of a slab right after its inverse FFT, pass B right before its FFT)
int pd_solve_nchunks(const pd_handle* h) {
  const SolvePlan* pl = reinterpret_cast<const SolvePlan*>(h->solve_plan);
  return pl->nlev >= 1 ? pl->rows[1] + 1 : 0;
}
int pd_solve_passA_range(pd_handle* h, cplx* w, int c0, int c1, cudaStream_t st) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, 0);
  sp.c0 = c0; sp.c1 = c1;
  const dim3 grid0 = stream_grid(h, sp.K, c1 - c0, 4);
  if (sp.al)
    PD_KLAUNCH((pd_solve_passA_kernel<true>), grid0, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, nullptr);
  else
    PD_KLAUNCH((pd_solve_passA_kernel<false>), grid0, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, nullptr);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}
int pd_solve_interface(pd_handle* h, cudaStream_t st) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, 0);
  return run_interface(h, sp, lv, st);
}
int pd_solve_passB_range(pd_handle* h, cplx* w, int c0, int c1, cudaStream_t st) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, 0);
  sp.c0 = c0; sp.c1 = c1;
  const dim3 grid0 = stream_grid(h, sp.K, c1 - c0, 2);
  if (sp.al)
    PD_KLAUNCH((pd_solve_passB_kernel<false, true>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  else
    PD_KLAUNCH((pd_solve_passB_kernel<false, false>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

cplx* pd_slab_lastl(pd_handle* h) { return plan_of(h) ? plan_of(h)->lastl : nullptr; }

bool pd_slab_half_supported(const pd_handle* h) {
  const SolvePlan* pl = reinterpret_cast<const SolvePlan*>(h->solve_plan);
  return pl && pl->slabcoef_h != nullptr;
}

// ---- peer-store exchange: set-up
static size_t comm_gath_bytes(int G, int64_t kmax) { return sizeof(cplx) * 2 * (size_t)G * 6 * (size_t)kmax; }

static SlabCommDev comm_dev_of(pd_handle* h) {
  SolvePlan* pl = plan_of(h);
  SlabCommDev cm;
  memset(&cm, 0, sizeof(cm));
  cm.G = pl->sg.G; cm.rank = pl->sg.rank; cm.kmax = pl->comm_kmax; cm.nflag = pl->comm_nflag;
  cm.epoch = pl->comm_epoch; cm.err = pl->comm_err;
  const size_t goff = comm_gath_bytes(cm.G, cm.kmax);
  for (int r = 0; r < cm.G; ++r) {
    cm.peer_gath[r] = reinterpret_cast<cplx*>(pl->comm_peer[r]);
    cm.peer_flag[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(pl->comm_peer[r]) + goff);
  }
  return cm;
}

// allocates this rank's symmetric buffer; ipc_handle_out (64 bytes, optional) receives its cudaIpcMemHandle_t
int pd_slab_comm_create_impl(pd_handle* h, void* ipc_handle_out, void** base_out) {
  SolvePlan* pl = plan_of(h);
  if (!pl->comm_base) {
    const int G = pl->sg.G;
    pl->comm_kmax = ((int64_t)h->cfg.N_t + 7) & ~7ll;  // covers the half spectrum's Kp as well (N_t >= 128)
    if (pl->comm_kmax < 8) pl->comm_kmax = 8;
    pl->comm_nflag = (int)((pl->comm_kmax + PD_FKB - 1) / PD_FKB);
    pl->comm_bytes = comm_gath_bytes(G, pl->comm_kmax) + sizeof(unsigned long long) * 2 * (size_t)G * pl->comm_nflag;
    PD_CUDA(cudaMalloc(&pl->comm_base, pl->comm_bytes));
    PD_CUDA(cudaMemset(pl->comm_base, 0, pl->comm_bytes));
    PD_CUDA(cudaMalloc(&pl->comm_epoch, sizeof(unsigned long long)));
    PD_CUDA(cudaMemset(pl->comm_epoch, 0, sizeof(unsigned long long)));
    PD_CUDA(cudaMalloc(&pl->comm_err, sizeof(int)));
    PD_CUDA(cudaMemset(pl->comm_err, 0, sizeof(int)));
    PD_CUDA(cudaDeviceSynchronize());
    h->ws_bytes += pl->comm_bytes;
  }
  if (ipc_handle_out) {
    cudaIpcMemHandle_t ih;
    PD_CUDA(cudaIpcGetMemHandle(&ih, pl->comm_base));
    static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &ih, sizeof(ih));
  }
  if (base_out) *base_out = pl->comm_base;
  return PD_OK;
}

// mode 0: `peers` = G cudaIpcMemHandle_t (64 bytes each, rank order; the own entry is ignored)
// mode 1: `peers` = G device pointers (void*[G]) already valid in this process (peer access is enabled here)
int pd_slab_comm_connect_impl(pd_handle* h, const void* peers, int mode, const int* peer_devices) {
  SolvePlan* pl = plan_of(h);
  if (!pl->comm_base) {
    pd_set_error("pd_slab_comm_connect: call pd_slab_comm_create first");
    return PD_ERR_INVALID;
  }
  const int G = pl->sg.G, me = pl->sg.rank;
  for (int r = 0; r < G; ++r) {
    if (r == me) {
      pl->comm_peer[r] = pl->comm_base;
      continue;
    }
    if (mode == 0) {
      cudaIpcMemHandle_t ih;
      memcpy(&ih, reinterpret_cast<const char*>(peers) + 64 * (size_t)r, sizeof(ih));
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        pd_set_error("pd_slab_comm_connect: cudaIpcOpenMemHandle of rank %d failed: %s", r, cudaGetErrorString(e));
        return PD_ERR_CUDA;
      }
      pl->comm_peer[r] = ptr;
      pl->comm_ipc[r] = true;
    } else {
      pl->comm_peer[r] = reinterpret_cast<void* const*>(peers)[r];
      if (peer_devices && peer_devices[r] != h->cfg.device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          pd_set_error("pd_slab_comm_connect: no peer access from device %d to %d: %s", h->cfg.device,
                       peer_devices[r], cudaGetErrorString(e));
          return PD_ERR_CUDA;
        }
        cudaGetLastError();
      }
    }
  }
  pl->comm_connected = true;
  return PD_OK;
}

bool pd_slab_comm_ready(const pd_handle* h) {
  const SolvePlan* pl = reinterpret_cast<const SolvePlan*>(h->solve_plan);
  return pl && pl->comm_connected;
}

// reads (and clears) the "a bounded wait expired" flag and the epoch; synchronises the device
int pd_slab_comm_status_impl(pd_handle* h, int* timed_out, unsigned long long* epoch) {
  SolvePlan* pl = plan_of(h);
  if (!pl->comm_base) {
    pd_set_error("pd_slab_comm_status: no exchange buffer");
    return PD_ERR_INVALID;
  }
  int e = 0;
  unsigned long long ep = 0;
  PD_CUDA(cudaMemcpy(&e, pl->comm_err, sizeof(int), cudaMemcpyDeviceToHost));
  PD_CUDA(cudaMemcpy(&ep, pl->comm_epoch, sizeof(ep), cudaMemcpyDeviceToHost));
  if (e) PD_CUDA(cudaMemset(pl->comm_err, 0, sizeof(int)));
  if (timed_out) *timed_out = e;
  if (epoch) *epoch = ep;
  return PD_OK;
}

// slab mode, first half: pass A, interface levels, slab functionals -> out[6][K]
// (out == nullptr: pushed to every rank's exchange buffer instead, see SlabCommDev)
int pd_slab_reduce_launch(pd_handle* h, cplx* w, cplx* out, cudaStream_t st, int half_spectrum, cudaEvent_t* ev,
                          int passA_done, int koff, int kend) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, half_spectrum);
  if (kend > koff) {  // column range [koff, kend) only (frequency halves of one apply on two streams)
    sp.koff = koff;
    sp.kend = kend < sp.K ? kend : sp.K;
  }
  const int ncol = sp.kend - sp.koff;
  const dim3 grid0 = stream_grid(h, ncol, sp.rows[1] + 1, 4);
  const int kblocks = (ncol + PD_KB - 1) / PD_KB;
  SolvePlan* pl = plan_of(h);
  if (!out && !pl->comm_connected) {
    pd_set_error("slab apply: the peer-store exchange is not connected (pd_slab_comm_create / _connect)");
    return PD_ERR_INVALID;
  }
  bool pushed = false;
  if (sp.nlev >= 1) {
    if (!passA_done) {
      if (sp.al)
        PD_KLAUNCH((pd_solve_passA_kernel<true>), grid0, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, sl.lastl);
      else
        PD_KLAUNCH((pd_solve_passA_kernel<false>), grid0, PD_KB, 0, st, w, lv.F[0], lv.R[1], sp, sl.lastl);
      PD_CHECK_LAUNCH();
      h->launches++;
    }
    if (ev) cudaEventRecord(ev[0], st);
    PushCtx pc;
    if (!out) {
      pc.w = w; pc.sl = sl; pc.cm = comm_dev_of(h);
    }
    int rc = run_interface(h, sp, lv, st, out ? nullptr : &pc, &pushed);
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[1], st);
  } else {
    // a single chunk: its first / last entries come from one forward sweep; reuse pass A with a
    // private F buffer (zout is free at this point: 4K entries >= 2K)
    lv.F[0] = pl->zout;
    if (sp.al)
      PD_KLAUNCH((pd_solve_passA_kernel<true>), grid0, PD_KB, 0, st, w, lv.F[0], nullptr, sp, sl.lastl);
    else
      PD_KLAUNCH((pd_solve_passA_kernel<false>), grid0, PD_KB, 0, st, w, lv.F[0], nullptr, sp, sl.lastl);
    PD_CHECK_LAUNCH();
    h->launches++;
    if (ev) { cudaEventRecord(ev[0], st); cudaEventRecord(ev[1], st); }
  }
  if (pushed) return PD_OK;  // the PCR kernel has already pushed the functionals
  if (out) {
    SlabCommDev none;
    memset(&none, 0, sizeof(none));
    PD_KLAUNCH((pd_slab_functionals_kernel<false>), kblocks, PD_KB, 0, st, w, lv, sp, sl, out, none);
  } else {
    PD_KLAUNCH((pd_slab_functionals_kernel<true>), kblocks, PD_KB, 0, st, w, lv, sp, sl, nullptr, comm_dev_of(h));
  }
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

int pd_slab_epoch_bump_launch(pd_handle* h, cudaStream_t st) {
  SolvePlan* pl = plan_of(h);
  PD_KLAUNCH(pd_slab_epoch_bump_kernel, 1, 1, 0, st, pl->comm_epoch, h->pdl ? h->pdl_early : 0);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// slab mode, second half: global separator solve from the gathered functionals, then pass B
// (gathered == nullptr: waits for the peers' stores into this rank's exchange buffer)
int pd_slab_finish_launch(pd_handle* h, cplx* w, const cplx* gathered, cudaStream_t st, int half_spectrum, cudaEvent_t* ev,
                          int koff, int kend, int bump_epoch) {
  SolveParams sp; Levels lv; SlabPtrs sl;
  fill_params(h, sp, lv, sl, half_spectrum);
  if (bump_epoch && !gathered) sl.epoch_bump = plan_of(h)->comm_epoch;
  if (kend > koff) {
    sp.koff = koff;
    sp.kend = kend < sp.K ? kend : sp.K;
  }
  SolvePlan* pl = plan_of(h);
  const int ncol = sp.kend - sp.koff;
  const dim3 grid0 = stream_grid(h, ncol, sp.rows[1] + 1, 2);
  const int kblocks = (ncol + PD_KB - 1) / PD_KB;
  const cplx* coef = half_spectrum ? pl->slabcoef_h : pl->slabcoef;
  // enough CTAs in y for the correction sweep over the interior separators (rows[1] x 2 x K values)
  // (one batch of 4 rows per thread where the interface is short: all its loads are issued before the flag wait)
  int gy = sp.rows[1] / 4;
  if (gy < 1) gy = 1;
  if (gy > 32) gy = 32;
  const dim3 ggrid(kblocks, gy);
  cplx* zsep = sp.nlev >= 1 ? lv.R[1] : nullptr;
#define PD_GLOBAL_LAUNCH(W, GT_, GPTR, GSTR, CM)                                                                   \
  PD_KLAUNCH((pd_slab_global_kernel<W, GT_>), ggrid, PD_KB, 0, st, GPTR, GSTR, sp, pl->sg, coef, pl->zout, CM, zsep, sl.green)
#define PD_GLOBAL_DISPATCH(W, GPTR, GSTR, CM)                                                                      \
  switch (pl->sg.G) {                                                                                              \
    case 2: PD_GLOBAL_LAUNCH(W, 2, GPTR, GSTR, CM); break;                                                         \
    case 4: PD_GLOBAL_LAUNCH(W, 4, GPTR, GSTR, CM); break;                                                         \
    case 8: PD_GLOBAL_LAUNCH(W, 8, GPTR, GSTR, CM); break;                                                         \
    default: PD_GLOBAL_LAUNCH(W, 0, GPTR, GSTR, CM); break;                                                        \
  }
  if (gathered) {
    SlabCommDev none;
    memset(&none, 0, sizeof(none));
    PD_GLOBAL_DISPATCH(false, gathered, (int64_t)sp.K, none);
  } else {
    if (!pl->comm_connected) {
      pd_set_error("slab apply: the peer-store exchange is not connected (pd_slab_comm_create / _connect)");
      return PD_ERR_INVALID;
    }
    const SlabCommDev cmd = comm_dev_of(h);
    PD_GLOBAL_DISPATCH(true, nullptr, pl->comm_kmax, cmd);
  }
  PD_CHECK_LAUNCH();
  if (ev) cudaEventRecord(ev[0], st);
  if (sp.al)
    PD_KLAUNCH((pd_solve_passB_kernel<true, true>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  else
    PD_KLAUNCH((pd_solve_passB_kernel<true, false>), grid0, PD_KB, 0, st, w, lv.R[1], sp, sl);
  PD_CHECK_LAUNCH();
  h->launches += 2;
  return PD_OK;
}
