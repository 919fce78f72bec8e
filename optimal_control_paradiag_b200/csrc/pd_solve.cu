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
// Algorithm (layout [field][node][k], k fastest, so a warp reads 512 contiguous
// bytes per node row): the interior nodes are cut into chunks of L rows with
// one separator row between chunks.
//   pass A  (streaming, 1 read sweep): per (k, chunk) a register-resident
//           forward recurrence yields the first and last entry of the local
//           solve Tt_L^-1 rho_chunk.
//   PCR     the separators of one k form a tridiagonal interface system
//           (P = m/(L+1) unknowns); it is solved by parallel cyclic reduction
//           held in shared memory, one CTA per 1..4 frequencies.
//   pass B  (streaming, 1 read + 1 write sweep): per (k, chunk) Thomas with the
//           now-known separator values, rotation back, store in place.
// The pivots m_i of the chunk-local factorisation depend on (k, i) only; each
// CTA regenerates them once into a per-thread shared-memory column and reuses
// them for all the chunks it visits.
#include "pd_common.cuh"

#define PD_L 16          // chunk length (rows held in registers)
#define PD_KB 128        // frequencies per CTA in the streaming passes
#define PD_PCR_THREADS 256
#define PD_PCR_MAXROWS 8  // rows per thread in the PCR kernel

struct SolveParams {
  int n, m, K, kbegin, N_t;
  int P, Llast;
  double h, dt2, c;
  int64_t plane;  // elements per field plane = n * K
};

struct KCoef {
  cplx a, b;     // off-diagonal / diagonal of Tt
  cplx zc;       // conj(z) = e^{-i theta}
  double sigma;  // sign(cos theta)
};

__device__ __forceinline__ KCoef make_coef(int kglob, const SolveParams& sp) {
  KCoef kc;
  double st, ct, sh, chh;
  sincospi(2.0 * (double)kglob / (double)sp.N_t, &st, &ct);
  sincospi((double)kglob / (double)sp.N_t, &sh, &chh);
  kc.sigma = ct >= 0.0 ? 1.0 : -1.0;
  const double sre = -4.0 * sh * sh;
  const double sim = sp.c * kc.sigma;
  const double kap = sp.dt2 * ct;
  kc.a = cmake(sre * (sp.h / 6.0) - kap / sp.h, sim * (sp.h / 6.0));
  kc.b = cmake(sre * (2.0 * sp.h / 3.0) + 2.0 * kap / sp.h, sim * (2.0 * sp.h / 3.0));
  kc.zc = cmake(ct, -st);
  return kc;
}

// rho_+ and conj(rho_-) from (u-hat, p-hat)
__device__ __forceinline__ void rotate_in(const KCoef& kc, cplx u, cplx p, cplx& rp, cplx& rm) {
  cplx uz = cmul(u, kc.zc);
  cplx ip = cmake(-p.y * kc.sigma, p.x * kc.sigma);  // i sigma p
  rp = cmake(0.5 * (uz.x + ip.x), 0.5 * (uz.y + ip.y));
  rm = cmake(0.5 * (uz.x - ip.x), -0.5 * (uz.y - ip.y));  // conjugated
}
// (w_u, w_p) from zeta_+ and conj(zeta_-)
__device__ __forceinline__ void rotate_out(const KCoef& kc, cplx zp, cplx zmc, cplx& wu, cplx& wp) {
  cplx zm = cconj(zmc);
  wu = cadd(zp, zm);
  cplx d = csub(zp, zm);
  cplx t = cmulc(d, kc.zc);                         // d * z
  wp = cmake(t.y * kc.sigma, -t.x * kc.sigma);      // -i sigma (d z)
}

// pivots of the chunk-local LU: m_1 = 1/b, m_i = 1/(b - a^2 m_{i-1})
__device__ __forceinline__ void fill_pivots(const KCoef& kc, cplx (*mtab)[PD_KB], int tid) {
  const cplx a2 = cmul(kc.a, kc.a);
  cplx m = crcp(kc.b);
  mtab[0][tid] = m;
#pragma unroll
  for (int i = 1; i < PD_L; ++i) {
    m = crcp(cfms(a2, m, kc.b));
    mtab[i][tid] = m;
  }
}

// ------------------------------------------------------------------- pass A
__global__ void __launch_bounds__(PD_KB)
pd_solve_passA_kernel(const cplx* __restrict__ w, cplx* __restrict__ red, SolveParams sp) {
  __shared__ cplx mtab[PD_L][PD_KB];
  const int tid = threadIdx.x;
  const int kk = blockIdx.x * PD_KB + tid;
  const bool valid = kk < sp.K;
  const int kc_idx = valid ? kk : sp.K - 1;
  const KCoef kc = make_coef(sp.kbegin + kc_idx, sp);
  fill_pivots(kc, mtab, tid);
  const cplx* wu = w + kc_idx;
  const cplx* wp = w + sp.plane + kc_idx;
  for (int c = blockIdx.y; c <= sp.P; c += gridDim.y) {
    const int Lc = c < sp.P ? PD_L : sp.Llast;
    const int j0 = c * (PD_L + 1) + 1;
    cplx ru[PD_L], rp_[PD_L];
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        ru[i] = wu[(int64_t)(j0 + i) * sp.K];
        rp_[i] = wp[(int64_t)(j0 + i) * sp.K];
      }
    }
    cplx dP = cmake(0, 0), dM = cmake(0, 0), fP = cmake(0, 0), fM = cmake(0, 0);
    cplx pi = cmake(1, 0);
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        cplx rP, rM;
        rotate_in(kc, ru[i], rp_[i], rP, rM);
        const cplx mi = mtab[i][tid];
        if (i > 0) {
          const cplx cp = cmul(kc.a, mtab[i - 1][tid]);  // c'_{i-1}
          pi = cneg(cmul(pi, cp));
        }
        dP = cmul(cfms(kc.a, dP, rP), mi);
        dM = cmul(cfms(kc.a, dM, rM), mi);
        fP = cfma(pi, dP, fP);
        fM = cfma(pi, dM, fM);
      }
    }
    if (valid) {
      cplx* r = red + ((int64_t)c * 4) * sp.K + kk;
      r[0] = fP;
      r[sp.K] = dP;
      r[2 * (int64_t)sp.K] = fM;
      r[3 * (int64_t)sp.K] = dM;
    }
  }
}

// ------------------------------------------------------ interface system (PCR)
// Rows are kept normalised (unit diagonal): (lo, 1, up | rP, rM).
template <int KPB>
__global__ void __launch_bounds__(PD_PCR_THREADS)
pd_solve_pcr_kernel(cplx* __restrict__ w, const cplx* __restrict__ red, cplx* __restrict__ zsep,
                    SolveParams sp) {
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  const int P = sp.P;
  const int rows = P * KPB;
  cplx* s_lo = reinterpret_cast<cplx*>(pd_smem_raw);
  cplx* s_up = s_lo + rows;
  cplx* s_rp = s_up + rows;
  cplx* s_rm = s_rp + rows;
  const int tid = threadIdx.x;
  const int kk0 = blockIdx.x * KPB;

  // build: row index idx -> (q = idx / KPB, ks = idx % KPB); smem slot = ks * P + q
  for (int idx = tid; idx < rows; idx += PD_PCR_THREADS) {
    const int q = idx / KPB, ks = idx - q * KPB;
    int kk = kk0 + ks;
    if (kk >= sp.K) kk = sp.K - 1;
    const KCoef kc = make_coef(sp.kbegin + kk, sp);
    const cplx a2 = cmul(kc.a, kc.a);
    // alpha_L = m_L, beta_L = pi_L m_L; alpha of the (possibly shorter) last chunk
    cplx m = crcp(kc.b), pi = cmake(1, 0), alast = cmake(0, 0);
    if (sp.Llast == 1) alast = m;
#pragma unroll
    for (int i = 1; i < PD_L; ++i) {
      pi = cneg(cmul(pi, cmul(kc.a, m)));
      m = crcp(cfms(a2, m, kc.b));
      if (i + 1 == sp.Llast) alast = m;
    }
    const cplx alpha = m, beta = cmul(pi, m);
    const cplx aright = (q + 1 < P) ? alpha : alast;
    cplx di = cfms(a2, cadd(alpha, aright), kc.b);
    const cplx dinv = crcp(di);
    const cplx off = cneg(cmul(a2, beta));
    const int j = q * (PD_L + 1) + PD_L + 1;  // node of separator q
    cplx rP, rM;
    rotate_in(kc, w[(int64_t)j * sp.K + kk], w[sp.plane + (int64_t)j * sp.K + kk], rP, rM);
    const cplx* r0 = red + ((int64_t)q * 4) * sp.K + kk;        // chunk q   : (f+, l+, f-, l-)
    const cplx* r1 = red + ((int64_t)(q + 1) * 4) * sp.K + kk;  // chunk q+1
    rP = cfms(kc.a, cadd(r0[sp.K], r1[0]), rP);
    rM = cfms(kc.a, cadd(r0[3 * (int64_t)sp.K], r1[2 * (int64_t)sp.K]), rM);
    const int slot = ks * P + q;
    s_lo[slot] = q > 0 ? cmul(off, dinv) : cmake(0, 0);
    s_up[slot] = q + 1 < P ? cmul(off, dinv) : cmake(0, 0);
    s_rp[slot] = cmul(rP, dinv);
    s_rm[slot] = cmul(rM, dinv);
  }
  __syncthreads();

  for (int delta = 1; delta < P; delta <<= 1) {
    cplx nlo[PD_PCR_MAXROWS], nup[PD_PCR_MAXROWS], nrp[PD_PCR_MAXROWS], nrm[PD_PCR_MAXROWS];
#pragma unroll
    for (int it = 0; it < PD_PCR_MAXROWS; ++it) {
      const int idx = tid + it * PD_PCR_THREADS;
      if (idx < rows) {
        const int q = idx / KPB, ks = idx - q * KPB;
        const int slot = ks * P + q;
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
        if (q + delta < P) {
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
        const int q = idx / KPB, ks = idx - q * KPB;
        const int slot = ks * P + q;
        s_lo[slot] = nlo[it];
        s_up[slot] = nup[it];
        s_rp[slot] = nrp[it];
        s_rm[slot] = nrm[it];
      }
    }
    __syncthreads();
  }

  // write the interface values (for pass B) and the finished separator rows
  for (int idx = tid; idx < rows; idx += PD_PCR_THREADS) {
    const int q = idx / KPB, ks = idx - q * KPB;
    const int kk = kk0 + ks;
    if (kk >= sp.K) continue;
    const int slot = ks * P + q;
    const cplx zp = s_rp[slot], zmc = s_rm[slot];
    zsep[((int64_t)q * 2) * sp.K + kk] = zp;
    zsep[((int64_t)q * 2 + 1) * sp.K + kk] = zmc;
    const KCoef kc = make_coef(sp.kbegin + kk, sp);
    cplx wu, wp;
    rotate_out(kc, zp, zmc, wu, wp);
    const int j = q * (PD_L + 1) + PD_L + 1;
    w[(int64_t)j * sp.K + kk] = wu;
    w[sp.plane + (int64_t)j * sp.K + kk] = wp;
  }
}

// ------------------------------------------------------------------- pass B
__global__ void __launch_bounds__(PD_KB)
pd_solve_passB_kernel(cplx* __restrict__ w, const cplx* __restrict__ zsep, SolveParams sp) {
  __shared__ cplx mtab[PD_L][PD_KB];
  const int tid = threadIdx.x;
  const int kk = blockIdx.x * PD_KB + tid;
  const bool valid = kk < sp.K;
  const int kc_idx = valid ? kk : sp.K - 1;
  const KCoef kc = make_coef(sp.kbegin + kc_idx, sp);
  fill_pivots(kc, mtab, tid);
  cplx* wu = w + kc_idx;
  cplx* wp = w + sp.plane + kc_idx;
  const cplx zero = cmake(0, 0);
  for (int c = blockIdx.y; c <= sp.P; c += gridDim.y) {
    const int Lc = c < sp.P ? PD_L : sp.Llast;
    const int j0 = c * (PD_L + 1) + 1;
    cplx dP[PD_L], dM[PD_L];
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        dP[i] = wu[(int64_t)(j0 + i) * sp.K];
        dM[i] = wp[(int64_t)(j0 + i) * sp.K];
      }
    }
    cplx zlP = zero, zlM = zero, zrP = zero, zrM = zero;
    if (c > 0) {
      zlP = zsep[((int64_t)(c - 1) * 2) * sp.K + kc_idx];
      zlM = zsep[((int64_t)(c - 1) * 2 + 1) * sp.K + kc_idx];
    }
    if (c < sp.P) {
      zrP = zsep[((int64_t)c * 2) * sp.K + kc_idx];
      zrM = zsep[((int64_t)c * 2 + 1) * sp.K + kc_idx];
    }
    // forward elimination (in place: d_i overwrites rho_i)
    cplx pP = zlP, pM = zlM;  // "d_{-1}" = known left neighbour value
#pragma unroll
    for (int i = 0; i < PD_L; ++i) {
      if (i < Lc) {
        cplx rP, rM;
        rotate_in(kc, dP[i], dM[i], rP, rM);
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
        rotate_out(kc, nP, nM, ou, op);
        if (valid) {
          wu[(int64_t)(j0 + i) * sp.K] = ou;
          wp[(int64_t)(j0 + i) * sp.K] = op;
        }
      }
    }
    // Dirichlet rows: output exactly 0 (:482, bcs :44-45)
    if (valid && c == 0) {
      wu[0] = zero;
      wp[0] = zero;
    }
    if (valid && c == sp.P) {
      wu[(int64_t)(sp.n - 1) * sp.K] = zero;
      wp[(int64_t)(sp.n - 1) * sp.K] = zero;
    }
  }
}

// --------------------------------------------------------------- host side
static int pcr_kpb(int P) {
  // largest KPB in {4,2,1} with rows <= threads*maxrows and <= ~96 KB of shared memory
  for (int kpb = 4; kpb >= 1; kpb >>= 1) {
    size_t rows = (size_t)P * kpb;
    if (rows <= (size_t)PD_PCR_THREADS * PD_PCR_MAXROWS && rows * 64 <= 96 * 1024) return kpb;
  }
  return 1;
}

int pd_solve_plan(pd_handle* h) {
  h->L = PD_L;
  h->P = h->m / (PD_L + 1);
  h->Llast = h->m % (PD_L + 1);
  const size_t K = (size_t)h->kcount;
  if ((size_t)h->P > (size_t)PD_PCR_THREADS * PD_PCR_MAXROWS || (size_t)h->P * 64 > 227 * 1024) {
    pd_set_error("N_x = %d gives an interface system of %d rows per frequency; the shared-memory PCR "
                 "kernel supports at most %d", h->cfg.N_x, h->P, PD_PCR_THREADS * PD_PCR_MAXROWS);
    return PD_ERR_INVALID;
  }
  size_t red_bytes = sizeof(cplx) * (size_t)(h->P + 1) * 4 * K;
  size_t zs_bytes = sizeof(cplx) * (size_t)(h->P > 0 ? h->P : 1) * 2 * K;
  PD_CUDA(cudaMalloc(&h->red, red_bytes));
  PD_CUDA(cudaMalloc(&h->zsep, zs_bytes));
  h->ws_bytes += red_bytes + zs_bytes;
  return PD_OK;
}

int pd_solve_launch(pd_handle* h, cplx* w, cudaStream_t st) {
  SolveParams sp;
  sp.n = h->n; sp.m = h->m; sp.K = h->kcount; sp.kbegin = h->kbegin; sp.N_t = h->cfg.N_t;
  sp.P = h->P; sp.Llast = h->Llast;
  sp.h = h->h; sp.dt2 = h->dt * h->dt; sp.c = h->c;
  sp.plane = (int64_t)h->n * h->kcount;
  const int kblocks = (sp.K + PD_KB - 1) / PD_KB;
  int nchunks = sp.P + 1;
  // enough CTAs for ~6 resident per SM; more chunks than that are looped over
  int ny = (h->num_sms * 6 + kblocks - 1) / kblocks;
  if (ny > nchunks) ny = nchunks;
  if (ny < 1) ny = 1;
  if (ny > 65535) ny = 65535;
  dim3 grid(kblocks, ny);
  if (sp.P > 0) {
    pd_solve_passA_kernel<<<grid, PD_KB, 0, st>>>(w, h->red, sp);
    PD_CHECK_LAUNCH();
    h->launches++;
    const int kpb = pcr_kpb(sp.P);
    const size_t smem = (size_t)sp.P * kpb * 64;
    const int nblk = (sp.K + kpb - 1) / kpb;
    switch (kpb) {
      case 4:
        PD_CUDA(cudaFuncSetAttribute(pd_solve_pcr_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        pd_solve_pcr_kernel<4><<<nblk, PD_PCR_THREADS, smem, st>>>(w, h->red, h->zsep, sp);
        break;
      case 2:
        PD_CUDA(cudaFuncSetAttribute(pd_solve_pcr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        pd_solve_pcr_kernel<2><<<nblk, PD_PCR_THREADS, smem, st>>>(w, h->red, h->zsep, sp);
        break;
      default:
        PD_CUDA(cudaFuncSetAttribute(pd_solve_pcr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        pd_solve_pcr_kernel<1><<<nblk, PD_PCR_THREADS, smem, st>>>(w, h->red, h->zsep, sp);
        break;
    }
    PD_CHECK_LAUNCH();
    h->launches++;
  }
  pd_solve_passB_kernel<<<grid, PD_KB, 0, st>>>(w, h->zsep, sp);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}
