// Device-side building blocks of the frequency-domain stage (shared by pd_solve.cu and pd_fused.cu):
// per-frequency coefficients regenerated from k, the S / S^-1 rotations, the cancellation-free level systems
// and the pass-A body.  See pd_solve.cu for the algorithm and the upstream lines it replaces.
#pragma once
#include "pd_common.cuh"

#define PD_L 16            // level-0 chunk length (rows held in registers)
#define PD_LG 8            // chunk length of the generic interface levels
#define PD_KB 128          // frequencies per CTA in the streaming passes
#define PD_AG 4            // rows per software-pipeline group in pass A
#define PD_PCR_MAX 32      // largest interface system handed to the PCR kernel (many frequencies)
#define PD_PCR_MAX_SMALLK 128  // ... when there are few frequencies (launch-bound sizes: fewer kernels win)
#define PD_PCR_THREADS 256
#define PD_PCR_MAXROWS 4   // rows per thread in the PCR kernel
#define PD_MAX_LEVELS 8
#define PD_MAX_SLABS 16

struct SolveParams {
  int n, m, K, kbegin, N_t;
  double h, dt2, c;
  int64_t plane;            // elements per field plane = n * K
  int nlev;                 // top interface level (0: single chunk, no interface)
  int rows[PD_MAX_LEVELS];  // rows[l] = size of the level-l system (rows[0] = m)
  // geometry of the local node rows: row 0 precedes the body, rows 1..m are the body.
  // Single-GPU: row 0 and row m+1 are the Dirichlet nodes.  Slab mode (x-slab r of G): row 0 is the
  // inter-slab separator (r > 0) and the last body row is followed by the next slab's separator (r < G-1).
  int first_dirichlet, last_dirichlet;
  int freq_perm;  // 1: columns hold [k mod 4 = 0 | 1 | 2 | 3] (the N_t = 16384 FFT kernel's frequency order)
  int koff, kend; // column range [koff, kend) this launch works on (row stride stays K)
  int c0, c1;     // level-0 chunk range [c0, c1) this launch of pass A / pass B works on (all: 0, rows[1] + 1)
  int al;         // 1: alpha != 1 (extension, see make_coef<true>); 0: the upstream operator
  double lna;     // ln(alpha) / N_t
  int pdl_early;  // bit mask: which kernels release their programmatic-launch dependents at once (pd_handle::pdl_early)
};

// Slab-mode extras (device pointers; all null in single-GPU mode)
struct SlabPtrs {
  cplx* lastl;        // [2][K]     last entry of the last chunk's local solve (pass A)
  const cplx* green;  // [P][2][K]  interface Green's vectors T_1^-1 e_0 (slot 0), T_1^-1 e_{P-1} (slot 1)
  const cplx* zout;   // [4][K]     outer separator values: left (+, -), right (+, -)
  unsigned long long* epoch_bump;  // peer-store exchange: pass B marks the apply complete (else null)
};

struct KCoef {
  cplx a;        // off-diagonal of Tt (the diagonal is b = sh - 2a)
  cplx sh;       // s h = b + 2a: the detuning from the discrete resonance, cancellation-free
  // rotation, alpha = 1 (the upstream operator)
  cplx zc;       // conj(z) = e^{-i theta}
  double sigma;  // sign(cos theta)
  // rotation, alpha != 1 (general 2x2 eigen-decomposition, see make_coef<true>)
  double gp, gm;    // g_+- = (d +- beta) / (2 d)
  cplx e;           // i c e^{i phi} / (2 d)
  cplx eic;         // e^{-i phi}
  double bmd, bpd;  // (beta -+ d) / c
};

// frequency index of column `kk` of this handle's frequency block
__device__ __forceinline__ int freq_of(const SolveParams& sp, int kk) {
  const int col = sp.kbegin + kk;
  if (!sp.freq_perm) return col;
  const int quarter = sp.N_t >> 2;
  return 4 * (col & (quarter - 1)) + col / quarter;
}

// AL = false: the upstream (alpha = 1) operator, division-free closed forms.
// AL = true : the alpha extension (oracle/pc_alpha.py; no upstream counterpart).  With a = alpha^(1/N_t),
//   l1 = (1 - a z)^2,  l2 = 1 + a^2 z^2 = |l2| e^{i phi},  mu = l1 e^{-i phi},  beta = Im mu,  d = sqrt(beta^2 + c^2):
//   Tt = s M + kap K,  s = Re mu + i d,  kap = dt^2/2 |l2|   (conj(Tt) serves the second eigenvalue)
//   rho_+ = g_+ uh + e ph,  rho_- = g_- uh - e ph;  wh_u = e^{-i phi}(zeta_+ + zeta_-),
//   wh_p = i [(beta - d) zeta_+ + (beta + d) zeta_-] / c.     |l2| >= 1 - a^2 > 0 for alpha < 1.
// Unused members are dead code in each instantiation (everything is inlined).
template <bool AL>
__device__ __forceinline__ KCoef make_coef(int kglob, const SolveParams& sp) {
  KCoef kc;
  double st, ct, sh, chh;
  sincospi(2.0 * (double)kglob / (double)sp.N_t, &st, &ct);
  sincospi((double)kglob / (double)sp.N_t, &sh, &chh);
  if (!AL) {
    kc.sigma = ct >= 0.0 ? 1.0 : -1.0;
    const double sre = -4.0 * sh * sh;
    const double sim = sp.c * kc.sigma;
    const double kap = sp.dt2 * ct;
    kc.a = cmake(sre * (sp.h / 6.0) - kap / sp.h, sim * (sp.h / 6.0));
    kc.sh = cmake(sre * sp.h, sim * sp.h);  // b + 2a = s (2h/3 + 2 h/6): the stiffness parts cancel exactly
    kc.zc = cmake(ct, -st);
  } else {
    const double a = exp(sp.lna), oma = -expm1(sp.lna), oma2 = -expm1(2.0 * sp.lna);  // a, 1 - a, 1 - a^2
    const cplx q = cmake(oma + 2.0 * a * sh * sh, -a * st);                              // 1 - a z
    const cplx l1 = cmul(q, q);
    const cplx l2 = cmake(oma2 + 2.0 * a * a * ct * ct, 2.0 * a * a * st * ct);
    const double al2 = sqrt(l2.x * l2.x + l2.y * l2.y);
    const cplx eiphi = cmake(l2.x / al2, l2.y / al2);
    const cplx mu = cmulc(l1, eiphi);  // l1 e^{-i phi}
    const double beta = mu.y, d = sqrt(beta * beta + sp.c * sp.c);
    const double kap = 0.5 * sp.dt2 * al2;
    kc.a = cmake(mu.x * (sp.h / 6.0) - kap / sp.h, d * (sp.h / 6.0));
    kc.sh = cmake(mu.x * sp.h, d * sp.h);
    const double i2d = 0.5 / d;
    kc.gp = (d + beta) * i2d;
    kc.gm = (d - beta) * i2d;
    kc.e = cmake(-sp.c * eiphi.y * i2d, sp.c * eiphi.x * i2d);  // i c e^{i phi} / (2 d)
    kc.eic = cconj(eiphi);
    kc.bmd = (beta - d) / sp.c;
    kc.bpd = (beta + d) / sp.c;
  }
  return kc;
}
// kernels that only need (a, sh): one run-time switch
__device__ __forceinline__ KCoef make_coef(int kglob, const SolveParams& sp) {
  return sp.al ? make_coef<true>(kglob, sp) : make_coef<false>(kglob, sp);
}

// rho_+ and conj(rho_-) from (u-hat, p-hat)
template <bool AL>
__device__ __forceinline__ void rotate_in(const KCoef& kc, cplx u, cplx p, cplx& rp, cplx& rm) {
  if (!AL) {
    cplx uz = cmul(u, kc.zc);
    cplx ip = cmake(-p.y * kc.sigma, p.x * kc.sigma);  // i sigma p
    rp = cmake(0.5 * (uz.x + ip.x), 0.5 * (uz.y + ip.y));
    rm = cmake(0.5 * (uz.x - ip.x), -0.5 * (uz.y - ip.y));  // conjugated
  } else {
    const cplx ep = cmul(kc.e, p);
    rp = cmake(kc.gp * u.x + ep.x, kc.gp * u.y + ep.y);
    rm = cmake(kc.gm * u.x - ep.x, -(kc.gm * u.y - ep.y));  // conjugated
  }
}
// (w_u, w_p) from zeta_+ and conj(zeta_-)
template <bool AL>
__device__ __forceinline__ void rotate_out(const KCoef& kc, cplx zp, cplx zmc, cplx& wu, cplx& wp) {
  cplx zm = cconj(zmc);
  if (!AL) {
    wu = cadd(zp, zm);
    cplx d = csub(zp, zm);
    cplx t = cmulc(d, kc.zc);                         // d * z
    wp = cmake(t.y * kc.sigma, -t.x * kc.sigma);      // -i sigma (d z)
  } else {
    wu = cmul(kc.eic, cadd(zp, zm));
    const cplx t = cmake(kc.bmd * zp.x + kc.bpd * zm.x, kc.bmd * zp.y + kc.bpd * zm.y);
    wp = cmake(-t.y, t.x);                            // i t
  }
}

// A level system: tridiag(off, d, off) with n rows, d = dmain except the last row (dlast).
//
// Near a discrete wave resonance dmain ~ -2 off, and everything that matters sits in the small
// "detuning" det = dmain + 2 off (at level 0: det = b + 2a = s h EXACTLY, the lumped mass term).
// Forming pivots 1/(dmain - off^2 m) or Schur complements dmain - 2 off^2 alpha from dmain itself
// loses det to rounding (relative error eps |off/det|, ~1e-8 * eps^-1... i.e. 1e-9 at N_x = 4096).
// The system is therefore carried as (off, det, glast = dlast - dmain) and all chunk quantities come
// from the cancellation-free recurrence for V_i = (-1)^i U_i(dmain / (2 off)) (Chebyshev U):
//     eta = det/off,  V_0 = 1,  E_1 = -eta,  V_i = V_{i-1} + 1 + E_i,  E_{i+1} = E_i - eta V_i
// with  pivot m_i = -V_{i-1}/(off V_i),  prod_{t<i}(-off m_t) = 1/V_{i-1},  (T_L^-1)_{1L} = -1/(off V_L).
// Measured against an 80-bit solve this is ~100x more accurate than plain fp64 LU (Thomas) of the
// same systems.
struct Sys {
  cplx off, det, glast;
  int n;
};
__device__ __forceinline__ cplx sys_dmain(const Sys& s) { return cmake(s.det.x - 2.0 * s.off.x, s.det.y - 2.0 * s.off.y); }

__device__ __host__ __forceinline__ int chunk_len(int level) { return level == 0 ? PD_L : PD_LG; }

// The (V, E) recurrence in homogeneous form: (V, e, one) may be rescaled together at any time, only
// ratios are ever used.  Far from resonance |eta| is large, V grows geometrically (the interface
// couplings decay accordingly) and would overflow after two levels without the rescaling; when the
// coupling is below 1e-60 of the diagonal the system is treated as decoupled (`diag`).
struct VRec {
  cplx eta, V, e, E;  // E: the increment used in the last step (E_i = V_i - V_{i-1} - one)
  double one;
  bool diag;
  __device__ __forceinline__ void init(cplx off, cplx det, cplx extra /* added to eta in the first step */) {
    const double mo = fabs(off.x) + fabs(off.y), md = fabs(det.x) + fabs(det.y);
    // |eta| <= 1e60 keeps eta * V (|V| <= 1e100 after rescaling, one step of growth |eta|) finite
    diag = !(mo > md * 1e-60);
    eta = diag ? cmake(0, 0) : cmul(det, crcp(off));
    V = cmake(1, 0);
    one = 1.0;
    e = cneg(cadd(eta, extra));
    E = e;
  }
  __device__ __forceinline__ void step() {
    E = e;
    V = cmake(V.x + one + E.x, V.y + E.y);
    e = cfms(eta, V, E);
    const double mag = fmax(fabs(V.x), fabs(V.y));
    if (mag > 1e100) {
      const double sc = 1.0 / mag;
      V = cscale(V, sc); e = cscale(e, sc); E = cscale(E, sc); one *= sc;
    }
  }
};

// Interface system obtained by cutting `s` into chunks of L rows + one separator each.
__device__ __forceinline__ Sys reduce_sys(const Sys& s, int L) {
  const int P = s.n / (L + 1), Llast = s.n - P * (L + 1);
  Sys r;
  r.n = P;
  VRec v;
  v.init(s.off, s.det, cmake(0, 0));
  if (v.diag) {  // no coupling left: the interface rows are plain diagonal equations
    r.off = cmake(0, 0);
    r.det = sys_dmain(s);
    r.glast = Llast == 0 ? s.glast : cmake(0, 0);
    return r;
  }
  for (int i = 1; i <= L; ++i) v.step();              // V_L, E_L
  const cplx rV = crcp(v.V);
  const cplx DLVL = cmul(cmake(v.one + v.E.x, v.E.y), rV);  // (V_L - V_{L-1}) / V_L
  r.off = cmul(s.off, cscale(rV, v.one));                                // -off^2 (T_L^-1)_{1L} = off / V_L
  r.det = cmul(s.off, cfms(cmake(2.0 * v.E.x, 2.0 * v.E.y), rV, v.eta));  // off (eta - 2 E_L / V_L)
  if (Llast > 0) {
    // last chunk (Llast rows, bottom diagonal dmain + glast), counted from its bottom row: W_i
    VRec w;
    w.init(s.off, s.det, cmul(s.glast, crcp(s.off)));
    for (int i = 1; i <= Llast; ++i) w.step();
    const cplx DW = cmul(cmake(w.one + w.E.x, w.E.y), crcp(w.V));
    r.glast = cmul(s.off, csub(DLVL, DW));                               // off^2 (alpha - alpha_first)
  } else {
    // the last separator is the last row of s itself: glast' = glast - off V_{L-1}/V_L
    r.glast = csub(s.glast, cmul(s.off, cmake(1.0 - DLVL.x, -DLVL.y)));
  }
  return r;
}

__device__ __forceinline__ Sys level_sys(const KCoef& kc, const SolveParams& sp, int level) {
  Sys s;
  s.off = kc.a; s.det = kc.sh; s.glast = cmake(0, 0); s.n = sp.m;
  for (int l = 0; l < level; ++l) s = reduce_sys(s, chunk_len(l));
  return s;
}

// Running generator of the chunk-local pivots m_i = -V_{i-1}/(off V_i), i = 1, 2, ...
struct PivotGen {
  VRec v;
  cplx roff, mdiag;
  __device__ __forceinline__ void init(const Sys& s) {
    v.init(s.off, s.det, cmake(0, 0));
    roff = v.diag ? cmake(0, 0) : crcp(s.off);
    mdiag = v.diag ? crcp(sys_dmain(s)) : cmake(0, 0);
  }
  __device__ __forceinline__ cplx next() {
    if (v.diag) return mdiag;
    const cplx Vp = v.V;
    // ratio V_{i-1} / V_i taken before any rescaling of V_i
    const cplx Vn = cmake(Vp.x + v.one + v.e.x, Vp.y + v.e.y);
    const cplx m = cneg(cmul(cmul(Vp, roff), crcp(Vn)));
    v.step();
    return m;
  }
};
// pivot of the very last row of a level system (diagonal dmain + glast) from its regular value
__device__ __forceinline__ cplx last_row_pivot(cplx m_reg, cplx glast) {
  return cmul(m_reg, crcp(cfma(glast, m_reg, cmake(1, 0))));
}

// pivots of the level-0 chunk-local LU, one shared-memory column per thread
template <int KBT>
__device__ __forceinline__ void fill_pivots(const KCoef& kc, cplx (*mtab)[KBT], int tid) {
  Sys s;
  s.off = kc.a; s.det = kc.sh; s.glast = cmake(0, 0); s.n = 0;
  PivotGen pg;
  pg.init(s);
#pragma unroll
  for (int i = 0; i < PD_L; ++i) mtab[i][tid] = pg.next();
}

// ------------------------------------------------------------------- pass A, one (k, chunk)
// Level-0 reduce of chunk c for the frequency this thread owns: forward recurrences in registers -> first entry
// f_c of the local solve (F0), last entry folded into the following separator's right-hand side (R1); S^-1
// rotation fused on load.  KBT = threads per CTA (row length of the pivot table); LDCG: read w through L2 only
// (ld.global.cg) -- needed when w was written by OTHER CTAs of the same kernel (pd_fused.cu), harmless otherwise.
template <bool LDCG>
__device__ __forceinline__ cplx ldw(const cplx* p) {
  return LDCG ? __ldcg(p) : *p;
}
template <bool AL, int KBT, bool LDCG>
__device__ __forceinline__ void passA_chunk(const cplx* __restrict__ wu, const cplx* __restrict__ wp,
                                            cplx* __restrict__ F0, cplx* __restrict__ R1, const SolveParams& sp,
                                            cplx* __restrict__ lastl, const KCoef& kc, cplx (*mtab)[KBT], int tid,
                                            int kk, bool valid, int c) {
  const int64_t K = sp.K;
  const int P = sp.rows[1], Llast = sp.m - P * (PD_L + 1);
  const int Lc = c < P ? PD_L : Llast;
  const int j0 = c * (PD_L + 1) + 1;
  const int nrows = Lc + (c < P ? 1 : 0);  // chunk rows + the separator row that follows
  // Software pipeline over groups of PD_AG rows: the next group's loads are in flight while the current
  // group's recurrences run.  (Loading all 17 rows up front costs 254 registers = 2 CTAs per SM.)
  // (loads are unconditional on a clamped row index -- a row past the chunk is re-read, never used: predicated or
  // branched loads are not issued back to back and cost a third of the kernel's bandwidth)
  // (an empty last chunk of an x-slab that does not end with the Dirichlet row has NO row j0: stay inside the array)
  const int rlast = min(nrows > 0 ? nrows - 1 : 0, sp.n - 1 - j0);
  cplx bu[2][PD_AG], bp[2][PD_AG];
#pragma unroll
  for (int r = 0; r < PD_AG; ++r) {
    const int64_t ro = (int64_t)(j0 + min(r, rlast)) * K;
    bu[0][r] = ldw<LDCG>(wu + ro);
    bp[0][r] = ldw<LDCG>(wp + ro);
  }
  cplx dP = cmake(0, 0), dM = cmake(0, 0), fP = cmake(0, 0), fM = cmake(0, 0);
  cplx pi = cmake(1, 0), sP = cmake(0, 0), sM = cmake(0, 0);
#pragma unroll 1
  for (int g = 0; g < (PD_L + 1 + PD_AG - 1) / PD_AG; g += 2) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int base = (g + half) * PD_AG;
      // prefetch the following group into the other buffer
#pragma unroll
      for (int r = 0; r < PD_AG; ++r) {
        const int64_t ro = (int64_t)(j0 + min(base + PD_AG + r, rlast)) * K;
        bu[half ^ 1][r] = ldw<LDCG>(wu + ro);
        bp[half ^ 1][r] = ldw<LDCG>(wp + ro);
      }
#pragma unroll
      for (int r = 0; r < PD_AG; ++r) {
        const int i = base + r;
        if (i < Lc) {
          cplx rP, rM;
          rotate_in<AL>(kc, bu[half][r], bp[half][r], rP, rM);
          const cplx mi = mtab[i][tid];
          if (i > 0) {
            const cplx cp = cmul(kc.a, mtab[i - 1][tid]);  // c'_{i-1}
            pi = cneg(cmul(pi, cp));
          }
          dP = cmul(cfms(kc.a, dP, rP), mi);
          dM = cmul(cfms(kc.a, dM, rM), mi);
          fP = cfma(pi, dP, fP);
          fM = cfma(pi, dM, fM);
        } else if (i == PD_L && c < P) {
          rotate_in<AL>(kc, bu[half][r], bp[half][r], sP, sM);
        }
      }
    }
  }
  if (valid) {
    F0[((int64_t)c * 2) * K + kk] = fP;
    F0[((int64_t)c * 2 + 1) * K + kk] = fM;
    if (c < P) {
      R1[((int64_t)c * 2) * K + kk] = cfms(kc.a, dP, sP);      // rho_sep - a l_c
      R1[((int64_t)c * 2 + 1) * K + kk] = cfms(kc.a, dM, sM);
    } else if (lastl) {
      lastl[kk] = dP;
      lastl[K + kk] = dM;
    }
  }
}

// Workspace of the interface levels (device pointers, by value in kernel params).
//   R[l], l >= 1 : [rows[l]][2][K]       written by level l-1 as  rhs_sep - off_{l-1} l_c ; the
//                                        final right-hand side of row q is R[l][q] - off_{l-1} F[l-1][q+1];
//                                        overwritten by the solution on the way back
//   F[l], l >= 0 : [rows[l+1] + 1][2][K] f_c = first entry of chunk c's local solve
struct Levels {
  cplx* R[PD_MAX_LEVELS];
  cplx* F[PD_MAX_LEVELS];
};

