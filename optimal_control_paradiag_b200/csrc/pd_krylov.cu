// All-at-once matvec, manufactured right-hand side and the device-resident
// GMRES loop (complex128, sm_100a).
//
//   pd_matvec_launch : Jacobian action of Build_L, Control_Wave_PC.py:86-179
//                      (pc=True branches), Dirichlet rows as identity.
//   pd_rhs_launch    : Build_f / Build_g / Build_Initial_Condition (:48-83)
//                      folded through the residual (:118, :139, :144, :93-95).
//   pd_gmres         : KSPGMRES as configured at :347-359 (left PC, classical
//                      Gram-Schmidt without refinement, zero initial guess,
//                      preconditioned-residual test relative to ||P^-1 b||).
#include <math.h>
#include <string.h>

#include <vector>

#include "pd_common.cuh"

// --------------------------------------------------------------------- matvec
struct OpParams {
  int n, N_t;      // n: global node count
  int nloc, j0;    // local rows and the global index of local row 0 (slab mode; else n, 0)
  const cplx* halo_lo;  // [2][N_t] rows j0-1 of (u, p), from the left neighbour (slab mode)
  const cplx* halo_hi;  // [2][N_t] rows j0+nloc
  double h, dt2h;  // dt^2 / 2
  double c;        // dt^2 / sqrt(gamma)
  double qlast;    // sqrt(gamma) if bug138 else 1
  int circulant;   // 1: the block-circulant operator P the preconditioner inverts
  int64_t plane;
};

// scalar helpers so that the stencil kernels exist for complex128 and for float64 vectors
__device__ __forceinline__ double vzero(double) { return 0.0; }
__device__ __forceinline__ cplx vzero(cplx) { return cmake(0, 0); }
__device__ __forceinline__ double vsub(double a, double b) { return a - b; }
__device__ __forceinline__ cplx vsub(cplx a, cplx b) { return csub(a, b); }
__device__ __forceinline__ double vadd(double a, double b) { return a + b; }
__device__ __forceinline__ cplx vadd(cplx a, cplx b) { return cadd(a, b); }
__device__ __forceinline__ double vscale(double a, double s) { return a * s; }
__device__ __forceinline__ cplx vscale(cplx a, double s) { return cscale(a, s); }
// w_off (a + c) + w_dia b
__device__ __forceinline__ double vlin3(double a, double b, double c, double wo, double wd) { return wo * (a + c) + wd * b; }
__device__ __forceinline__ cplx vlin3(cplx a, cplx b, cplx c, double wo, double wd) {
  return cmake(wo * (a.x + c.x) + wd * b.x, wo * (a.y + c.y) + wd * b.y);
}
// a + s1 b - s2 c   (the state row)  /  a + s1 b + s2 c (the adjoint row, pass -s2)
__device__ __forceinline__ double vcomb(double a, double s1, double b, double s2, double c) { return a + s1 * b - s2 * c; }
__device__ __forceinline__ cplx vcomb(cplx a, double s1, cplx b, double s2, cplx c) {
  return cmake(a.x + s1 * b.x - s2 * c.x, a.y + s1 * b.y - s2 * c.y);
}

// v: local plane of field f; j: GLOBAL node index
template <class T>
__device__ __forceinline__ T ld_or_zero(const T* __restrict__ v, int f, int j, int i, const OpParams& op) {
  // Dirichlet columns are dropped: boundary-node values never enter interior rows
  if (j < 1 || j > op.n - 2) return vzero(T());
  if (i < 0 || i >= op.N_t) {
    if (!op.circulant) return vzero(T());
    i = i < 0 ? i + op.N_t : i - op.N_t;  // C1, C2 wrap around (mat_test.ipynb cells 8-9)
  }
  const int jl = j - op.j0;
  if (jl < 0) return reinterpret_cast<const T*>(op.halo_lo)[(int64_t)f * op.N_t + i];
  if (jl >= op.nloc) return reinterpret_cast<const T*>(op.halo_hi)[(int64_t)f * op.N_t + i];
  return v[(int64_t)jl * op.N_t + i];
}

#define PD_MV_TJ 16  // nodes per thread: rows j-1, j, j+1 slide through registers, each row is loaded once per tile

template <class T>
__global__ void __launch_bounds__(256)
pd_matvec_kernel(const T* __restrict__ x, T* __restrict__ y, OpParams op) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= op.N_t) return;
  const T* u = x;
  const T* p = x + op.plane;
  const double m_off = op.h / 6.0, m_dia = 2.0 * op.h / 3.0;
  const double ih = 1.0 / op.h;
  const double d_i = (i == 0 && !op.circulant) ? 0.5 : 1.0;               // :117
  const double e_i = (i == op.N_t - 1 && !op.circulant) ? 0.5 : 1.0;      // :143
  const double q_i = (i == op.N_t - 1 && !op.circulant) ? op.qlast : 1.0; // :138
  // Rounding matters here: the Krylov vectors of this problem are smooth, so both the second
  // difference in time and the stiffness stencil in space cancel to O(dt^2), O(h^2) of their operands,
  // and the ill-conditioned preconditioner amplifies whatever noise the matvec leaves (measured: extra
  // GMRES iterations at N_x >= 4096).  Every cancelling combination is therefore formed from
  // differences of neighbouring values -- exact in floating point (Sterbenz) -- before any scaling:
  //   time:   D2 v = (v_i - v_{i-1}) - (v_{i-1} - v_{i-2})        per node, then M in space
  //   space:  K v  = ((v_C - v_L) + (v_C - v_R)) / h               per time level, then summed
  const int jl0 = blockIdx.y * PD_MV_TJ;
  const int jl1 = min(jl0 + PD_MV_TJ, op.nloc);
  T uv[3][3], pv[3][3];  // [node L,C,R][time level 0,1,2]: u at i-t, p at i+t
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      uv[s + 1][t] = ld_or_zero<T>(u, 0, op.j0 + jl0 - 1 + s, i - t, op);
      pv[s + 1][t] = ld_or_zero<T>(p, 1, op.j0 + jl0 - 1 + s, i + t, op);
    }
  for (int jl = jl0; jl < jl1; ++jl) {
    const int j = op.j0 + jl;  // global node
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      uv[0][t] = uv[1][t]; uv[1][t] = uv[2][t];
      pv[0][t] = pv[1][t]; pv[1][t] = pv[2][t];
      uv[2][t] = ld_or_zero<T>(u, 0, j + 1, i - t, op);
      pv[2][t] = ld_or_zero<T>(p, 1, j + 1, i + t, op);
    }
    const int64_t o = (int64_t)jl * op.N_t + i;
    if (j == 0 || j == op.n - 1) {  // Dirichlet rows: identity (the centre value bypasses the column mask)
      y[o] = u[o];
      y[op.plane + o] = p[o];
      continue;
    }
    T d2u[3], d2p[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      d2u[s] = vsub(vsub(uv[s][0], uv[s][1]), vsub(uv[s][1], uv[s][2]));
      d2p[s] = vsub(vsub(pv[s][0], pv[s][1]), vsub(pv[s][1], pv[s][2]));
    }
    const T Ku0 = vscale(vadd(vsub(uv[1][0], uv[0][0]), vsub(uv[1][0], uv[2][0])), ih);
    const T Ku2 = vscale(vadd(vsub(uv[1][2], uv[0][2]), vsub(uv[1][2], uv[2][2])), ih);
    const T Kp0 = vscale(vadd(vsub(pv[1][0], pv[0][0]), vsub(pv[1][0], pv[2][0])), ih);
    const T Kp2 = vscale(vadd(vsub(pv[1][2], pv[0][2]), vsub(pv[1][2], pv[2][2])), ih);
    const T Md2u = vlin3(d2u[0], d2u[1], d2u[2], m_off, m_dia);
    const T Md2p = vlin3(d2p[0], d2p[1], d2p[2], m_off, m_dia);
    const T Mu0 = vlin3(uv[0][0], uv[1][0], uv[2][0], m_off, m_dia);
    const T Mp0 = vlin3(pv[0][0], pv[1][0], pv[2][0], m_off, m_dia);
    // state row: M(u_i - 2u_{i-1} + u_{i-2}) + q dt^2/2 K(u_i + u_{i-2}) - d c M p_i
    y[o] = vcomb(Md2u, q_i * op.dt2h, vadd(Ku0, Ku2), d_i * op.c, Mp0);
    // adjoint row: e c M u_i + M(p_i - 2p_{i+1} + p_{i+2}) + dt^2/2 K(p_i + p_{i+2})
    y[op.plane + o] = vcomb(Md2p, op.dt2h, vadd(Kp0, Kp2), -e_i * op.c, Mu0);
  }
}

int pd_matvec_launch(pd_handle* h, const cplx* x, cplx* y, cudaStream_t st, int circulant,
                     const cplx* halo_lo, const cplx* halo_hi, int real_vectors) {
  OpParams op;
  op.circulant = circulant;
  op.n = h->cfg.N_x + 1; op.N_t = h->cfg.N_t; op.h = h->h; op.dt2h = 0.5 * h->dt * h->dt; op.c = h->c;
  op.nloc = h->n; op.j0 = h->node_begin; op.halo_lo = halo_lo; op.halo_hi = halo_hi;
  op.qlast = h->cfg.bug138 ? sqrt(h->cfg.gamma) : 1.0;
  op.plane = (int64_t)h->n * h->cfg.N_t;
  if (h->slab_count > 1 && ((h->slab_rank > 0 && !halo_lo) || (h->slab_rank < h->slab_count - 1 && !halo_hi))) {
    pd_set_error("matvec in slab mode needs the neighbour rows (halo_lo / halo_hi)");
    return PD_ERR_INVALID;
  }
  dim3 grid((op.N_t + 255) / 256, (h->n + PD_MV_TJ - 1) / PD_MV_TJ);
  if (real_vectors)
    pd_matvec_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(x), reinterpret_cast<double*>(y), op);
  else
    pd_matvec_kernel<cplx><<<grid, 256, 0, st>>>(x, y, op);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// ------------------------------------------------------------- (A - P) x, the residual-correction operator
// A (Build_L, :86-179) and the block-circulant P that DiagFFTPC inverts differ only where the time stencils
// wrap around and in the three special factors (:117, :143, :138):
//   state row i   : + 2 M u_{i-1+N} [i = 0],  - (M + dt^2/2 K) u_{i-2+N} [i = 0, 1],  + c/2 M p_0 [i = 0],
//                   + (q - 1) dt^2/2 K (u_i + u_{i-2}) [i = N-1, q = sqrt(gamma) with bug138]
//   adjoint row i : + 2 M p_{i+1-N} [i = N-1],  - (M + dt^2/2 K) p_{i+2-N} [i = N-2, N-1],  - c/2 M u_{N-1} [i = N-1]
// so (A - P) x is non-zero on at most three time levels per field.  With it the preconditioned operator is
//   P^-1 A v = v + P^-1 (A - P) v      (v with zero Dirichlet rows, as every Krylov vector is)
// which never forms the cancelling second differences of A v: the rounding noise of the matvec -- amplified by the
// ill-conditioned P^-1, it is what sets the GMRES iteration count at N_x >= 2000 -- is gone.
// One thread per (node, slot); slots 0..2 = state rows i = 0, 1, N-1, slots 3..4 = adjoint rows i = N-2, N-1.  The
// output vector must be zero everywhere else (the caller keeps one such vector: only these entries are ever written).
template <class T>
__global__ void __launch_bounds__(128)
pd_delta_kernel(const T* __restrict__ x, T* __restrict__ d, OpParams op) {
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = blockIdx.y;
  if (jl >= op.nloc) return;
  const int N = op.N_t;
  const int j = op.j0 + jl;
  const bool state = slot < 3;
  const int i = slot == 0 ? 0 : slot == 1 ? 1 : slot == 2 ? N - 1 : slot == 3 ? N - 2 : N - 1;
  // duplicates for tiny N_t (N = 3: state rows {0, 1, 2}, adjoint rows {1, 2} are all distinct; nothing to skip)
  if (state && slot == 2 && i <= 1) return;
  if (!state && slot == 3 && i == N - 1) return;
  const int64_t o = (int64_t)jl * N + i;
  T* out = d + (state ? 0 : op.plane);
  if (j == 0 || j == op.n - 1) {  // Dirichlet rows: A and P are both the identity there
    out[o] = vzero(T());
    return;
  }
  const T* u = x;
  const T* p = x + op.plane;
  const double m_off = op.h / 6.0, m_dia = 2.0 * op.h / 3.0, ih = 1.0 / op.h;
  auto Mv = [&](const T* v, int f, int t) {
    return vlin3(ld_or_zero<T>(v, f, j - 1, t, op), ld_or_zero<T>(v, f, j, t, op), ld_or_zero<T>(v, f, j + 1, t, op),
                 m_off, m_dia);
  };
  auto Kv = [&](const T* v, int f, int t) {
    const T c = ld_or_zero<T>(v, f, j, t, op);
    return vscale(vadd(vsub(c, ld_or_zero<T>(v, f, j - 1, t, op)), vsub(c, ld_or_zero<T>(v, f, j + 1, t, op))), ih);
  };
  T acc = vzero(T());
  if (state) {
    if (i - 1 < 0) acc = vadd(acc, vscale(Mv(u, 0, i - 1 + N), 2.0));
    if (i - 2 < 0) acc = vsub(acc, vadd(Mv(u, 0, i - 2 + N), vscale(Kv(u, 0, i - 2 + N), op.dt2h)));
    if (i == 0) acc = vadd(acc, vscale(Mv(p, 1, 0), 0.5 * op.c));
    if (i == N - 1 && op.qlast != 1.0) {
      T k2 = Kv(u, 0, i);
      if (i - 2 >= 0) k2 = vadd(k2, Kv(u, 0, i - 2));
      acc = vadd(acc, vscale(k2, (op.qlast - 1.0) * op.dt2h));
    }
  } else {
    if (i + 1 >= N) acc = vadd(acc, vscale(Mv(p, 1, i + 1 - N), 2.0));
    if (i + 2 >= N) acc = vsub(acc, vadd(Mv(p, 1, i + 2 - N), vscale(Kv(p, 1, i + 2 - N), op.dt2h)));
    if (i == N - 1) acc = vsub(acc, vscale(Mv(u, 0, N - 1), 0.5 * op.c));
  }
  out[o] = acc;
}

int pd_delta_launch(pd_handle* h, const cplx* x, cplx* d, cudaStream_t st, const cplx* halo_lo, const cplx* halo_hi,
                    int real_vectors) {
  OpParams op;
  op.circulant = 0;
  op.n = h->cfg.N_x + 1; op.N_t = h->cfg.N_t; op.h = h->h; op.dt2h = 0.5 * h->dt * h->dt; op.c = h->c;
  op.nloc = h->n; op.j0 = h->node_begin; op.halo_lo = halo_lo; op.halo_hi = halo_hi;
  op.qlast = h->cfg.bug138 ? sqrt(h->cfg.gamma) : 1.0;
  op.plane = (int64_t)h->n * h->cfg.N_t;
  if (h->slab_count > 1 && ((h->slab_rank > 0 && !halo_lo) || (h->slab_rank < h->slab_count - 1 && !halo_hi))) {
    pd_set_error("(A - P) x in slab mode needs the neighbour rows (halo_lo / halo_hi)");
    return PD_ERR_INVALID;
  }
  dim3 grid((h->n + 127) / 128, 5);
  if (real_vectors)
    pd_delta_kernel<double><<<grid, 128, 0, st>>>(reinterpret_cast<const double*>(x), reinterpret_cast<double*>(d), op);
  else
    pd_delta_kernel<cplx><<<grid, 128, 0, st>>>(x, d, op);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// ------------------------------------------------------------------------ rhs
struct RhsParams {
  int n, N_t, N_x;  // n: global node count
  int j0;           // global index of local row 0
  double h, dt, T, gamma;
  int64_t plane;
};

__device__ __forceinline__ void rhs_store(cplx* b, int64_t o, double v) { b[o] = cmake(v, 0.0); }
__device__ __forceinline__ void rhs_store(double* b, int64_t o, double v) { b[o] = v; }

template <class T>
__global__ void __launch_bounds__(256)
pd_rhs_kernel(T* __restrict__ b, RhsParams rp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = rp.j0 + blockIdx.y;
  if (i >= rp.N_t) return;
  const int64_t o = (int64_t)blockIdx.y * rp.N_t + i;
  if (j == 0 || j == rp.n - 1) {
    rhs_store(b, o, 0.0);
    rhs_store(b, rp.plane + o, 0.0);
    return;
  }
  // nodal sin(pi x) at j-1, j, j+1 (full vectors: boundary nodal values enter M f)
  const double sl = sinpi((double)(j - 1) / rp.N_x), sc = sinpi((double)j / rp.N_x),
               sr = sinpi((double)(j + 1) / rp.N_x);
  const double Ms = rp.h / 6.0 * (sl + 4.0 * sc + sr);
  const double Ks = (-sl + 2.0 * sc - sr) / rp.h;
  const double sg = sqrt(rp.gamma), dt2 = rp.dt * rp.dt, eT = exp(rp.T);
  // f_i at t = i dt, scaled by sqrt(gamma) (:55-57); g_i at t = (i+1) dt (:69-72)
  const double tf = i * rp.dt, tg = (i + 1) * rp.dt;
  const double ef = exp(tf) - eT;
  const double F = -(1.0 / rp.gamma) * ef * ef * sg;
  const double eg = exp(tg) - eT;
  const double G = 2.0 * (2.0 * exp(2.0 * tg) - exp(rp.T + tg)) + M_PI * M_PI * eg * eg + cospi(tg);
  double bu;
  if (i == 0) {
    bu = dt2 * Ms * (0.5 * F + sg / dt2);                 // :118 (u_1 = 0, :80)
  } else {
    bu = dt2 * Ms * F;                                    // :139, :159
    if (i == 1) bu += -Ms * sg - 0.5 * dt2 * Ks * sg;     // :93-95 with :157-158
  }
  double bp = dt2 * Ms * G;                               // :123, :164
  if (i == rp.N_t - 1) bp *= 0.5;                         // :144
  rhs_store(b, o, bu);
  rhs_store(b, rp.plane + o, bp);
}

int pd_rhs_launch(pd_handle* h, cplx* b, cudaStream_t st, int real_vectors) {
  RhsParams rp;
  rp.n = h->cfg.N_x + 1; rp.j0 = h->node_begin; rp.N_t = h->cfg.N_t; rp.N_x = h->cfg.N_x; rp.h = h->h; rp.dt = h->dt;
  rp.T = h->cfg.T; rp.gamma = h->cfg.gamma; rp.plane = (int64_t)h->n * h->cfg.N_t;
  dim3 grid((rp.N_t + 255) / 256, h->n);
  if (real_vectors)
    pd_rhs_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<double*>(b), rp);
  else
    pd_rhs_kernel<cplx><<<grid, 256, 0, st>>>(b, rp);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// ------------------------------------------------------------- BLAS-1 kernels
#define PD_RED_THREADS 256
#define PD_MAXB 16  // basis vectors handled per launch (w is re-read once per batch)

struct VecBatch {
  const cplx* v[PD_MAXB];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[blk * NV + i] = sum over this block's elements of conj(V_i[e]) * w[e]
template <int NV>
__global__ void __launch_bounds__(PD_RED_THREADS)
pd_mdot_kernel(VecBatch vb, const cplx* __restrict__ w, int64_t len, cplx* __restrict__ partial) {
  double ar[NV], ai[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ar[i] = ai[i] = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len;
       e += (int64_t)gridDim.x * blockDim.x) {
    const cplx we = w[e];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const cplx ve = vb.v[i][e];
      ar[i] += ve.x * we.x + ve.y * we.y;
      ai[i] += ve.x * we.y - ve.y * we.x;
    }
  }
  __shared__ double sr[NV][PD_RED_THREADS / 32], si[NV][PD_RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double r = warp_sum(ar[i]), im = warp_sum(ai[i]);
    if (lane == 0) { sr[i][wid] = r; si[i][wid] = im; }
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double r = 0, im = 0;
    for (int k = 0; k < PD_RED_THREADS / 32; ++k) { r += sr[threadIdx.x][k]; im += si[threadIdx.x][k]; }
    partial[(int64_t)blockIdx.x * NV + threadIdx.x] = cmake(r, im);
  }
}

// out[i] = sum_blk partial[blk * nv + i]   (fixed order: deterministic)
__global__ void pd_reduce_partials_kernel(const cplx* __restrict__ partial, int nblk, int nv,
                                          cplx* __restrict__ out, int real_only) {
  const int i = blockIdx.x;
  double r = 0, im = 0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    cplx p = partial[(int64_t)b * nv + i];
    r += p.x; im += p.y;
  }
  __shared__ double sr[32], si[32];
  r = warp_sum(r); im = warp_sum(im);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sr[wid] = r; si[wid] = im; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { tr += sr[k]; ti += si[k]; }
    out[i] = cmake(tr, real_only ? 0.0 : ti);
  }
}

// w += sign * sum_i coef[i] * V_i ; optionally partial[blk] = sum |w_new|^2
template <int NV, bool NORM>
__global__ void __launch_bounds__(PD_RED_THREADS)
pd_maxpy_kernel(VecBatch vb, const cplx* __restrict__ coef, double sign, cplx* __restrict__ w,
                int64_t len, cplx* __restrict__ partial) {
  cplx cf[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) cf[i] = cscale(coef[i], sign);
  double acc = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len;
       e += (int64_t)gridDim.x * blockDim.x) {
    cplx we = w[e];
#pragma unroll
    for (int i = 0; i < NV; ++i) we = cfma(cf[i], vb.v[i][e], we);
    w[e] = we;
    if (NORM) acc += we.x * we.x + we.y * we.y;
  }
  if (NORM) {
    __shared__ double sr[PD_RED_THREADS / 32];
    double r = warp_sum(acc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sr[wid] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int k = 0; k < PD_RED_THREADS / 32; ++k) t += sr[k];
      partial[blockIdx.x] = cmake(t, 0.0);
    }
  }
}

// y = a*x + b*y elementwise with real scalars (b = 0: y = a*x)
__global__ void __launch_bounds__(256)
pd_axpby_kernel(double a, const cplx* __restrict__ x, double b, cplx* __restrict__ y, int64_t len) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len;
       e += (int64_t)gridDim.x * blockDim.x) {
    cplx xe = x[e];
    if (b == 0.0) {
      y[e] = cscale(xe, a);
    } else {
      cplx ye = y[e];
      y[e] = cmake(a * xe.x + b * ye.x, a * xe.y + b * ye.y);
    }
  }
}

// v *= 1/sqrt(norm2[0].x)
__global__ void __launch_bounds__(256)
pd_normalize_kernel(cplx* __restrict__ v, const cplx* __restrict__ norm2, int64_t len) {
  const double s = rsqrt(norm2[0].x);
  const double inv = s;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len;
       e += (int64_t)gridDim.x * blockDim.x)
    v[e] = cscale(v[e], inv);
}

static int red_blocks(const pd_handle* h, int64_t len) {
  int64_t nb = (len + PD_RED_THREADS - 1) / PD_RED_THREADS;
  int64_t cap = (int64_t)h->num_sms * 8;
  return (int)(nb < cap ? nb : cap);
}

template <int NV>
static int mdot_batch(pd_handle* h, const cplx* const* vs, const cplx* w, int64_t len, cplx* out,
                      cudaStream_t st) {
  VecBatch vb;
  for (int i = 0; i < PD_MAXB; ++i) vb.v[i] = vs[i < NV ? i : 0];
  const int nb = red_blocks(h, len);
  pd_mdot_kernel<NV><<<nb, PD_RED_THREADS, 0, st>>>(vb, w, len, h->kry_partial);
  PD_CHECK_LAUNCH();
  pd_reduce_partials_kernel<<<NV, 256, 0, st>>>(h->kry_partial, nb, NV, out, h->kry_real);
  PD_CHECK_LAUNCH();
  h->launches += 2;
  return PD_OK;
}

static int ensure_partial(pd_handle* h) {
  if (!h->kry_partial) {
    size_t bytes = sizeof(cplx) * (size_t)h->num_sms * 8 * PD_MAXB;
    PD_CUDA(cudaMalloc(&h->kry_partial, bytes));
    h->ws_bytes += bytes;
  }
  return PD_OK;
}

// out[i] = V_i^H w for i < nv, vectors given by pointer list
static int mdot_list(pd_handle* h, const cplx* const* vs, int nv, const cplx* w, int64_t len, cplx* out,
                     cudaStream_t st) {
  int rc = ensure_partial(h);
  if (rc) return rc;
  for (int base = 0; base < nv; base += PD_MAXB) {
    const int cnt = nv - base < PD_MAXB ? nv - base : PD_MAXB;
    const cplx* const* v = vs + base;
    switch (cnt) {
      case 1: rc = mdot_batch<1>(h, v, w, len, out + base, st); break;
      case 2: rc = mdot_batch<2>(h, v, w, len, out + base, st); break;
      case 3: rc = mdot_batch<3>(h, v, w, len, out + base, st); break;
      case 4: rc = mdot_batch<4>(h, v, w, len, out + base, st); break;
      case 5: rc = mdot_batch<5>(h, v, w, len, out + base, st); break;
      case 6: rc = mdot_batch<6>(h, v, w, len, out + base, st); break;
      case 7: rc = mdot_batch<7>(h, v, w, len, out + base, st); break;
      case 8: rc = mdot_batch<8>(h, v, w, len, out + base, st); break;
      case 9: rc = mdot_batch<9>(h, v, w, len, out + base, st); break;
      case 10: rc = mdot_batch<10>(h, v, w, len, out + base, st); break;
      case 11: rc = mdot_batch<11>(h, v, w, len, out + base, st); break;
      case 12: rc = mdot_batch<12>(h, v, w, len, out + base, st); break;
      case 13: rc = mdot_batch<13>(h, v, w, len, out + base, st); break;
      case 14: rc = mdot_batch<14>(h, v, w, len, out + base, st); break;
      case 15: rc = mdot_batch<15>(h, v, w, len, out + base, st); break;
      default: rc = mdot_batch<16>(h, v, w, len, out + base, st); break;
    }
    if (rc) return rc;
  }
  return PD_OK;
}

template <int NV>
static int maxpy_batch(pd_handle* h, const cplx* const* vs, const cplx* coef, double sign, cplx* w,
                       int64_t len, bool norm, cplx* norm_out, cudaStream_t st) {
  VecBatch vb;
  for (int i = 0; i < PD_MAXB; ++i) vb.v[i] = vs[i < NV ? i : 0];
  const int nb = red_blocks(h, len);
  if (norm) {
    pd_maxpy_kernel<NV, true><<<nb, PD_RED_THREADS, 0, st>>>(vb, coef, sign, w, len, h->kry_partial);
    PD_CHECK_LAUNCH();
    pd_reduce_partials_kernel<<<1, 256, 0, st>>>(h->kry_partial, nb, 1, norm_out, 0);
    PD_CHECK_LAUNCH();
    h->launches += 2;
  } else {
    pd_maxpy_kernel<NV, false><<<nb, PD_RED_THREADS, 0, st>>>(vb, coef, sign, w, len, h->kry_partial);
    PD_CHECK_LAUNCH();
    h->launches += 1;
  }
  return PD_OK;
}

// w += sign * sum_i coef[i] V_i ; if norm_out: norm_out[0] = ||w_new||^2
static int maxpy_list(pd_handle* h, const cplx* const* vs, int nv, const cplx* coef, double sign, cplx* w,
                      int64_t len, cplx* norm_out, cudaStream_t st) {
  int rc = ensure_partial(h);
  if (rc) return rc;
  for (int base = 0; base < nv; base += PD_MAXB) {
    const int cnt = nv - base < PD_MAXB ? nv - base : PD_MAXB;
    const bool lastb = base + cnt >= nv;
    const bool norm = lastb && norm_out != nullptr;
    const cplx* const* v = vs + base;
    const cplx* cf = coef + base;
    switch (cnt) {
      case 1: rc = maxpy_batch<1>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 2: rc = maxpy_batch<2>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 3: rc = maxpy_batch<3>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 4: rc = maxpy_batch<4>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 5: rc = maxpy_batch<5>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 6: rc = maxpy_batch<6>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 7: rc = maxpy_batch<7>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 8: rc = maxpy_batch<8>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 9: rc = maxpy_batch<9>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 10: rc = maxpy_batch<10>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 11: rc = maxpy_batch<11>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 12: rc = maxpy_batch<12>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 13: rc = maxpy_batch<13>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 14: rc = maxpy_batch<14>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      case 15: rc = maxpy_batch<15>(h, v, cf, sign, w, len, norm, norm_out, st); break;
      default: rc = maxpy_batch<16>(h, v, cf, sign, w, len, norm, norm_out, st); break;
    }
    if (rc) return rc;
  }
  return PD_OK;
}

extern "C" int pd_mdot(pd_handle* h, const void* V_dev, int64_t ld, int nv, const void* w_dev, int64_t len,
                       void* out_dev, void* stream) {
  if (!h || !V_dev || !w_dev || !out_dev || nv < 0) {
    pd_set_error("pd_mdot: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  std::vector<const cplx*> vs(nv);
  for (int i = 0; i < nv; ++i) vs[i] = (const cplx*)V_dev + (int64_t)i * ld;
  h->kry_real = h->opt_kry_real;  // float64 vectors seen as complex pairs: only the real part of the sums is meaningful
  return mdot_list(h, vs.data(), nv, (const cplx*)w_dev, len, (cplx*)out_dev, (cudaStream_t)stream);
}

// w += sign * sum_i coef[i] V_i (coefficients on the device); norm2_out (optional) <- ||w_new||^2 (local)
extern "C" int pd_maxpy(pd_handle* h, const void* V_dev, int64_t ld, int nv, const void* coef_dev, double sign,
                        void* w_dev, int64_t len, void* norm2_out_dev, void* stream) {
  if (!h || !V_dev || !w_dev || !coef_dev || nv < 1) {
    pd_set_error("pd_maxpy: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  std::vector<const cplx*> vs(nv);
  for (int i = 0; i < nv; ++i) vs[i] = (const cplx*)V_dev + (int64_t)i * ld;
  return maxpy_list(h, vs.data(), nv, (const cplx*)coef_dev, sign, (cplx*)w_dev, len, (cplx*)norm2_out_dev,
                    (cudaStream_t)stream);
}

// ---------------------------------------------------------------------- GMRES
struct hcplx {
  double re, im;
};
static inline hcplx hmul(hcplx a, hcplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline hcplx hconj(hcplx a) { return {a.re, -a.im}; }
static inline hcplx hadd(hcplx a, hcplx b) { return {a.re + b.re, a.im + b.im}; }
static inline hcplx hsub(hcplx a, hcplx b) { return {a.re - b.re, a.im - b.im}; }
static inline double habs(hcplx a) { return hypot(a.re, a.im); }
static inline hcplx hdiv(hcplx a, hcplx b) {
  double d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// The small host-side part of restarted GMRES as KSPGMRES keeps it (Control_Wave_PC.py:347-359 selects KSPGMRES with
// classical Gram-Schmidt): the Hessenberg matrix of one cycle, reduced to triangular form column by column with
// Givens rotations, the rotated right-hand side g whose last entry is the residual-norm estimate, and the back
// substitution at the end of the cycle.  ONE implementation serves pd_gmres / pd_gmres_real here and, through the
// pd_hess_* entry points, the distributed Krylov loop of dist.py (which owns the collectives but not this algebra).
struct pd_hessenberg {
  int restart;
  int ncol;                   // columns pushed in the current cycle
  std::vector<hcplx> H;       // column-major, stride restart + 1; grows with the columns pushed (restart = 300
                              // would otherwise cost a 1.4 MB zero-fill per solve)
  std::vector<hcplx> g, cs, sn;
  explicit pd_hessenberg(int m) : restart(m), ncol(0), g(m + 1), cs(m), sn(m) {}
  hcplx& at(int i, int j) { return H[(size_t)j * (restart + 1) + i]; }
  void start(double beta) {
    for (auto& e : g) e = {0, 0};
    g[0] = {beta, 0};
    ncol = 0;
  }
  // column j = ncol: hcol[0..j] = the Gram-Schmidt coefficients, hcol[j + 1].re = SQUARED norm of the orthogonalised
  // vector.  Returns the residual-norm estimate |g_{j+1}|; *hnorm = h_{j+1,j}.
  double push(const hcplx* hcol, double* hnorm) {
    const int j = ncol;
    if (H.size() < (size_t)(j + 1) * (restart + 1)) H.resize((size_t)(j + 1) * (restart + 1));
    for (int i = 0; i <= j; ++i) at(i, j) = hcol[i];
    const double hn = sqrt(fmax(hcol[j + 1].re, 0.0));
    at(j + 1, j) = {hn, 0};
    for (int i = 0; i < j; ++i) {
      const hcplx a_ = at(i, j), b_ = at(i + 1, j);
      at(i, j) = hadd(hmul(hconj(cs[i]), a_), hmul(hconj(sn[i]), b_));
      at(i + 1, j) = hsub(hmul(cs[i], b_), hmul(sn[i], a_));
    }
    const hcplx a_ = at(j, j), b_ = at(j + 1, j);
    const double den = sqrt(a_.re * a_.re + a_.im * a_.im + b_.re * b_.re + b_.im * b_.im);
    if (den == 0.0) { cs[j] = {1, 0}; sn[j] = {0, 0}; }
    else { cs[j] = {a_.re / den, a_.im / den}; sn[j] = {b_.re / den, b_.im / den}; }
    at(j, j) = hadd(hmul(hconj(cs[j]), a_), hmul(hconj(sn[j]), b_));
    at(j + 1, j) = {0, 0};
    const hcplx gj = g[j];
    g[j + 1] = hmul({-sn[j].re, -sn[j].im}, gj);
    g[j] = hmul(hconj(cs[j]), gj);
    ncol = j + 1;
    if (hnorm) *hnorm = hn;
    return habs(g[j + 1]);
  }
  // y = H^-1 g for the ncol columns of this cycle (upper triangular after the rotations)
  void solve(hcplx* y) {
    for (int i = ncol - 1; i >= 0; --i) {
      hcplx s = g[i];
      for (int k = i + 1; k < ncol; ++k) s = hsub(s, hmul(at(i, k), y[k]));
      y[i] = hdiv(s, at(i, i));
    }
  }
};

extern "C" int pd_hess_create(int restart, pd_hessenberg** out) {
  if (restart < 1 || !out) {
    pd_set_error("pd_hess_create: invalid argument");
    return PD_ERR_INVALID;
  }
  *out = new pd_hessenberg(restart);
  return PD_OK;
}
extern "C" int pd_hess_destroy(pd_hessenberg* q) {
  delete q;
  return PD_OK;
}
extern "C" int pd_hess_start(pd_hessenberg* q, double beta) {
  if (!q) {
    pd_set_error("pd_hess_start: invalid argument");
    return PD_ERR_INVALID;
  }
  q->start(beta);
  return PD_OK;
}
extern "C" int pd_hess_push(pd_hessenberg* q, const void* hcol, double* resnorm_out, double* hnorm_out) {
  if (!q || !hcol || !resnorm_out || q->ncol >= q->restart) {
    pd_set_error("pd_hess_push: invalid argument or cycle full");
    return PD_ERR_INVALID;
  }
  *resnorm_out = q->push(reinterpret_cast<const hcplx*>(hcol), hnorm_out);
  return PD_OK;
}
extern "C" int pd_hess_solve(pd_hessenberg* q, void* y_out, int* ncol_out) {
  if (!q || !y_out) {
    pd_set_error("pd_hess_solve: invalid argument");
    return PD_ERR_INVALID;
  }
  q->solve(reinterpret_cast<hcplx*>(y_out));
  if (ncol_out) *ncol_out = q->ncol;
  return PD_OK;
}

static int ensure_basis(pd_handle* h, std::vector<cplx*>& V, int need, int64_t len) {
  // basis vectors are cached on the handle as one allocation each
  while ((int)V.size() < need) {
    cplx* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(cplx) * (size_t)len);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return PD_ERR_NOMEM;
    }
    h->ws_bytes += sizeof(cplx) * (size_t)len;
    V.push_back(p);
  }
  return PD_OK;
}

struct KrylovCache {
  std::vector<cplx*> V;
  int64_t veclen = 0;  // complex elements per cached vector
};

static KrylovCache* cache_of(pd_handle* h) {
  if (!h->kry_V) h->kry_V = reinterpret_cast<cplx*>(new KrylovCache());
  return reinterpret_cast<KrylovCache*>(h->kry_V);
}

void pd_krylov_free(pd_handle* h) {
  if (h->kry_V) {
    KrylovCache* kc = reinterpret_cast<KrylovCache*>(h->kry_V);
    for (cplx* p : kc->V) cudaFree(p);
    delete kc;
    h->kry_V = nullptr;
  }
  if (h->kry_partial) cudaFree(h->kry_partial);
  if (h->kry_d) cudaFree(h->kry_d);
  h->kry_d = nullptr;
  h->kry_d_mode = 0;
  if (h->kry_h) cudaFree(h->kry_h);
  if (h->kry_t) cudaFree(h->kry_t);
  if (h->kry_host) cudaFreeHost(h->kry_host);
  h->kry_partial = h->kry_h = h->kry_t = nullptr;
  h->kry_host = nullptr;
}

static int gmres_impl(pd_handle* h, const void* b_dev, void* x_dev, double rtol, double atol, int restart,
                      int max_it, int* its_out, double* hist, int* reason_out, void* stream, int real_vectors) {
  if (!h || !b_dev || !x_dev || restart < 1 || max_it < 0) {
    pd_set_error("pd_gmres: invalid argument");
    return PD_ERR_INVALID;
  }
  if (h->kcount != h->cfg.N_t || h->nloc != h->n) {
    pd_set_error("pd_gmres: handle is sharded (k_count/n_local set); use the stage API");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  // real vectors: 2 n N_t doubles, handled by the BLAS-1 kernels as n N_t complex pairs (the real part of
  // a pair-wise conj(v) w sum is the real inner product; imaginary parts are zeroed in the reduction)
  const int64_t len = (real_vectors ? 1 : 2) * (int64_t)h->n * h->cfg.N_t;
  h->kry_real = real_vectors;
  auto PC = [&](const cplx* in, cplx* out) {
    return real_vectors ? pd_pc_apply_real(h, in, out, stream) : pd_pc_apply(h, in, out, stream);
  };
  const cplx* b = (const cplx*)b_dev;
  cplx* x = (cplx*)x_dev;
  KrylovCache* kc = cache_of(h);
  if (kc->veclen < len) {  // cached basis vectors from a smaller (real) solve: start over
    for (cplx* p : kc->V) cudaFree(p);
    kc->V.clear();
    kc->veclen = len;
  }
  const int64_t alloc_len = kc->veclen;
  int rc;
  if ((rc = ensure_partial(h))) return rc;
  const int hcap = restart + 2;
  if (!h->kry_h || h->kry_cap < hcap) {
    if (h->kry_h) cudaFree(h->kry_h);
    if (h->kry_host) cudaFreeHost(h->kry_host);
    PD_CUDA(cudaMalloc(&h->kry_h, sizeof(cplx) * (size_t)hcap));
    PD_CUDA(cudaMallocHost(&h->kry_host, sizeof(cplx) * (size_t)hcap));
    h->kry_cap = hcap;
  }
  if (!h->kry_t) {
    const size_t full = 2 * (size_t)h->n * h->cfg.N_t;  // sized for the complex solve
    PD_CUDA(cudaMalloc(&h->kry_t, sizeof(cplx) * full));
    h->ws_bytes += sizeof(cplx) * full;
  }
  cplx* t = h->kry_t;
  // residual-correction mode (opt-in, pd_set_option "gmres_residual_correction"): w = v + P^-1 (A - P) v
  const int correction = h->opt_gmres_correction;
  if (correction) {
    const size_t full = 2 * (size_t)h->n * h->cfg.N_t;
    if (!h->kry_d) {
      PD_CUDA(cudaMalloc(&h->kry_d, sizeof(cplx) * full));
      h->ws_bytes += sizeof(cplx) * full;
      h->kry_d_mode = 0;
    }
    if (h->kry_d_mode != (real_vectors ? 2 : 1)) {  // the few non-zero entries sit elsewhere in the other layout
      PD_CUDA(cudaMemsetAsync(h->kry_d, 0, sizeof(cplx) * full, st));
      h->kry_d_mode = real_vectors ? 2 : 1;
    }
  }
  cplx* hdev = h->kry_h;
  hcplx* hhost = reinterpret_cast<hcplx*>(h->kry_host);
  const int nb1 = (int)((len + 255) / 256 < (int64_t)h->num_sms * 16 ? (len + 255) / 256
                                                                      : (int64_t)h->num_sms * 16);

  int its = 0, reason = -3;
  double beta0 = 0.0, target = 0.0;
  bool converged = false, first = true;
  PD_CUDA(cudaMemsetAsync(x, 0, sizeof(cplx) * (size_t)len, st));

  pd_hessenberg hess(restart);   // Hessenberg / Givens state of one cycle (shared with the distributed loop)
  std::vector<hcplx> yk(restart);

  while (!converged && (its < max_it || first)) {
    if ((rc = ensure_basis(h, kc->V, 1, alloc_len))) { pd_set_error("pd_gmres: out of device memory for the Krylov basis"); return rc; }
    cplx* v0 = kc->V[0];
    // r = P^-1 (b - A x)   (x = 0 on the first cycle)
    if (first) {
      if ((rc = PC(b, v0))) return rc;
    } else {
      if ((rc = pd_matvec_launch(h, x, t, st, 0, nullptr, nullptr, real_vectors))) return rc;
      pd_axpby_kernel<<<nb1, 256, 0, st>>>(1.0, b, -1.0, t, len);
      PD_CHECK_LAUNCH();
      h->launches++;
      if ((rc = PC(t, v0))) return rc;
    }
    const cplx* vs0[1] = {v0};
    if ((rc = mdot_list(h, vs0, 1, v0, len, hdev, st))) return rc;
    PD_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(cplx), cudaMemcpyDeviceToHost, st));
    PD_CUDA(cudaStreamSynchronize(st));
    const double beta = sqrt(hhost[0].re);
    if (first) {
      beta0 = beta;
      target = fmax(rtol * beta0, atol);
      if (hist) hist[0] = beta0;
      first = false;
      if (beta0 <= target || beta0 == 0.0) {
        converged = true;
        reason = beta0 <= atol ? 3 : 2;
        break;
      }
      if (max_it == 0) break;
    }
    pd_normalize_kernel<<<nb1, 256, 0, st>>>(v0, hdev, len);
    PD_CHECK_LAUNCH();
    h->launches++;
    hess.start(beta);
    int jdone = 0;
    for (int j = 0; j < restart; ++j) {
      // a basis slot for w; if memory runs out, restart early with what we have
      rc = ensure_basis(h, kc->V, j + 2, alloc_len);
      if (rc == PD_ERR_NOMEM) {
        if (j == 0) { pd_set_error("pd_gmres: out of device memory for the Krylov basis"); return rc; }
        break;
      }
      cplx* w = kc->V[j + 1];
      if (correction) {
        if ((rc = pd_delta_launch(h, kc->V[j], h->kry_d, st, nullptr, nullptr, real_vectors))) return rc;
        if ((rc = PC(h->kry_d, w))) return rc;
        pd_axpby_kernel<<<nb1, 256, 0, st>>>(1.0, kc->V[j], 1.0, w, len);
        PD_CHECK_LAUNCH();
        h->launches++;
      } else {
        if ((rc = pd_matvec_launch(h, kc->V[j], t, st, 0, nullptr, nullptr, real_vectors))) return rc;
        if ((rc = PC(t, w))) return rc;
      }
      // classical Gram-Schmidt: all inner products against the unmodified w first
      if ((rc = mdot_list(h, kc->V.data(), j + 1, w, len, hdev, st))) return rc;
      if ((rc = maxpy_list(h, kc->V.data(), j + 1, hdev, -1.0, w, len, hdev + (j + 1), st))) return rc;
      PD_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(cplx) * (size_t)(j + 2), cudaMemcpyDeviceToHost, st));
      PD_CUDA(cudaStreamSynchronize(st));
      double hn = 0.0;
      const double rn = hess.push(hhost, &hn);
      ++its;
      jdone = j + 1;
      if (hist) hist[its] = rn;
      if (rn <= target) {
        converged = true;
        reason = rn > atol ? 2 : 3;
        break;
      }
      if (its >= max_it || hn == 0.0) break;
      pd_normalize_kernel<<<nb1, 256, 0, st>>>(w, hdev + (j + 1), len);
      PD_CHECK_LAUNCH();
      h->launches++;
    }
    // y = H^-1 g (upper triangular), x += V y
    hess.solve(yk.data());
    if (jdone > 0) {
      for (int i = 0; i < jdone; ++i) hhost[i] = yk[i];
      PD_CUDA(cudaMemcpyAsync(hdev, hhost, sizeof(cplx) * (size_t)jdone, cudaMemcpyHostToDevice, st));
      if ((rc = maxpy_list(h, kc->V.data(), jdone, hdev, 1.0, x, len, nullptr, st))) return rc;
      PD_CUDA(cudaStreamSynchronize(st));
    }
    if (its >= max_it) break;
  }
  PD_CUDA(cudaStreamSynchronize(st));
  if (its_out) *its_out = its;
  if (reason_out) *reason_out = reason;
  if (!converged) {
    pd_set_error("pd_gmres: not converged after %d iterations", its);
    return PD_ERR_NOT_CONVERGED;
  }
  return PD_OK;
}

extern "C" int pd_gmres(pd_handle* h, const void* b_dev, void* x_dev, double rtol, double atol, int restart,
                        int max_it, int* its_out, double* hist, int* reason_out, void* stream) {
  return gmres_impl(h, b_dev, x_dev, rtol, atol, restart, max_it, its_out, hist, reason_out, stream, 0);
}

// Same solve on float64 vectors (2 n N_t doubles): real matvec, half-spectrum preconditioner, half the bytes
// in every BLAS-1 sweep.  Valid because b, A and P are real for this problem (the imaginary parts the complex
// solve carries are rounding noise).
extern "C" int pd_gmres_real(pd_handle* h, const void* b_dev, void* x_dev, double rtol, double atol, int restart,
                             int max_it, int* its_out, double* hist, int* reason_out, void* stream) {
  if (h && !pd_rfft_supported(h)) {
    pd_set_error("pd_gmres_real: needs N_t >= 8 (got %d); use pd_gmres", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  return gmres_impl(h, b_dev, x_dev, rtol, atol, restart, max_it, its_out, hist, reason_out, stream, 1);
}

// Run-time options of a handle (name, value).  Unknown names are an error.
//   "gmres_residual_correction" : 0 (default) pd_gmres forms P^-1 (A v) as KSP does; 1: v + P^-1 ((A - P) v)
//   "krylov_real_vectors"       : 0 (default); 1: pd_mdot treats its vectors as float64 data in complex pairs (the
//                                 imaginary parts of the sums are zeroed) -- the float64 distributed Krylov loop
//   "slab_overlap"              : 1 (default): pd_slab_apply runs its two frequency halves on two streams; 0: one stream
//   "host_register"             : 0 (default); 1: pd_pc_apply_host page-locks each host buffer once (see pd_capi.cu)
extern "C" int pd_set_option(pd_handle* h, const char* name, double value) {
  if (!h || !name) {
    pd_set_error("pd_set_option: invalid argument");
    return PD_ERR_INVALID;
  }
  if (!strcmp(name, "gmres_residual_correction")) {
    h->opt_gmres_correction = value != 0.0;
    return PD_OK;
  }
  if (!strcmp(name, "krylov_real_vectors")) {
    h->opt_kry_real = value != 0.0;
    return PD_OK;
  }
  if (!strcmp(name, "slab_overlap")) {
    h->opt_slab_no_overlap = value == 0.0 ? 1 : (value == 2.0 ? 2 : 0);  // 2: split, but on one stream (tests)
    return PD_OK;
  }
  if (!strcmp(name, "host_register")) {
    h->opt_host_register = value != 0.0;
    return PD_OK;
  }
  pd_set_error("pd_set_option: unknown option '%s'", name);
  return PD_ERR_INVALID;
}

// d = (A - P) x on device vectors (complex128, or float64 with real_vectors != 0); d must be zero on entry outside
// the <= 3 time levels per field the operator touches (see pd_delta_kernel).  x and d must not alias.
extern "C" int pd_delta(pd_handle* h, const void* x_dev, void* d_dev, int real_vectors, void* stream) {
  if (!h || !x_dev || !d_dev || x_dev == d_dev || h->slab_count > 1) {
    pd_set_error("pd_delta: invalid argument (distinct device vectors, unsharded handle)");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_delta_launch(h, (const cplx*)x_dev, (cplx*)d_dev, (cudaStream_t)stream, nullptr, nullptr, real_vectors);
}
