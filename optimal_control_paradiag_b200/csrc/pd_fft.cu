// Batched time-axis DFT kernels (complex128) for sm_100a.
//
// Replaces scipy.fft.ifft / fft along axis=1 in DiagFFTPC.apply
// (Control_Wave_PC.py:500-501 and :547-548).  Convention (mat_test.ipynb cells
// 5-9): forward = sum_j x_j e^{-2 pi i jk/N}, inverse = (1/N) sum_j x_j e^{+2 pi i jk/N}.
//
// Two implementations behind pd_fft_launch:
//   kind 1  power-of-two N_t: register-resident radix-16/8/4/2 Stockham passes,
//           one shared-memory exchange between passes (padded against bank
//           conflicts), coalesced 128-bit global loads in the first pass and
//           stores in the last.
//   kind 0  any other N_t (the upstream default is 81 = 3^4): mixed-radix
//           Stockham with direct (O(R) per output) passes ping-ponging between
//           two shared-memory buffers; prime factors of any size are accepted.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <atomic>

#include "pd_common.cuh"


namespace cg = cooperative_groups;

// ------------------------------------------------------------ twiddle table
__global__ void pd_twiddle_kernel(cplx* tw, int N) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) {
    double s, c;
    sincospi(-2.0 * (double)j / (double)N, &s, &c);
    tw[j] = cmake(c, s);
  }
}

// ---------------------------------------------------- Gamma_alpha time weights (alpha != 1, an extension)
// The upstream operator has no alpha; with pd_config.alpha != 1 the apply becomes
// Gamma^-1 fft_t [ per-frequency solves with alpha-shifted symbols ] ifft_t Gamma,  Gamma = diag(a^j),
// a = alpha^(1/N_t) (oracle/pc_alpha.py).  gam[0][j] = a^j, gam[1][j] = a^-j.
__global__ void pd_gamma_table_kernel(double* gam, int N, double lna) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) {
    gam[j] = exp(lna * (double)j);
    gam[N + j] = exp(-lna * (double)j);
  }
}
// out[line][j] = in[line][j] * g[j]; one thread per element, lines are contiguous (in == out allowed)
__global__ void __launch_bounds__(256)
pd_gamma_scale_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int64_t total, int N,
                      const double* __restrict__ g) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    out[i] = cscale(in[i], g[(int)(i % N)]);
}

struct PassList {
  int n;
  int r[PD_MAX_FFT_PASSES];
};

// ------------------------------------------------------------ generic kernel
template <bool INV>
__global__ void __launch_bounds__(256)
pd_fft_generic_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int N,
                      int64_t nlines, const cplx* __restrict__ tw, PassList pl, double scale,
                      const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  cplx* buf0 = reinterpret_cast<cplx*>(pd_smem_raw);
  cplx* buf1 = buf0 + N;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const cplx* src_g = in + line * (int64_t)N;
    cplx* dst_g = out + line * (int64_t)N;
    // gam (alpha != 1): Gamma on load for the inverse transform, Gamma^-1 on store for the forward one
    for (int i = tid; i < N; i += nth) buf0[i] = (INV && gam) ? cscale(src_g[i], gam[i]) : src_g[i];
    __syncthreads();
    cplx* s = buf0;
    cplx* d = buf1;
    int Ns = 1;
    for (int p = 0; p < pl.n; ++p) {
      const int R = pl.r[p];
      const int NR = N / R;
      const int tws = N / (Ns * R);
      const bool last = (p == pl.n - 1);
      for (int o = tid; o < N; o += nth) {
        int jl = o % Ns, t = o / Ns;
        int r = t % R, jh = t / R;
        int j = jh * Ns + jl;
        int step = (int)(((int64_t)jl * tws + (int64_t)r * NR) % N);
        int e = 0;
        cplx acc = cmake(0.0, 0.0);
        for (int q = 0; q < R; ++q) {
          cplx w = tw[e];
          if (INV) w.y = -w.y;
          acc = cfma(s[j + q * NR], w, acc);
          e += step;
          if (e >= N) e -= N;
        }
        if (last)
          dst_g[o] = cscale(acc, (!INV && gam) ? scale * gam[o] : scale);
        else
          d[o] = acc;
      }
      __syncthreads();
      cplx* tmp = s; s = d; d = tmp;
      Ns *= R;
    }
  }
}

// ------------------------------------------------------------ generic real-input pair kernel
// The two-for-one real transform of pd_rfft_pair_kernel (further down) for the time-axis lengths the register
// pipelines do not cover -- every non-power-of-two N_t (the upstream default N_t = 81 included) and the powers of two
// below 128: the u-line and the p-line of one node as ONE complex line c = u + i p through the shared-memory
// passes of pd_fft_generic_kernel.  Half spectra hold the frequencies k = 0 .. N_t/2 (integer division) in rows of
// KP = (N_t/2 + 1 rounded up to 8) complex numbers, padding columns zero.  gam: see pd_rfft_pair_kernel.
__device__ __forceinline__ cplx* generic_forward_passes(cplx* s, cplx* d, int N, const cplx* __restrict__ tw,
                                                        const PassList& pl, int tid, int nth) {
  int Ns = 1;
  for (int p = 0; p < pl.n; ++p) {
    const int R = pl.r[p];
    const int NR = N / R;
    const int tws = N / (Ns * R);
    for (int o = tid; o < N; o += nth) {
      const int jl = o % Ns, t = o / Ns;
      const int r = t % R, jh = t / R;
      const int j = jh * Ns + jl;
      const int step = (int)(((int64_t)jl * tws + (int64_t)r * NR) % N);
      int e = 0;
      cplx acc = cmake(0.0, 0.0);
      for (int q = 0; q < R; ++q) {
        acc = cfma(s[j + q * NR], tw[e], acc);
        e += step;
        if (e >= N) e -= N;
      }
      d[o] = acc;
    }
    __syncthreads();
    cplx* tmp = s; s = d; d = tmp;
    Ns *= R;
  }
  return s;
}

template <bool TO_FREQ>
__global__ void __launch_bounds__(256)
pd_rfft_pair_generic_kernel(const void* __restrict__ in_, void* __restrict__ out_, int N, int64_t nnodes,
                            const cplx* __restrict__ tw, PassList pl, const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  cplx* buf0 = reinterpret_cast<cplx*>(pd_smem_raw);
  cplx* buf1 = buf0 + N;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int H = N / 2;
  const int KP = (H + 1 + 7) & ~7;
  for (int64_t node = blockIdx.x; node < nnodes; node += gridDim.x) {
    if (TO_FREQ) {
      const double* xu = reinterpret_cast<const double*>(in_) + node * N;
      const double* xp = xu + nnodes * N;
      cplx* gu = reinterpret_cast<cplx*>(out_) + node * KP;
      cplx* gp = gu + nnodes * KP;
      for (int i = tid; i < N; i += nth) {
        const double g = gam ? gam[i] : 1.0;
        buf0[i] = cmake(xu[i] * g, xp[i] * g);
      }
      __syncthreads();
      const cplx* C = generic_forward_passes(buf0, buf1, N, tw, pl, tid, nth);
      // u-hat[k] = conj(A)/N, p-hat[k] = conj(B)/N with A = (C[k] + conj C[N-k])/2, B = -i (C[k] - conj C[N-k])/2
      const double sc = 0.5 / (double)N;
      for (int k = tid; k < KP; k += nth) {
        if (k <= H) {
          const cplx a = C[k], b = cconj(C[k == 0 ? 0 : N - k]);
          gu[k] = cmake((a.x + b.x) * sc, -(a.y + b.y) * sc);
          gp[k] = cmake((a.y - b.y) * sc, (a.x - b.x) * sc);
        } else {
          gu[k] = cmake(0.0, 0.0);
          gp[k] = cmake(0.0, 0.0);
        }
      }
      __syncthreads();
    } else {
      const cplx* wu = reinterpret_cast<const cplx*>(in_) + node * KP;
      const cplx* wp = wu + nnodes * KP;
      double* yu = reinterpret_cast<double*>(out_) + node * N;
      double* yp = yu + nnodes * N;
      // D[k] = Wu[k] + i Wp[k] (k <= N/2), the Hermitian extension of both above; y_u + i y_p = FFT(D)
      for (int k = tid; k < N; k += nth) {
        const bool lo = k <= H;
        const int kk = lo ? k : N - k;
        cplx a = wu[kk], b = wp[kk];
        if (!lo) { a.y = -a.y; b.y = -b.y; }
        buf0[k] = cmake(a.x - b.y, a.y + b.x);
      }
      __syncthreads();
      const cplx* Y = generic_forward_passes(buf0, buf1, N, tw, pl, tid, nth);
      for (int i = tid; i < N; i += nth) {
        const double g = gam ? gam[i] : 1.0;
        yu[i] = Y[i].x * g;
        yp[i] = Y[i].y * g;
      }
      __syncthreads();
    }
  }
}

#include "pd_fft_dev.cuh"

// N = R0 * R1 * R2 * R3 (unused radices = 1); T = N/16 threads per line,
// LPB lines per block.
template <int R0, int R1, int R2, int R3, bool INV, bool GAM>
__global__ void __launch_bounds__(512)
pd_fft_pow2_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int64_t nlines,
                   const cplx* __restrict__ tw, double scale, int64_t seg_lines, int64_t seg_stride,
                   const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  const int lpb = blockDim.x / T;
  const int lane_line = threadIdx.x / T;
  const int t = threadIdx.x - lane_line * T;
  cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw) + (size_t)lane_line * (N + N / 16);
  for (int64_t line0 = (int64_t)blockIdx.x * lpb; line0 < nlines; line0 += (int64_t)gridDim.x * lpb) {
    const int64_t line = line0 + lane_line;
    // whole block must take the same path through __syncthreads: clamp the line
    int64_t ln = line < nlines ? line : nlines - 1;
    // the lines may come in equal segments `seg_stride` lines apart (the same node rows of both fields)
    if (ln >= seg_lines) ln = (ln / seg_lines) * seg_stride + ln % seg_lines;
    const cplx* gsrc = in + ln * N;
    cplx* gdst = out + ln * N;
    // a partially filled last block still runs every pass (block-wide barriers)
    // on the clamped line and only skips the final store
    const bool live = line < nlines;
    constexpr bool L0 = (R1 == 1);
    // (INV: Gamma on the loads of the first pass; forward: Gamma^-1 on the stores of the last pass)
    const double* g_in = INV ? gam : nullptr;
    const double* g_out = INV ? nullptr : gam;
    pow2_pass<R0, INV, true, L0, false, false, false, GAM && INV>(gsrc, gdst, sm, tw, N, 1, t, T, scale, live, nullptr, g_in);
    if (R1 > 1) {
      constexpr bool L1 = (R2 == 1);
      pow2_pass<(R1 > 1 ? R1 : 2), INV, false, L1, false, false, false, GAM && !INV && L1>(gsrc, gdst, sm, tw, N, R0, t, T, scale, live, nullptr, g_out);
    }
    if (R2 > 1) {
      constexpr bool L2 = (R3 == 1);
      pow2_pass<(R2 > 1 ? R2 : 2), INV, false, L2, false, false, false, GAM && !INV && L2>(gsrc, gdst, sm, tw, N, R0 * R1, t, T, scale, live, nullptr, g_out);
    }
    if (R3 > 1) {
      pow2_pass<(R3 > 1 ? R3 : 2), INV, false, true, false, false, false, GAM && !INV>(gsrc, gdst, sm, tw, N, R0 * R1 * R2, t, T, scale, live, nullptr, g_out);
    }
    __syncthreads();
  }
}

// N_t = 16384: a line is 256 KiB, more than one CTA's shared memory.  Two kernels exist; both leave the
// frequency axis as [k = 0 mod 4 | 1 mod 4 | 2 mod 4 | 3 mod 4] (N = 4 x 4096, Q = N/4, w = W_N):
//   TO_FREQ (time -> frequency, decimation in frequency):
//       y_q[j] = w^{jq} sum_m x[j + Q m] (-i)^{mq},   X[4k' + q] = FFT_Q(y_q)[k']
//   !TO_FREQ (frequency -> time, decimation in time), input in that [q][k'] order:
//       Z_q = FFT_Q(Y_q),   x[n' + Q m] = sum_q (-i)^{mq} w^{n'q} Z_q[n'].
// The per-frequency solves are independent and only need the index map (freq_of in pd_solve.cu), so no
// reordering pass exists.
//
// (1) pd_fft_16k_l2_kernel (default, further down): one CTA per line, the radix-4 stage is an in-place pass
//     over the line in global memory that L2 absorbs (eviction-priority hints).  4.2 TB/s on B200.
// (2) pd_fft_16k_kernel (PD_FFT16K=cluster): a 4-CTA thread-block cluster per line.  Every CTA runs the
//     4096-point register pipeline (256 threads, 68 KiB of shared memory, two CTAs per SM) and the radix-4
//     stage is an all-to-all between the four CTAs done with REMOTE STORES into distributed shared memory
//     (3/4 of a line crosses the SM-to-SM network once; stores are fire-and-forget, no remote-load latency is
//     exposed).  TO_FREQ: CTA c loads x[j + Q m] for its j-quarter and all m straight from global memory,
//     does the 4-point DFTs and twiddles in registers and stores y_q[j] into CTA q's shared memory; CTA q
//     transforms y_q and writes the q-th quarter of the line.  !TO_FREQ: CTA q transforms its quarter and
//     stores Z_q[n'] into the shared memory of the CTA that owns n'; that CTA twiddles, does the 4-point DFTs
//     and writes four contiguous 16 KiB pieces of the time line.  3.1-3.4 TB/s: the barrier coupling of four
//     CTAs that each share their SM with a CTA of another cluster costs more than the L2 round trip of (1)
//     (without the exchange and the barriers the same pipeline streams at 4.9 TB/s; ncu shows no pipe above
//     50 %).  Measured alternatives for (2): a 2-CTA cluster of 8192-point pipelines reading the peer's half
//     through distributed shared memory (one CTA per SM) 2.5 TB/s; st.async + mbarrier instead of the
//     release/acquire barrier, an L2 prefetch of the next line, three CTAs per SM (80 registers) and
//     512-thread CTAs with radix-8 passes each equal or slower.
#define PD_BIGN 16384
template <bool INV, bool TO_FREQ>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(256, 2)
pd_fft_16k_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int64_t nlines,
                  const cplx* __restrict__ tw, const cplx* __restrict__ tw_q, double scale) {
  constexpr int N = PD_BIGN, Q = N / 4, T = Q / 16, J = Q / 4;  // T = 256 threads, J = 1024 per j-quarter
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw);
  const int t = threadIdx.x;
  const int64_t ncl = gridDim.x / 4;
  // barrier phase A of a line: arrive right after the last local pass has read the shared memory (inside
  // pow2_pass), wait just before the exchange stores -- the wait is hidden behind that pass's arithmetic and
  // global stores and the next line's global loads.  Phase B brackets the exchange.
  cluster_arrive_relaxed();  // A of the first line: nothing to protect yet
  for (int64_t line = blockIdx.x / 4; line < nlines; line += ncl) {
    if (TO_FREQ) {
      const cplx* g = in + line * N + c * J + t;
      cplx v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          cplx x = g[T * u + Q * m];
          if (INV) x.y = -x.y;
          v[u][m] = x;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dft_pow2<4>(v[u]);  // v[u][q] = sum_m x[j + Q m] (-i)^{mq}
        const cplx w1 = tw[c * J + T * u + t];
        const cplx w2 = cmul(w1, w1);
        v[u][1] = cmul(v[u][1], w1);
        v[u][2] = cmul(v[u][2], w2);
        v[u][3] = cmul(v[u][3], cmul(w2, w1));
      }
      cluster_wait();  // A: every CTA of the cluster is done with the previous line
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t dst = cluster_map(sm, q);
#pragma unroll
        for (int u = 0; u < 4; ++u) cluster_store(dst + 16u * (uint32_t)pad16(c * J + T * u + t), v[u][q]);
      }
      cluster_arrive_release();
      cluster_wait_acquire();  // B: y_c is complete in this CTA's shared memory
      cplx io[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) io[q] = sm[pad16(t + T * q)];
      __syncthreads();
      cplx* gdst = out + line * N + (int64_t)c * Q;
      pow2_pass<16, false, true, false, true, false>(nullptr, gdst, sm, tw_q, Q, 1, t, T, scale, true, io);
      pow2_pass<16, false, false, false>(nullptr, gdst, sm, tw_q, Q, 16, t, T, scale, true);
      pow2_pass<16, INV, false, true, false, false, true>(nullptr, gdst, sm, tw_q, Q, 256, t, T, scale, true);
    } else {
      const cplx* gsrc = in + line * N + (int64_t)c * Q;
      cplx io[16];
      pow2_pass<16, INV, true, false>(gsrc, nullptr, sm, tw_q, Q, 1, t, T, scale, true);
      pow2_pass<16, false, false, false>(gsrc, nullptr, sm, tw_q, Q, 16, t, T, scale, true);
      pow2_pass<16, false, false, true, false, true, true>(gsrc, nullptr, sm, tw_q, Q, 256, t, T, scale, true, io);
      // io[r] <-> Z_c[n'], n' = t + 256 r (the twiddle w^{n'c} is applied by the CTA that owns n')
      cluster_wait();  // A: every CTA of the cluster is done with its local passes
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const uint32_t dst = cluster_map(sm, o);
#pragma unroll
        for (int u = 0; u < 4; ++u) cluster_store(dst + 16u * (uint32_t)pad16(c * J + t + T * u), io[4 * o + u]);
      }
      cluster_arrive_release();
      cluster_wait_acquire();  // B: slot [q][s] of this CTA = Z_q[n'], n' = 1024 c + s
      cplx* gdst = out + line * N + c * J + t;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cplx v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = sm[pad16(q * J + T * u + t)];
        const cplx w1 = tw[c * J + T * u + t];  // W_N^{n'}
        const cplx w2 = cmul(w1, w1);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], cmul(w2, w1));
        dft_pow2<4>(v);  // v[m] = x[n' + Q m]
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          cplx y = v[m];
          if (INV) y.y = -y.y;
          gdst[T * u + Q * m] = cscale(y, scale);
        }
      }
      __syncthreads();  // the combine reads are done before the next line's first pass overwrites them
    }
  }
  // balance the last A arrive; after it no peer stores into this CTA any more (its B wait of the last line
  // has seen every store), so the CTA may leave
  cluster_wait();
}

// L2 eviction-priority hints (createpolicy / .L2::cache_hint): the cluster-free 16k kernel parks a 256 KiB
// intermediate line in L2 for a few microseconds; "evict_last" on it and "evict_first" on the streaming
// input keep the intermediates of all resident CTAs (148 x 2 x 256 KiB = 76 MB of the 126 MB L2) from being
// written back to HBM and fetched again.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// L1-bypassing load / store of one complex number with an L2 cache policy
__device__ __forceinline__ cplx ld_cg_hint(const cplx* p, uint64_t pol) {
  cplx v;
  asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_hint(cplx* p, cplx v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// N_t = 16384 WITHOUT a cluster (the default; PD_FFT16K=cluster selects the kernel above): one 256-thread CTA
// owns a whole line and does the radix-4 stage as an in-place pass over the line in GLOBAL memory -- the
// 256 KiB line it has just written (or is about to re-read) sits in L2, so HBM still sees one read and one
// write of the line while no CTA ever waits for another one.  Same [k mod 4] frequency order.
//   TO_FREQ : butterflies (reads `in`, writes y_q[j] to out[q Q + j]), then four 4096-point pipelines in place;
//   !TO_FREQ: four 4096-point pipelines (in -> out), then the butterflies in place on out.
// Every global access is explicit here (L1-bypassing, with an L2 policy): the local pipelines run with their
// first pass fed from registers and their last pass kept in registers.  Measured (B200, ncu): without the
// policies 1.57 + 1.70 GB of DRAM traffic per 1.07 GB sweep and 3.7-3.9 TB/s; with them 1.08 + 1.06 GB and
// 4.2 TB/s (an L2 prefetch of the next input piece and a persistent grid changed nothing).
template <bool INV, bool TO_FREQ, bool GAM>
__global__ void __launch_bounds__(256, 2)
pd_fft_16k_l2_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int64_t nlines,
                     const cplx* __restrict__ tw, const cplx* __restrict__ tw_q, double scale,
                     const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  constexpr int N = PD_BIGN, Q = N / 4, T = Q / 16, J = Q / 4;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw);
  const int t = threadIdx.x;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  auto ld_in = [&](const cplx* p) { return ld_cg_hint(p, pol_stream); };     // streaming input
  auto ld_mid = [&](const cplx* p) { return ld_cg_hint(p, pol_keep); };      // parked intermediate line
  auto st_mid = [&](cplx* p, cplx v) { st_hint(p, v, pol_keep); };
  auto st_out = [&](cplx* p, cplx v) { st_hint(p, v, pol_stream); };         // streaming output
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const cplx* src = in + line * N;
    cplx* dst = out + line * N;
    if (TO_FREQ) {
#pragma unroll 1
      for (int jq = 0; jq < 4; ++jq) {
        cplx v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            cplx x = ld_in(src + jq * J + T * u + t + Q * m);
            if (GAM) x = cscale(x, gam[jq * J + T * u + t + Q * m]);  // Gamma on load (alpha != 1)
            if (INV) x.y = -x.y;
            v[u][m] = x;
          }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dft_pow2<4>(v[u]);
          const cplx w1 = tw[jq * J + T * u + t];
          const cplx w2 = cmul(w1, w1);
          v[u][1] = cmul(v[u][1], w1);
          v[u][2] = cmul(v[u][2], w2);
          v[u][3] = cmul(v[u][3], cmul(w2, w1));
#pragma unroll
          for (int q = 0; q < 4; ++q) st_mid(dst + q * Q + jq * J + T * u + t, v[u][q]);
        }
      }
      __syncthreads();  // the line of y is visible to the whole CTA (read back through L2 below)
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        cplx io[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) io[r] = ld_mid(dst + q * Q + t + T * r);
        pow2_pass<16, false, true, false, true, false>(nullptr, nullptr, sm, tw_q, Q, 1, t, T, scale, true, io);
        pow2_pass<16, false, false, false>(nullptr, nullptr, sm, tw_q, Q, 16, t, T, scale, true);
        pow2_pass<16, false, false, true, false, true>(nullptr, nullptr, sm, tw_q, Q, 256, t, T, scale, true, io);
        // io[r] <-> X[4 (t + 256 r) + q]
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          cplx y = io[r];
          if (INV) y.y = -y.y;
          st_out(dst + q * Q + t + T * r, cscale(y, scale));
        }
        __syncthreads();  // the last pass's shared-memory reads are done before the next quarter's first pass
      }
    } else {
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        cplx io[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          cplx x = ld_in(src + q * Q + t + T * r);
          if (INV) x.y = -x.y;
          io[r] = x;
        }
        pow2_pass<16, false, true, false, true, false>(nullptr, nullptr, sm, tw_q, Q, 1, t, T, 1.0, true, io);
        pow2_pass<16, false, false, false>(nullptr, nullptr, sm, tw_q, Q, 16, t, T, 1.0, true);
        pow2_pass<16, false, false, true, false, true>(nullptr, nullptr, sm, tw_q, Q, 256, t, T, 1.0, true, io);
#pragma unroll
        for (int r = 0; r < 16; ++r) st_mid(dst + q * Q + t + T * r, io[r]);   // Z_q[t + 256 r]
        __syncthreads();
      }
      // (the __syncthreads above also makes the four quarter transforms visible to the whole CTA)
#pragma unroll 1
      for (int jq = 0; jq < 4; ++jq) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = jq * J + T * u + t;
          cplx v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = ld_mid(dst + q * Q + n);
          const cplx w1 = tw[n];
          const cplx w2 = cmul(w1, w1);
          v[1] = cmul(v[1], w1);
          v[2] = cmul(v[2], w2);
          v[3] = cmul(v[3], cmul(w2, w1));
          dft_pow2<4>(v);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            cplx y = v[m];
            if (INV) y.y = -y.y;
            st_out(dst + m * Q + n, cscale(y, GAM ? scale * gam[m * Q + n] : scale));  // Gamma^-1 on store
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// (3) pd_fft_16k_tma_kernel (PD_FFT16K=tma): the same 4 x 4096 decomposition with ALL global traffic on the bulk
//     asynchronous copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) instead of the load/store units.
// ncu on (1): no pipe is saturated (DRAM 48 %, L1TEX 57-61 %, fp64 39 %, 16 of 64 warps resident) -- the kernel
// is latency-bound: a CTA alternates between waiting for its 128-bit loads and computing, and two CTAs per SM
// are all that registers and shared memory allow.  Here ONE persistent 256-thread CTA per SM runs a software
// pipeline of "stages" over two 64 KiB shared-memory buffers:
//      bulk load of stage s+1 / s+2   ||   compute of stage s (in place in its buffer)   ||   bulk store of stage s-1
//   BFLY stage (one j-quarter): 4 strided 16 KiB pieces in, radix-4 butterflies + twiddles on registers, the 4 pieces
//        of the result written back to the same shared-memory slots, 4 pieces out;
//   FFTQ stage (one quarter)  : 64 KiB contiguous in, the 4096-point register pipeline (passes exchange through a third,
//        padded buffer), result written back in place, 64 KiB contiguous out.
// TO_FREQ lines are BFLY x4 (in -> mid, parked in L2 inside `out`) then FFTQ x4 (mid -> out, in place); !TO_FREQ lines
// are FFTQ x4 then BFLY x4.  The second kind of a line may only be loaded once the stores of its first kind have
// completed (cp.async.bulk.wait_group), so the stages of consecutive lines are interleaved (first kind of line i
// between the second-kind stages of line i-1): the dependency is then always at least two stages old and costs no
// bubble.  Loads and stores carry L2 eviction-priority hints like (1).  Every wait is bounded (trap, never a hang).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(b, parity))
    if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#define PD_TMA_BUF (PD_BIGN / 4 * 16)                                    // 64 KiB: one stage
#define PD_TMA_X ((PD_BIGN / 4 + PD_BIGN / 64) * 16)                     // padded exchange buffer of the 4096-point passes
#define PD_TMA_SMEM (2 * PD_TMA_BUF + PD_TMA_X + 64)

template <bool INV, bool TO_FREQ>
__global__ void __launch_bounds__(256, 1)
pd_fft_16k_tma_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int64_t nlines,
                      const cplx* __restrict__ tw, const cplx* __restrict__ tw_q, double scale) {
  constexpr int N = PD_BIGN, Q = N / 4, T = Q / 16, J = Q / 4;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  cplx* bufs[2] = {reinterpret_cast<cplx*>(pd_smem_raw), reinterpret_cast<cplx*>(pd_smem_raw + PD_TMA_BUF)};
  cplx* X = reinterpret_cast<cplx*>(pd_smem_raw + 2 * PD_TMA_BUF);
  uint64_t* full = reinterpret_cast<uint64_t*>(pd_smem_raw + 2 * PD_TMA_BUF + PD_TMA_X);
  const int t = threadIdx.x;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  // this CTA's lines: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t cnt = (nlines - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (cnt <= 0) return;
  const int64_t total = 8 * cnt;
  if (t == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // stage s -> (line index i, second kind?, j)
  auto decode = [&](int64_t s, int64_t& i, bool& second, int& j) {
    if (s < 4) { i = 0; second = false; j = (int)s; return; }
    if (s >= total - 4) { i = cnt - 1; second = true; j = (int)(s - (total - 4)); return; }
    const int64_t r = s - 4, pair = r >> 1;
    j = (int)(pair & 3);
    second = (r & 1) != 0;
    i = (pair >> 2) + (second ? 0 : 1);
  };
  // index of the last first-kind stage of line i (the stores a second-kind stage of line i depends on)
  auto last_first = [&](int64_t i) -> int64_t { return i == 0 ? 3 : 4 + 2 * (4 * (i - 1) + 3); };
  // the stage after whose compute the load of stage s is issued (-1: in the prologue)
  auto issue_at = [&](int64_t s) -> int64_t {
    int64_t i; bool second; int j;
    decode(s, i, second, j);
    const int64_t dep = second ? last_first(i) : -1;
    return (s - 2 > dep) ? s - 2 : dep;
  };
  // BFLY pieces are strided by Q in global memory and packed [4][J] in the buffer; FFTQ is one contiguous quarter
  const bool first_is_bfly = TO_FREQ;
  auto issue_load = [&](int64_t s) {  // thread 0 only
    int64_t i; bool second; int j;
    decode(s, i, second, j);
    const int64_t line = blockIdx.x + i * (int64_t)gridDim.x;
    cplx* b = bufs[s & 1];
    uint64_t* bar = &full[s & 1];
    const bool bfly = (second != first_is_bfly);
    // the first kind reads the caller's input (streaming), the second kind reads the parked intermediate in `out`
    const cplx* src = (second ? out : in) + line * N;
    const uint64_t pol = second ? pol_keep : pol_stream;
    mbar_expect_tx(bar, PD_TMA_BUF);
    if (bfly) {
#pragma unroll
      for (int m = 0; m < 4; ++m) bulk_g2s(b + m * J, src + (int64_t)m * Q + j * J, J * 16, bar, pol);
    } else {
      bulk_g2s(b, src + (int64_t)j * Q, Q * 16, bar, pol);
    }
  };
  auto issue_store = [&](int64_t s) {  // thread 0 only
    int64_t i; bool second; int j;
    decode(s, i, second, j);
    const int64_t line = blockIdx.x + i * (int64_t)gridDim.x;
    cplx* b = bufs[s & 1];
    const bool bfly = (second != first_is_bfly);
    cplx* dst = out + line * N;
    const uint64_t pol = second ? pol_stream : pol_keep;  // first kind writes the parked intermediate
    if (bfly) {
#pragma unroll
      for (int m = 0; m < 4; ++m) bulk_s2g(dst + (int64_t)m * Q + j * J, b + m * J, J * 16, pol);
    } else {
      bulk_s2g(dst + (int64_t)j * Q, b, Q * 16, pol);
    }
    bulk_commit();
  };

  if (t == 0) {
    issue_load(0);
    if (total > 1 && issue_at(1) < 0) issue_load(1);
  }
  for (int64_t s = 0; s < total; ++s) {
    int64_t li; bool second; int j;
    decode(s, li, second, j);
    cplx* b = bufs[s & 1];
    mbar_wait(&full[s & 1], (uint32_t)((s >> 1) & 1));
    const bool bfly = (second != first_is_bfly);
    if (bfly) {
      if (TO_FREQ) {
        // y_q[jJ + n] = w^{(jJ+n) q} sum_m x[jJ + n + Q m] (-i)^{mq}, in place: slot [m][n] -> slot [q][n]
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cplx v[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            cplx x = b[m * J + T * u + t];
            if (INV) x.y = -x.y;
            v[m] = x;
          }
          dft_pow2<4>(v);
          const cplx w1 = tw[j * J + T * u + t];
          const cplx w2 = cmul(w1, w1);
          v[1] = cmul(v[1], w1);
          v[2] = cmul(v[2], w2);
          v[3] = cmul(v[3], cmul(w2, w1));
#pragma unroll
          for (int q = 0; q < 4; ++q) b[q * J + T * u + t] = v[q];
        }
      } else {
        // x[n' + Q m] = sum_q (-i)^{mq} w^{n'q} Z_q[n'], n' = jJ + n, in place: slot [q][n] -> slot [m][n]
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = j * J + T * u + t;
          cplx v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = b[q * J + T * u + t];
          const cplx w1 = tw[n];
          const cplx w2 = cmul(w1, w1);
          v[1] = cmul(v[1], w1);
          v[2] = cmul(v[2], w2);
          v[3] = cmul(v[3], cmul(w2, w1));
          dft_pow2<4>(v);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            cplx y = v[m];
            if (INV) y.y = -y.y;
            b[m * J + T * u + t] = cscale(y, scale);
          }
        }
      }
    } else {
      cplx io[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        cplx x = b[t + T * r];
        if (!TO_FREQ && INV) x.y = -x.y;
        io[r] = x;
      }
      pow2_pass<16, false, true, false, true, false>(nullptr, nullptr, X, tw_q, Q, 1, t, T, 1.0, true, io);
      pow2_pass<16, false, false, false>(nullptr, nullptr, X, tw_q, Q, 16, t, T, 1.0, true);
      pow2_pass<16, false, false, true, false, true>(nullptr, nullptr, X, tw_q, Q, 256, t, T, 1.0, true, io);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        cplx y = io[r];
        if (TO_FREQ) {
          if (INV) y.y = -y.y;
          y = cscale(y, scale);
        }
        b[t + T * r] = y;
      }
    }
    fence_proxy_async();  // this thread's shared-memory writes become visible to the bulk-copy engine
    __syncthreads();      // ... of every thread; also: all reads of X are done before the next FFTQ stage writes it
    if (t == 0) {
      issue_store(s);
      bulk_wait_read0();  // the engine has read buffer s & 1: it may be refilled
      // loads whose turn it is: stage s+1 (if its dependency delayed it) and stage s+2
      for (int64_t s2 = s + 1; s2 <= s + 2 && s2 < total; ++s2) {
        if (issue_at(s2) != s) continue;
        int64_t i2; bool sec2; int j2;
        decode(s2, i2, sec2, j2);
        if (sec2) {  // its input = the stores of the first kind of the same line: complete, not merely issued
          if (s - last_first(i2) >= 1) bulk_wait1(); else bulk_wait0();
        }
        issue_load(s2);
      }
    }
  }
  if (t == 0) bulk_wait0();  // the last stores have left shared memory and reached global memory
}

// ------------------------------------------------------------ real-input fast path
// The Krylov vectors of this (real) optimal-control problem are real, so their time spectra are
// Hermitian and the frequencies k = 0..N_t/2 suffice: half the bytes in every stage.  A real line of
// N = 2M samples is read as M complex numbers z[n] = x[2n] + i x[2n+1]; with Z = FFT_M(z) and
//     G(A)[k] = (A[k] + conj A[M-k])/2 - (i/2) W_N^k (A[k] - conj A[M-k])           (indices mod M)
//   TO_FREQ  : x-hat[k] = conj(G(Z)[k]) / N for k = 0..M          (= scipy ifft(x)[k], :500-501)
//   !TO_FREQ : y = fft of the Hermitian extension of Y[0..M]: F = FFT_M(G(Y)[0..M-1]),
//              y[2n] + i y[2n+1] = 2 conj(F[n])                    (= scipy fft, :547-548, real part)
// Both directions reuse the M-point register pipeline above plus one extra shared-memory exchange.
// GAM (alpha != 1, an extension): gam[j] multiplies the real sample j as it is loaded (TO_FREQ: Gamma) or stored
// (!TO_FREQ: Gamma^-1), see pow2_pass.
template <int R0, int R1, int R2, int R3, bool TO_FREQ, bool GAM>
__global__ void __launch_bounds__(512)
pd_rfft_kernel(const void* __restrict__ in_, void* __restrict__ out_, int64_t nlines,
               const cplx* __restrict__ twN, const cplx* __restrict__ twM, const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  constexpr int G2 = GAM ? 2 : 0;
  constexpr int M = R0 * R1 * R2 * R3;  // complex length = N_t / 2
  constexpr int T = M / 16;
  constexpr int KP = (M + 1 + 7) & ~7;   // row stride of a half spectrum: M + 1 rounded up to 128 bytes
  constexpr int NL = (R3 > 1) ? R3 : (R2 > 1 ? R2 : R1);  // radix of the last pass
  constexpr int NsL = M / NL;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  const int lpb = blockDim.x / T;
  const int lane_line = threadIdx.x / T;
  const int t = threadIdx.x - lane_line * T;
  cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw) + (size_t)lane_line * (M + M / 16 + 16);
  const double invN = 1.0 / (2.0 * M);
  for (int64_t line0 = (int64_t)blockIdx.x * lpb; line0 < nlines; line0 += (int64_t)gridDim.x * lpb) {
    const int64_t line = line0 + lane_line;
    const int64_t ln = line < nlines ? line : nlines - 1;
    const bool live = line < nlines;
    cplx io[16];
    if (TO_FREQ) {
      const cplx* gsrc = reinterpret_cast<const cplx*>(in_) + ln * M;       // N doubles = M complex
      cplx* gdst = reinterpret_cast<cplx*>(out_) + ln * KP;
      // M-point transform, last pass kept in registers
      if (R1 == 1) {
        pow2_pass<R0, false, true, true, false, true, false, G2>(gsrc, gdst, sm, twM, M, 1, t, T, 1.0, live, io, gam);
      } else {
        pow2_pass<R0, false, true, false, false, false, false, G2>(gsrc, gdst, sm, twM, M, 1, t, T, 1.0, live, nullptr, gam);
        if (R2 == 1) {
          pow2_pass<(R1 > 1 ? R1 : 2), false, false, true, false, true>(gsrc, gdst, sm, twM, M, R0, t, T, 1.0, live, io);
        } else {
          pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(gsrc, gdst, sm, twM, M, R0, t, T, 1.0, live);
          if (R3 == 1) {
            pow2_pass<(R2 > 1 ? R2 : 2), false, false, true, false, true>(gsrc, gdst, sm, twM, M, R0 * R1, t, T, 1.0, live, io);
          } else {
            pow2_pass<(R2 > 1 ? R2 : 2), false, false, false>(gsrc, gdst, sm, twM, M, R0 * R1, t, T, 1.0, live);
            pow2_pass<(R3 > 1 ? R3 : 2), false, false, true, false, true>(gsrc, gdst, sm, twM, M, R0 * R1 * R2, t, T, 1.0, live, io);
          }
        }
      }
      __syncthreads();
      // io[u*NL + r] <-> Z[(t + u T) + r NsL]
#pragma unroll
      for (int u = 0; u < 16 / NL; ++u)
#pragma unroll
        for (int r = 0; r < NL; ++r) sm[pad16(t + u * T + r * NsL)] = io[u * NL + r];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int k = t + T * q;
        const cplx a = sm[pad16(k)], b = cconj(sm[pad16((M - k) & (M - 1))]);
        const cplx s2 = cadd(a, b), d2 = csub(a, b);
        const cplx wd = cmul(twN[k], d2);
        // G = s2/2 - (i/2) wd ; output conj(G)/N
        const cplx G = cmake(0.5 * (s2.x + wd.y), 0.5 * (s2.y - wd.x));
        if (live) gdst[k] = cmake(G.x * invN, -G.y * invN);
        if (k == 0 && live) {
          gdst[M] = cmake((a.x - a.y) * invN, 0.0);   // k = M: Re Z0 - Im Z0
          for (int kk = M + 1; kk < KP; ++kk) gdst[kk] = cmake(0.0, 0.0);  // padding columns
        }
      }
      __syncthreads();
    } else {
      const cplx* gsrc = reinterpret_cast<const cplx*>(in_) + ln * KP;
      cplx* gdst = reinterpret_cast<cplx*>(out_) + ln * M;
      // stage the half spectrum Y[0..M] in shared memory
      for (int k = t; k <= M; k += T) sm[pad16(k)] = gsrc[k];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int k = t + T * q;                                           // first-pass input order
        const cplx a = sm[pad16(k)], b = cconj(sm[pad16(M - k)]);
        const cplx s2 = cadd(a, b), d2 = csub(a, b);
        const cplx wd = cmul(twN[k], d2);
        io[q] = cmake(0.5 * (s2.x + wd.y), 0.5 * (s2.y - wd.x));
      }
      __syncthreads();
      // F = FFT_M(io); stored as 2 conj(F): the real samples y[2n], y[2n+1]
      if (R1 == 1) {
        pow2_pass<R0, true, true, true, true, false, false, G2>(gsrc, gdst, sm, twM, M, 1, t, T, 2.0, live, io, gam);
      } else {
        pow2_pass<R0, false, true, false, true, false>(gsrc, gdst, sm, twM, M, 1, t, T, 2.0, live, io);
        if (R2 == 1) {
          pow2_pass<(R1 > 1 ? R1 : 2), true, false, true, false, false, false, G2>(gsrc, gdst, sm, twM, M, R0, t, T, 2.0, live, nullptr, gam);
        } else {
          pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(gsrc, gdst, sm, twM, M, R0, t, T, 2.0, live);
          if (R3 == 1) {
            pow2_pass<(R2 > 1 ? R2 : 2), true, false, true, false, false, false, G2>(gsrc, gdst, sm, twM, M, R0 * R1, t, T, 2.0, live, nullptr, gam);
          } else {
            pow2_pass<(R2 > 1 ? R2 : 2), false, false, false>(gsrc, gdst, sm, twM, M, R0 * R1, t, T, 2.0, live);
            pow2_pass<(R3 > 1 ? R3 : 2), true, false, true, false, false, false, G2>(gsrc, gdst, sm, twM, M, R0 * R1 * R2, t, T, 2.0, live, nullptr, gam);
          }
        }
      }
      __syncthreads();
    }
  }
}

// --------------------------------------------------------------- host side
// cudaFuncSetAttribute once per kernel instantiation and device instead of on every launch (the launch-bound
// small configurations pay for every host call)
// (the mask is atomic: handles may be used from several host threads; a racing second call of
// cudaFuncSetAttribute with the same value is harmless)
#define PD_SET_SMEM_ONCE(kernel, bytes)                                                                  \
  do {                                                                                                   \
    static std::atomic<unsigned long long> pd_done_mask{0};                                              \
    const int pd_dev = h->cfg.device & 63;                                                               \
    if (!(pd_done_mask.load(std::memory_order_acquire) >> pd_dev & 1ull)) {                              \
      PD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      pd_done_mask.fetch_or(1ull << pd_dev, std::memory_order_release);                                  \
    }                                                                                                    \
  } while (0)

static void factorize(int N, PassList& pl) {
  pl.n = 0;
  int rem = N;
  // (9 before 3: N_t = 81, the upstream default, becomes two radix-9 passes instead of four radix-3 passes -- the
  // generic kernel's cost is its block barriers and index arithmetic per pass, not the 9-term sums)
  const int pref[] = {16, 8, 4, 2, 9, 3, 5, 7};
  for (int f : pref)
    while (rem % f == 0 && rem > 1 && pl.n < PD_MAX_FFT_PASSES - 1) {
      pl.r[pl.n++] = f;
      rem /= f;
    }
  for (int f = 11; rem > 1 && pl.n < PD_MAX_FFT_PASSES - 1; f += 2)
    while (rem % f == 0 && pl.n < PD_MAX_FFT_PASSES - 1) {
      pl.r[pl.n++] = f;
      rem /= f;
    }
  if (rem > 1) pl.r[pl.n++] = rem;
  if (pl.n == 0) pl.r[pl.n++] = 1;
}

static bool is_pow2(int N) { return N > 0 && (N & (N - 1)) == 0; }

// the real-input (half-spectrum) path: register pipelines for the powers of two in [128, 16384], the generic
// shared-memory pair kernel for every other N_t >= 8 (a half-spectrum row, N_t/2 + 1 rounded up to 8 columns, must fit
// the N_t columns the workspaces are sized for)
static bool rfft_register_path(const pd_handle* h) {
  const int N = h->cfg.N_t;
  return is_pow2(N) && N >= 128 && N <= 16384 && h->twiddle_half != nullptr;
}
bool pd_rfft_supported(const pd_handle* h) {
  return rfft_register_path(h) || (h->cfg.N_t >= 8 && h->twiddle != nullptr);
}

int pd_fft_plan(pd_handle* h) {
  const int N = h->cfg.N_t;
  PD_CUDA(cudaMalloc(&h->twiddle, sizeof(cplx) * (size_t)N));
  h->ws_bytes += sizeof(cplx) * (size_t)N;
  pd_twiddle_kernel<<<(N + 255) / 256, 256>>>(h->twiddle, N);
  PD_CHECK_LAUNCH();
  if (is_pow2(N) && N >= 128) {  // N/2 table: real-input fast path
    PD_CUDA(cudaMalloc(&h->twiddle_half, sizeof(cplx) * (size_t)(N / 2)));
    h->ws_bytes += sizeof(cplx) * (size_t)(N / 2);
    pd_twiddle_kernel<<<(N / 2 + 255) / 256, 256>>>(h->twiddle_half, N / 2);
    PD_CHECK_LAUNCH();
  }
  if (N == PD_BIGN) {
    PD_CUDA(cudaMalloc(&h->twiddle_quarter, sizeof(cplx) * (size_t)(N / 4)));
    h->ws_bytes += sizeof(cplx) * (size_t)(N / 4);
    pd_twiddle_kernel<<<(N / 4 + 255) / 256, 256>>>(h->twiddle_quarter, N / 4);
    PD_CHECK_LAUNCH();
    const size_t smem = (size_t)(PD_BIGN / 4 + PD_BIGN / 64) * sizeof(cplx);
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_l2_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_l2_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_l2_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_l2_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_tma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PD_TMA_SMEM));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_16k_tma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PD_TMA_SMEM));
    {
      // PD_FFT16K = l2 (default) | tma | cluster
      const char* env = getenv("PD_FFT16K");
      h->fft16k_l2 = !(env && env[0] == 'c');
      h->fft16k_tma = (env && env[0] == 't') ? 1 : 0;
    }
    // co-resident 4-CTA clusters (GPC boundaries keep this a little below num_sms * 2 / 4: 71 on B200)
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(4 * (unsigned)h->num_sms, 1, 1);
    lc.blockDim = dim3(256, 1, 1);
    lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, pd_fft_16k_kernel<true, true>, &lc) != cudaSuccess) {
      cudaGetLastError();
      ncl = 0;
    }
    h->fft16k_clusters = ncl;
  }
  if (h->cfg.alpha != 1.0) {
    PD_CUDA(cudaMalloc(&h->gamma_tab, sizeof(double) * 2 * (size_t)N));
    h->ws_bytes += sizeof(double) * 2 * (size_t)N;
    pd_gamma_table_kernel<<<(N + 255) / 256, 256>>>(h->gamma_tab, N, log(h->cfg.alpha) / (double)N);
    PD_CHECK_LAUNCH();
  }
  PassList pl;
  factorize(N, pl);
  h->npass = pl.n;
  for (int i = 0; i < pl.n; ++i) h->radix[i] = pl.r[i];
  h->fft_kind = (is_pow2(N) && N >= 64 && N <= 16384) ? 1 : 0;
  if (h->fft_kind == 0) {
    size_t smem = 2 * sizeof(cplx) * (size_t)N;
    if (smem > 227 * 1024) {
      pd_set_error("N_t = %d is not supported by the time-axis FFT (needs %zu bytes of shared memory)", N,
                   smem);
      return PD_ERR_INVALID;
    }
    // the limit is a property of the kernel, shared by every handle of the process: always the maximum, so
    // that a later handle with a smaller N_t cannot lower it under an earlier one
    PD_CUDA(cudaFuncSetAttribute(pd_fft_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    PD_CUDA(cudaFuncSetAttribute(pd_fft_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    PD_CUDA(cudaFuncSetAttribute(pd_rfft_pair_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    PD_CUDA(cudaFuncSetAttribute(pd_rfft_pair_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
  }
  return PD_OK;
}

template <int R0, int R1, int R2, int R3>
static int launch_pow2(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse,
                       cudaStream_t st, int64_t seg_lines = 0, int64_t seg_stride = 0, int with_gamma = 0) {
  if (seg_lines <= 0) seg_lines = nlines;
  // alpha != 1: Gamma (inverse transform) / Gamma^-1 (forward) fused into the first / last pass
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (inverse ? 0 : R0 * R1 * R2 * R3) : nullptr;
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  int threads = T < 256 ? 256 : T;
  int lpb = threads / T;
  size_t smem = (size_t)lpb * (N + N / 16) * sizeof(cplx);
  int64_t nblk = (nlines + lpb - 1) / lpb;
  double scale = inverse ? 1.0 / (double)N : 1.0;
  if (inverse && !gam) {
    auto k = pd_fft_pow2_kernel<R0, R1, R2, R3, true, false>;
    PD_SET_SMEM_ONCE(k, smem);
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nlines, h->twiddle, scale, seg_lines, seg_stride, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  } else if (!gam) {
    auto k = pd_fft_pow2_kernel<R0, R1, R2, R3, false, false>;
    PD_SET_SMEM_ONCE(k, smem);
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nlines, h->twiddle, scale, seg_lines, seg_stride, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  } else if (inverse) {
    auto k = pd_fft_pow2_kernel<R0, R1, R2, R3, true, true>;
    PD_SET_SMEM_ONCE(k, smem);
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nlines, h->twiddle, scale, seg_lines, seg_stride, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  } else {
    auto k = pd_fft_pow2_kernel<R0, R1, R2, R3, false, true>;
    PD_SET_SMEM_ONCE(k, smem);
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nlines, h->twiddle, scale, seg_lines, seg_stride, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  }
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

static int launch_16k(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse, cudaStream_t st,
                      int with_gamma = 0) {
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (inverse ? 0 : PD_BIGN) : nullptr;
  const double scale = inverse ? 1.0 / (double)PD_BIGN : 1.0;
  // inverse (time -> frequency, :500-501) leaves the permuted frequency order, forward (:547-548) consumes it.
  const size_t smem = (size_t)(PD_BIGN / 4 + PD_BIGN / 64) * sizeof(cplx);
  if (h->fft16k_tma && !gam) {
    const unsigned grid = (unsigned)(nlines < (int64_t)h->num_sms ? nlines : (int64_t)h->num_sms);
    if (inverse)
      pd_fft_16k_tma_kernel<true, true><<<grid, 256, PD_TMA_SMEM, st>>>(in, out, nlines, h->twiddle, h->twiddle_quarter,
                                                                        scale);
    else
      pd_fft_16k_tma_kernel<false, false><<<grid, 256, PD_TMA_SMEM, st>>>(in, out, nlines, h->twiddle,
                                                                         h->twiddle_quarter, scale);
    PD_CHECK_LAUNCH();
    h->launches++;
    return PD_OK;
  }
  if (h->fft16k_l2) {
    const unsigned grid = (unsigned)(nlines < (int64_t)h->num_sms * 64 ? nlines : (int64_t)h->num_sms * 64);
    if (inverse && !gam)
      PD_KLAUNCH((pd_fft_16k_l2_kernel<true, true, false>), grid, 256, smem, st, in, out, nlines, h->twiddle, h->twiddle_quarter,
                                                                       scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
    else if (!gam)
      PD_KLAUNCH((pd_fft_16k_l2_kernel<false, false, false>), grid, 256, smem, st, in, out, nlines, h->twiddle,
                                                                        h->twiddle_quarter, scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
    else if (inverse)
      PD_KLAUNCH((pd_fft_16k_l2_kernel<true, true, true>), grid, 256, smem, st, in, out, nlines, h->twiddle, h->twiddle_quarter,
                                                                      scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
    else
      PD_KLAUNCH((pd_fft_16k_l2_kernel<false, false, true>), grid, 256, smem, st, in, out, nlines, h->twiddle,
                                                                       h->twiddle_quarter, scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
    PD_CHECK_LAUNCH();
    h->launches++;
    return PD_OK;
  }
  // 4-CTA clusters, persistent over the lines: as many clusters as can be co-resident
  int64_t ncl = h->fft16k_clusters > 0 ? h->fft16k_clusters : h->num_sms / 2;
  if (ncl > nlines) ncl = nlines;
  if (inverse)
    pd_fft_16k_kernel<true, true><<<(unsigned)(4 * ncl), 256, smem, st>>>(in, out, nlines, h->twiddle,
                                                                         h->twiddle_quarter, scale);
  else
    pd_fft_16k_kernel<false, false><<<(unsigned)(4 * ncl), 256, smem, st>>>(in, out, nlines, h->twiddle,
                                                                           h->twiddle_quarter, scale);
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// Two-for-one variant used by the apply: the u-line and the p-line of one node are transformed together
// as ONE complex N_t-point line c = u + i p through the full-size register pipeline (same bytes per CTA
// and the same efficiency as the complex kernel):
//   TO_FREQ : C = FFT(c);  A = (C[k] + conj C[N-k])/2,  B = -i (C[k] - conj C[N-k])/2,
//             u-hat[k] = conj(A)/N,  p-hat[k] = conj(B)/N          for k = 0..N/2
//   !TO_FREQ: D[k] = Wu[k] + i Wp[k] (k <= N/2), D[k] = conj(Wu[N-k]) + i conj(Wp[N-k]) (k > N/2);
//             y_u + i y_p = FFT(D)
// x / y are (2, n, N_t) float64, the half spectra (2, n, KP) complex; one "line" = one node.
// GAM (alpha != 1, an extension): the samples are scaled by gam[time index] on load (TO_FREQ: Gamma) or on store
// (!TO_FREQ: Gamma^-1).
template <int R0, int R1, int R2, int R3, bool TO_FREQ, bool GAM>
__global__ void __launch_bounds__(512)
pd_rfft_pair_kernel(const void* __restrict__ in_, void* __restrict__ out_, int64_t nnodes,
                    const cplx* __restrict__ tw, const double* __restrict__ gam, int pdl_early) {
  pd_pdl_enter(pdl_early != 0);  // programmatic dependent launch: see pd_common.cuh
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  constexpr int H = N / 2;
  constexpr int KP = (H + 1 + 7) & ~7;
  constexpr int NL = (R3 > 1) ? R3 : (R2 > 1 ? R2 : R1);
  constexpr int NsL = N / NL;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  const int lpb = blockDim.x / T;
  const int lane_line = threadIdx.x / T;
  const int t = threadIdx.x - lane_line * T;
  cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw) + (size_t)lane_line * (N + N / 16);
  const double invN = 1.0 / (double)N;
  for (int64_t n0 = (int64_t)blockIdx.x * lpb; n0 < nnodes; n0 += (int64_t)gridDim.x * lpb) {
    const int64_t node = n0 + lane_line;
    const int64_t nd = node < nnodes ? node : nnodes - 1;
    const bool live = node < nnodes;
    cplx io[16];
    if (TO_FREQ) {
      const double* xu = reinterpret_cast<const double*>(in_) + nd * N;
      const double* xp = xu + nnodes * N;
      cplx* gu = reinterpret_cast<cplx*>(out_) + nd * KP;
      cplx* gp = gu + nnodes * KP;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        io[q] = cmake(xu[t + T * q], xp[t + T * q]);
        if (GAM) io[q] = cscale(io[q], gam[t + T * q]);
      }
      if (R2 == 1) {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live, io);
      } else if (R3 == 1) {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live);
        pow2_pass<(R2 > 1 ? R2 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0 * R1, t, T, 1.0, live, io);
      } else {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live);
        pow2_pass<(R2 > 1 ? R2 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0 * R1, t, T, 1.0, live);
        pow2_pass<(R3 > 1 ? R3 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0 * R1 * R2, t, T, 1.0, live, io);
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 16 / NL; ++u)
#pragma unroll
        for (int r = 0; r < NL; ++r) sm[pad16(t + u * T + r * NsL)] = io[u * NL + r];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int k = t + T * q;
        if (k <= H) {
          const cplx a = sm[pad16(k)], b = cconj(sm[pad16((N - k) & (N - 1))]);
          // conj(A)/N and conj(B)/N with A = (a + b)/2, B = -i (a - b)/2
          const double s = 0.5 * invN;
          if (live) {
            gu[k] = cmake((a.x + b.x) * s, -(a.y + b.y) * s);
            gp[k] = cmake((a.y - b.y) * s, (a.x - b.x) * s);
          }
        }
      }
      if (t == 0 && live)
        for (int kk = H + 1; kk < KP; ++kk) { gu[kk] = cmake(0, 0); gp[kk] = cmake(0, 0); }
      __syncthreads();
    } else {
      const cplx* wu = reinterpret_cast<const cplx*>(in_) + nd * KP;
      const cplx* wp = wu + nnodes * KP;
      double* yu = reinterpret_cast<double*>(out_) + nd * N;
      double* yp = yu + nnodes * N;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int k = t + T * q;
        const bool lo = k <= H;
        const int kk = lo ? k : N - k;
        cplx a = wu[kk], b = wp[kk];
        if (!lo) { a.y = -a.y; b.y = -b.y; }
        io[q] = cmake(a.x - b.y, a.y + b.x);   // a + i b
      }
      if (R2 == 1) {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live, io);
      } else if (R3 == 1) {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live);
        pow2_pass<(R2 > 1 ? R2 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0 * R1, t, T, 1.0, live, io);
      } else {
        pow2_pass<R0, false, true, false, true, false>(nullptr, nullptr, sm, tw, N, 1, t, T, 1.0, live, io);
        pow2_pass<(R1 > 1 ? R1 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0, t, T, 1.0, live);
        pow2_pass<(R2 > 1 ? R2 : 2), false, false, false>(nullptr, nullptr, sm, tw, N, R0 * R1, t, T, 1.0, live);
        pow2_pass<(R3 > 1 ? R3 : 2), false, false, true, false, true>(nullptr, nullptr, sm, tw, N, R0 * R1 * R2, t, T, 1.0, live, io);
      }
      // io[u*NL + r] <-> time index (t + u T) + r NsL: real part -> u line, imaginary part -> p line
      if (live) {
#pragma unroll
        for (int u = 0; u < 16 / NL; ++u)
#pragma unroll
          for (int r = 0; r < NL; ++r) {
            const int i = t + u * T + r * NsL;
            const double g = GAM ? gam[i] : 1.0;
            yu[i] = GAM ? io[u * NL + r].x * g : io[u * NL + r].x;
            yp[i] = GAM ? io[u * NL + r].y * g : io[u * NL + r].y;
          }
      }
      __syncthreads();
    }
  }
}

template <int R0, int R1, int R2, int R3>
static int launch_rfft_pair(pd_handle* h, const void* in, void* out, int64_t nnodes, int to_freq, cudaStream_t st,
                            int with_gamma) {
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  int threads = T < 256 ? 256 : T;
  int lpb = threads / T;
  size_t smem = (size_t)lpb * (N + N / 16) * sizeof(cplx);
  int64_t nblk = (nnodes + lpb - 1) / lpb;
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (to_freq ? 0 : N) : nullptr;
#define PD_RFFT_PAIR_GO(TF, GM)                                            \
  do {                                                                     \
    auto k = pd_rfft_pair_kernel<R0, R1, R2, R3, TF, GM>;                  \
    PD_SET_SMEM_ONCE(k, smem);                                             \
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nnodes, h->twiddle, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0)); \
  } while (0)
  if (to_freq) {
    if (gam) PD_RFFT_PAIR_GO(true, true); else PD_RFFT_PAIR_GO(true, false);
  } else {
    if (gam) PD_RFFT_PAIR_GO(false, true); else PD_RFFT_PAIR_GO(false, false);
  }
#undef PD_RFFT_PAIR_GO
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// both fields of `nnodes` nodes at once: (2, nnodes, N_t) float64 <-> (2, nnodes, KP) complex half spectra
int pd_rfft_pair_launch(pd_handle* h, const void* in, void* out, int64_t nnodes, int to_freq, cudaStream_t st,
                        int with_gamma) {
  if (nnodes <= 0) return PD_OK;
  const int g = with_gamma;
  switch (h->cfg.N_t) {
    case 128:  return launch_rfft_pair<16, 8, 1, 1>(h, in, out, nnodes, to_freq, st, g);
    case 256:  return launch_rfft_pair<16, 16, 1, 1>(h, in, out, nnodes, to_freq, st, g);
    case 512:  return launch_rfft_pair<16, 8, 4, 1>(h, in, out, nnodes, to_freq, st, g);
    case 1024: return launch_rfft_pair<16, 16, 4, 1>(h, in, out, nnodes, to_freq, st, g);
    case 2048: return launch_rfft_pair<16, 16, 8, 1>(h, in, out, nnodes, to_freq, st, g);
    case 4096: return launch_rfft_pair<16, 16, 16, 1>(h, in, out, nnodes, to_freq, st, g);
    case 8192: return launch_rfft_pair<16, 16, 8, 4>(h, in, out, nnodes, to_freq, st, g);
    default: break;
  }
  if (rfft_register_path(h)) return -100;  // N_t = 16384: the caller falls back to the per-line packed kernel
  // every other length: the shared-memory pair kernel
  const int N = h->cfg.N_t;
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (to_freq ? 0 : N) : nullptr;
  PassList pl;
  pl.n = h->npass;
  for (int i = 0; i < pl.n; ++i) pl.r[i] = h->radix[i];
  const size_t smem = 2 * sizeof(cplx) * (size_t)N;
  const int64_t cap = (int64_t)h->num_sms * 8;
  const int64_t nblk = nnodes < cap ? nnodes : cap;
  if (to_freq)
    PD_KLAUNCH((pd_rfft_pair_generic_kernel<true>), (unsigned)nblk, 256, smem, st, in, out, N, nnodes, h->twiddle, pl, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  else
    PD_KLAUNCH((pd_rfft_pair_generic_kernel<false>), (unsigned)nblk, 256, smem, st, in, out, N, nnodes, h->twiddle, pl, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

template <int R0, int R1, int R2, int R3>
static int launch_rfft(pd_handle* h, const void* in, void* out, int64_t nlines, int to_freq, cudaStream_t st,
                       int with_gamma) {
  constexpr int M = R0 * R1 * R2 * R3;
  constexpr int T = M / 16;
  int threads = T < 256 ? 256 : T;
  int lpb = threads / T;
  size_t smem = (size_t)lpb * (M + M / 16 + 16) * sizeof(cplx);
  int64_t nblk = (nlines + lpb - 1) / lpb;
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (to_freq ? 0 : 2 * M) : nullptr;
#define PD_RFFT_GO(TF, GM)                                                                              \
  do {                                                                                                  \
    auto k = pd_rfft_kernel<R0, R1, R2, R3, TF, GM>;                                                    \
    PD_SET_SMEM_ONCE(k, smem);                                                                          \
    PD_KLAUNCH(k, (unsigned)nblk, threads, smem, st, in, out, nlines, h->twiddle, h->twiddle_half, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));        \
  } while (0)
  if (to_freq) {
    if (gam) PD_RFFT_GO(true, true); else PD_RFFT_GO(true, false);
  } else {
    if (gam) PD_RFFT_GO(false, true); else PD_RFFT_GO(false, false);
  }
#undef PD_RFFT_GO
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// real lines of N_t samples <-> half spectra of N_t/2 + 1 complex numbers (power-of-two N_t in [128, 16384])
int pd_rfft_launch(pd_handle* h, const void* in, void* out, int64_t nlines, int to_freq, cudaStream_t st,
                   int with_gamma) {
  if (nlines <= 0) return PD_OK;
  const int g = with_gamma;
  if (!rfft_register_path(h)) {
    pd_set_error("the per-line packed real transform needs a power-of-two N_t in [128, 16384] (got %d); "
                 "pd_stage_rfft_pair covers every N_t >= 8", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  switch (h->cfg.N_t / 2) {
    case 64:   return launch_rfft<16, 4, 1, 1>(h, in, out, nlines, to_freq, st, g);
    case 128:  return launch_rfft<16, 8, 1, 1>(h, in, out, nlines, to_freq, st, g);
    case 256:  return launch_rfft<16, 16, 1, 1>(h, in, out, nlines, to_freq, st, g);
    case 512:  return launch_rfft<16, 8, 4, 1>(h, in, out, nlines, to_freq, st, g);
    case 1024: return launch_rfft<16, 16, 4, 1>(h, in, out, nlines, to_freq, st, g);
    case 2048: return launch_rfft<16, 16, 8, 1>(h, in, out, nlines, to_freq, st, g);
    case 4096: return launch_rfft<16, 16, 16, 1>(h, in, out, nlines, to_freq, st, g);
    case 8192: return launch_rfft<16, 16, 8, 4>(h, in, out, nlines, to_freq, st, g);
    default: break;
  }
  pd_set_error("pd_rfft_launch: unsupported N_t %d", h->cfg.N_t);
  return PD_ERR_UNSUPPORTED;
}

// Gamma (inverse = 0) or Gamma^-1 (inverse = 1) on `nlines` time lines
int pd_gamma_launch(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse, cudaStream_t st) {
  if (nlines <= 0 || !h->gamma_tab) return PD_OK;
  const int N = h->cfg.N_t;
  const int64_t total = nlines * N;
  int64_t nblk = (total + 255) / 256;
  const int64_t cap = (int64_t)h->num_sms * 32;
  if (nblk > cap) nblk = cap;
  pd_gamma_scale_kernel<<<(unsigned)nblk, 256, 0, st>>>(in, out, total, N, h->gamma_tab + (inverse ? N : 0));
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// true when pd_fft_launch_segments exists for this N_t (the register kernels up to 8192 points)
bool pd_fft_segments_supported(const pd_handle* h) {
  return h->fft_kind == 1 && h->cfg.N_t <= 8192;
}

// `nseg` segments of `seg_lines` lines each, `seg_stride` lines apart (in and out alike): the same node rows of
// both fields in one launch
int pd_fft_launch_segments(pd_handle* h, const cplx* in, cplx* out, int64_t seg_lines, int nseg, int64_t seg_stride,
                           int inverse, cudaStream_t st) {
  const int64_t nlines = seg_lines * nseg;
  if (nlines <= 0) return PD_OK;
  switch (h->fft_kind == 1 ? h->cfg.N_t : 0) {
    case 64:   return launch_pow2<16, 4, 1, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 128:  return launch_pow2<16, 8, 1, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 256:  return launch_pow2<16, 16, 1, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 512:  return launch_pow2<16, 8, 4, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 1024: return launch_pow2<16, 16, 4, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 2048: return launch_pow2<16, 16, 8, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 4096: return launch_pow2<16, 16, 16, 1>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    case 8192: return launch_pow2<16, 16, 8, 4>(h, in, out, nlines, inverse, st, seg_lines, seg_stride);
    default: break;
  }
  pd_set_error("pd_fft_launch_segments: N_t = %d has no segmented kernel", h->cfg.N_t);
  return PD_ERR_UNSUPPORTED;
}

// true when pd_fft_launch can apply the Gamma_alpha weights inside the transform (with_gamma)
bool pd_fft_gamma_fused(const pd_handle* h) {
  return !(h->cfg.N_t == PD_BIGN && h->fft_kind == 1 && !h->fft16k_l2);  // every kernel but the 16k cluster variant
}

// with_gamma (alpha != 1): the transform also applies Gamma (inverse != 0: on its input) / Gamma^-1 (inverse == 0:
// on its output) -- the time-weight scaling of the alpha-circulant costs no sweep of its own
int pd_fft_launch(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse,
                  cudaStream_t st, int with_gamma) {
  const int N = h->cfg.N_t;
  if (nlines <= 0) return PD_OK;
  const int g = with_gamma;
  if (h->fft_kind == 1) {
    switch (N) {
      case 64:   return launch_pow2<16, 4, 1, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 128:  return launch_pow2<16, 8, 1, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 256:  return launch_pow2<16, 16, 1, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 512:  return launch_pow2<16, 8, 4, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 1024: return launch_pow2<16, 16, 4, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 2048: return launch_pow2<16, 16, 8, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 4096: return launch_pow2<16, 16, 16, 1>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 8192: return launch_pow2<16, 16, 8, 4>(h, in, out, nlines, inverse, st, 0, 0, g);
      case 16384: return launch_16k(h, in, out, nlines, inverse, st, g);
      default: break;
    }
  }
  const double* gam = (with_gamma && h->gamma_tab) ? h->gamma_tab + (inverse ? 0 : N) : nullptr;
  PassList pl;
  pl.n = h->npass;
  for (int i = 0; i < pl.n; ++i) pl.r[i] = h->radix[i];
  size_t smem = 2 * sizeof(cplx) * (size_t)N;
  int64_t nblk = nlines < (int64_t)h->num_sms * 8 ? nlines : (int64_t)h->num_sms * 8;
  double scale = inverse ? 1.0 / (double)N : 1.0;
  if (inverse)
    PD_KLAUNCH((pd_fft_generic_kernel<true>), (unsigned)nblk, 256, smem, st, in, out, N, nlines, h->twiddle, pl, scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  else
    PD_KLAUNCH((pd_fft_generic_kernel<false>), (unsigned)nblk, 256, smem, st, in, out, N, nlines, h->twiddle, pl, scale, gam, (int)(h->pdl ? (h->pdl_early & 1) : 0));
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}
