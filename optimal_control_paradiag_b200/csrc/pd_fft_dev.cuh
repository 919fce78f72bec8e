// Device-side building blocks of the time-axis FFT (shared by pd_fft.cu and pd_fused.cu): small DFTs on
// registers and one Stockham pass over the 16 elements a thread owns.  See pd_fft.cu.
#pragma once
#include "pd_common.cuh"

// ------------------------------------------------- power-of-two register kernel
// Small DFTs on registers (forward sign; the inverse transform conjugates on
// load and store).
__device__ __forceinline__ void bf2(cplx& a, cplx& b) {
  cplx t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
template <int R>
__device__ __forceinline__ void dft_pow2(cplx* v);
template <>
__device__ __forceinline__ void dft_pow2<2>(cplx* v) { bf2(v[0], v[1]); }
template <>
__device__ __forceinline__ void dft_pow2<4>(cplx* v) {
  bf2(v[0], v[2]);
  bf2(v[1], v[3]);
  v[3] = cmulni(v[3]);  // * -i
  bf2(v[0], v[1]);
  bf2(v[2], v[3]);
  // outputs in bit-reversed order: v0, v2, v1, v3 -> fix
  cplx t = v[1]; v[1] = v[2]; v[2] = t;
}
template <>
__device__ __forceinline__ void dft_pow2<8>(cplx* v) {
  const double r = 0.70710678118654752440;
  bf2(v[0], v[4]); bf2(v[1], v[5]); bf2(v[2], v[6]); bf2(v[3], v[7]);
  v[5] = cmake((v[5].x + v[5].y) * r, (v[5].y - v[5].x) * r);    // * e^{-i pi/4}
  v[6] = cmulni(v[6]);                                           // * -i
  v[7] = cmake((v[7].y - v[7].x) * r, (-v[7].x - v[7].y) * r);   // * e^{-3i pi/4}
  bf2(v[0], v[2]); bf2(v[1], v[3]); bf2(v[4], v[6]); bf2(v[5], v[7]);
  v[3] = cmulni(v[3]); v[7] = cmulni(v[7]);
  bf2(v[0], v[1]); bf2(v[2], v[3]); bf2(v[4], v[5]); bf2(v[6], v[7]);
  // bit reversal of 3 bits: 1<->4, 3<->6
  cplx t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[3]; v[3] = v[6]; v[6] = t;
}
template <>
__device__ __forceinline__ void dft_pow2<16>(cplx* v) {
  // 4 x 4 decomposition: columns, twiddle W16^{ab}, rows.
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
  const double r = 0.70710678118654752440;
  cplx a[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    cplx t[4] = {v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]};
    dft_pow2<4>(t);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) a[k1][n2] = t[k1];
  }
  // twiddles W16^{k1*n2}
  a[1][1] = cmul(a[1][1], cmake(c1, -s1));
  a[1][2] = cmake((a[1][2].x + a[1][2].y) * r, (a[1][2].y - a[1][2].x) * r);
  a[1][3] = cmul(a[1][3], cmake(s1, -c1));
  a[2][1] = cmake((a[2][1].x + a[2][1].y) * r, (a[2][1].y - a[2][1].x) * r);
  a[2][2] = cmulni(a[2][2]);
  a[2][3] = cmake((a[2][3].y - a[2][3].x) * r, (-a[2][3].x - a[2][3].y) * r);
  a[3][1] = cmul(a[3][1], cmake(s1, -c1));
  a[3][2] = cmake((a[3][2].y - a[3][2].x) * r, (-a[3][2].x - a[3][2].y) * r);
  a[3][3] = cmul(a[3][3], cmake(-c1, s1));  // W16^9 = e^{-9 i pi/8} = (-c1, +s1)
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    cplx t[4] = {a[k1][0], a[k1][1], a[k1][2], a[k1][3]};
    dft_pow2<4>(t);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
  }
}

// padded shared-memory index: one 16-byte pad every 16 elements
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// One Stockham pass of radix R over the 16 elements this thread owns.
//   FIRST : inputs come from global memory (or, with PRE, are already in `io`)
//   LAST  : outputs go to global memory (or, with KEEP, stay in `io`)
//   io    : 16 registers; input order io[u*R + q] <-> element (t + u T) + q N/R,
//           output order io[u*R + r] <-> element base(t + u T) + r Ns
// ---- thread-block-cluster plumbing of the N_t = 16384 kernel (sm_90+ PTX) ----
// Two barrier phases per line, both on the hardware cluster barrier, split into arrive and wait:
//   A "my shared memory may be overwritten": RELAXED arrive behind a block-scope fence -- it orders this
//     CTA's completed shared-memory reads against the peers' later remote stores (arrive.release would
//     stall every warp on a membar until its in-flight global stores are acknowledged device-wide);
//   B "the exchange has landed": release / acquire, the remote stores must be visible to the reader.
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 32-bit shared::cluster address of `smem_ptr` in CTA `rank` of the cluster (keeps the 64-bit generic
// pointers of cluster.map_shared_rank out of the register budget)
__device__ __forceinline__ uint32_t cluster_map(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_store(uint32_t addr, cplx v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <int R, bool INV, bool FIRST, bool LAST, bool PRE = false, bool KEEP = false, bool CLARRIVE = false,
          int GAM = 0>
__device__ __forceinline__ void pow2_pass(const cplx* __restrict__ gsrc, cplx* __restrict__ gdst,
                                          cplx* sm, const cplx* __restrict__ tw, int N, int Ns,
                                          int t, int T, double scale, bool live, cplx* io = nullptr,
                                          const double* __restrict__ gam = nullptr) {
  // GAM / gam (alpha != 1 only; a template flag so that the alpha = 1 kernels carry no trace of it -- a run-time
  // test on the pointer inside the load loop cost the inverse transform 17 %): Gamma_alpha time weights fused into the transform -- the samples are scaled
  // by gam[time index] as the FIRST pass loads them (Gamma before the inverse FFT) or as the LAST pass stores them
  // (Gamma^-1 after the forward FFT): no separate elementwise sweep.  GAM = 1: complex samples, one weight each;
  // GAM = 2: the packed real line of pd_rfft_kernel (element i holds the real samples 2i and 2i + 1).
  constexpr int NB = 16 / R;  // butterflies per thread
  const int NR = N / R;
  const int tws = N / (Ns * R);
  cplx v[NB][R];
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const int j = t + u * T;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      cplx x;
      if (PRE) {
        x = io[u * R + q];
      } else {
        x = FIRST ? gsrc[j + q * NR] : sm[pad16(j + q * NR)];
        if (GAM == 1 && FIRST) x = cscale(x, gam[j + q * NR]);
        if (GAM == 2 && FIRST) {
          x.x *= gam[2 * (j + q * NR)];
          x.y *= gam[2 * (j + q * NR) + 1];
        }
        if (FIRST && INV) x.y = -x.y;
      }
      v[u][q] = x;
    }
  }
  if (!FIRST) {
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int j = t + u * T;
      const int jl = j & (Ns - 1);
      if (R == 2) {
        v[u][1] = cmul(v[u][1], tw[jl * tws]);
      } else {
        // powers of w = W_{Ns R}^{jl}: table look-up for w, w^2, w^3 ..., depth-limited products
        const cplx w1 = tw[jl * tws];
        cplx w[R];
        w[1] = w1;
#pragma unroll
        for (int q = 2; q < R; ++q) w[q] = (q & 1) ? cmul(w[q - 1], w1) : cmul(w[q / 2], w[q / 2]);
#pragma unroll
        for (int q = 1; q < R; ++q) v[u][q] = cmul(v[u][q], w[q]);
      }
    }
    __syncthreads();  // all reads of sm done before anyone overwrites it
    // cluster kernels: this CTA's shared memory is free from here on (split-phase cluster barrier).  The
    // block-scope fence makes sure the shared-memory loads above have been PERFORMED, not merely issued: the
    // untwiddled element v[u][0] is first used after this point, and a peer's remote store (which does not
    // queue behind this SM's own shared-memory pipeline) overtook such a load about once in 5000 lines.
    if (CLARRIVE) {
      __threadfence_block();
      cluster_arrive_relaxed();
    }
  }
#pragma unroll
  for (int u = 0; u < NB; ++u) dft_pow2<R>(v[u]);
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const int j = t + u * T;
    const int base = (j / Ns) * Ns * R + (j & (Ns - 1));
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cplx x = v[u][r];
      if (KEEP) {
        io[u * R + r] = x;
      } else if (LAST) {
        if (INV) x.y = -x.y;
        if (GAM == 2) {
          x.x *= gam[2 * (base + r * Ns)];
          x.y *= gam[2 * (base + r * Ns) + 1];
        }
        if (live) gdst[base + r * Ns] = cscale(x, GAM == 1 ? scale * gam[base + r * Ns] : scale);
      } else {
        sm[pad16(base + r * Ns)] = x;
      }
    }
  }
  if (!LAST) __syncthreads();
}

