// Fused inverse time-FFT + pass A of the partition solve: ONE launch, two CTA roles, pass A fed out of L2.
//
// Replaces scipy.fft.ifft along axis=1 (Control_Wave_PC.py:500-501) followed by the first sweep of the
// per-frequency solves (:445-457 rotation + the local elimination of :460-484, :512) -- see pd_fft.cu and
// pd_solve.cu for the two halves.
//
// Why.  The partition method reads the transformed right-hand side twice (pass A, pass B); as two kernels the
// first of those reads is a full extra HBM sweep (7 sweeps per apply instead of the structural 6).  Running the
// FFT and pass A slab by slab as SEPARATE launches so that pass A hits L2 was measured slower (the machine drains
// and refills ~280 times per apply).  Here both run in the same grid:
//
//   * every CTA draws a ticket (atomicAdd) and derives its role from it, so roles are handed out in a fixed
//     order whatever order the hardware dispatches blocks in:
//         group g = [ FFT CTAs of node-slab g | pass-A CTAs of node-slab g - LAG ]
//     An FFT CTA transforms `lpb` lines exactly as pd_fft_pow2_kernel does, then publishes them
//     (__threadfence + atomicAdd on done[slab]).  A pass-A CTA waits until the node rows of its chunks are
//     published (ld.acquire), then runs the same passA_chunk as the stand-alone kernel, reading through L2 only
//     (ld.global.cg: other CTAs of this very kernel wrote the data, L1 may be stale).
//   * a consumer only ever waits for producers with SMALLER tickets, which are running or done: no deadlock,
//     no co-scheduling assumption.  With LAG = 1 the rows it needs were written ~one slab (a few MB .. tens of
//     MB) ago and are still L2 resident (126 MB): pass A costs no HBM traffic.
//   * both roles keep the resource shape they have as separate kernels (256 threads, <= 128 registers, 68 KiB of
//     shared memory: two CTAs per SM), and the FFT CTAs -- which are HBM-bound -- overlap with the pass-A CTAs,
//     which are fp64- and L2-bound.
#include "pd_fft_dev.cuh"
#include "pd_solve_dev.cuh"

struct FuseParams {
  int per;      // level-0 chunks per slab
  int R;        // node rows per FFT slab = 17 * per
  int nfs;      // FFT slabs   = ceil(n / R)
  int nas;      // pass-A slabs = ceil(nchunks / per)
  int nf;       // FFT CTAs per slab    = ceil(2 R / lpb)
  int na;       // pass-A CTAs per slab = kb * ceil(per / cpb)
  int kb;       // frequency blocks = K / THREADS
  int cpb;      // chunks per pass-A CTA
  int lag;      // pass A trails the FFT by this many slabs (>= 1)
  int n;        // node rows per field
  int nch;      // level-0 chunks (P + 1)
  unsigned* ticket;  // [1]   reset to 0 before every launch
  unsigned* done;    // [nfs] lines published per FFT slab, reset to 0 before every launch
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int R0, int R1, int R2, int R3, bool AL>
__global__ void __launch_bounds__((R0 * R1 * R2 * R3 / 16 <= 256 ? 256 : R0 * R1 * R2 * R3 / 16),
                                  (R0 * R1 * R2 * R3 / 16 <= 256 ? 2 : 1))
pd_fused_ifft_passA_kernel(const cplx* __restrict__ in, cplx* __restrict__ w, const cplx* __restrict__ tw,
                           double scale, cplx* __restrict__ F0, cplx* __restrict__ R1p, cplx* __restrict__ lastl,
                           SolveParams sp, FuseParams fp) {
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  constexpr int THREADS = T < 256 ? 256 : T;
  constexpr int LPB = THREADS / T;
  extern __shared__ __align__(16) unsigned char pd_smem_raw[];
  __shared__ unsigned s_vid;
  const int tid = threadIdx.x;
  if (tid == 0) s_vid = atomicAdd(fp.ticket, 1u);
  __syncthreads();
  const int gsz = fp.nf + fp.na;
  const int g = (int)(s_vid / (unsigned)gsz);
  const int r = (int)(s_vid - (unsigned)g * (unsigned)gsz);

  if (r < fp.nf) {
    // ------------------------------------------------------------ FFT role: `LPB` lines of node-slab g
    if (g >= fp.nfs) return;
    const int row0 = g * fp.R;
    const int rows_s = min(fp.R, fp.n - row0);
    const int lane_line = tid / T;
    const int t = tid - lane_line * T;
    const int l = r * LPB + lane_line;          // line within the slab: [field][row]
    if (r * LPB >= 2 * rows_s) return;          // whole CTA beyond the slab (last, shorter slab)
    const bool live = l < 2 * rows_s;
    const int lc = live ? l : 2 * rows_s - 1;   // clamp: every thread takes the same path through the barriers
    const int field = lc / rows_s;
    const int64_t line = (int64_t)field * fp.n + row0 + (lc - field * rows_s);
    cplx* sm = reinterpret_cast<cplx*>(pd_smem_raw) + (size_t)lane_line * (N + N / 16);
    const cplx* gsrc = in + line * N;
    cplx* gdst = w + line * N;
    constexpr bool L0 = (R1 == 1);
    pow2_pass<R0, true, true, L0>(gsrc, gdst, sm, tw, N, 1, t, T, scale, live);
    if (R1 > 1) {
      constexpr bool L1 = (R2 == 1);
      pow2_pass<(R1 > 1 ? R1 : 2), true, false, L1>(gsrc, gdst, sm, tw, N, R0, t, T, scale, live);
    }
    if (R2 > 1) {
      constexpr bool L2 = (R3 == 1);
      pow2_pass<(R2 > 1 ? R2 : 2), true, false, L2>(gsrc, gdst, sm, tw, N, R0 * R1, t, T, scale, live);
    }
    if (R3 > 1) {
      pow2_pass<(R3 > 1 ? R3 : 2), true, false, true>(gsrc, gdst, sm, tw, N, R0 * R1 * R2, t, T, scale, live);
    }
    // publish: this CTA's lines are complete and visible device-wide before the counter moves
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(fp.done + g, (unsigned)min(LPB, 2 * rows_s - r * LPB));
    return;
  }

  // ---------------------------------------------------------------- pass-A role: chunks of node-slab g - lag
  const int sa = g - fp.lag;
  if (sa < 0 || sa >= fp.nas) return;
  const int a = r - fp.nf;
  const int kblock = a % fp.kb;
  const int cg = a / fp.kb;
  const int cbeg = sa * fp.per + cg * fp.cpb;
  int cend = min(cbeg + fp.cpb, (sa + 1) * fp.per);
  cend = min(cend, fp.nch);
  if (cbeg >= cend) return;
  const int kk = kblock * THREADS + tid;
  const bool valid = kk < sp.kend;
  const int kc_idx = valid ? kk : sp.kend - 1;
  cplx(*mtab)[THREADS] = reinterpret_cast<cplx(*)[THREADS]>(pd_smem_raw);
  const KCoef kc = make_coef<AL>(freq_of(sp, kc_idx), sp);
  fill_pivots<THREADS>(kc, mtab, tid);  // (regenerating the pivots overlaps the wait below)
  if (tid == 0) {
    // rows 17 c + 1 .. 17 c + 17 of the chunks [cbeg, cend), clipped to the last interior row
    const int row_lo = 17 * cbeg + 1;
    const int row_hi = min(17 * (cend - 1) + 17, fp.n - 2);
    for (int s = row_lo / fp.R; s <= row_hi / fp.R; ++s) {
      const unsigned need = 2u * (unsigned)min(fp.R, fp.n - s * fp.R);
      // (bounded: the producers hold smaller tickets and are running or done, so this wait ends within
      // microseconds; if that reasoning were ever violated the kernel must fail loudly, never hang the GPU)
      const long long t0 = clock64();
      while (ld_acquire_gpu(fp.done + s) < need) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) __trap();
      }
    }
  }
  __syncthreads();
  const cplx* wu = w + kc_idx;
  const cplx* wp = w + sp.plane + kc_idx;
  for (int c = cbeg; c < cend; ++c)
    passA_chunk<AL, THREADS, true>(wu, wp, F0, R1p, sp, lastl, kc, mtab, tid, kk, valid, c);
}

// ------------------------------------------------------------------------------------------- host side
void pd_solve_fill_params(pd_handle* h, SolveParams& sp, Levels& lv, SlabPtrs& sl, int half_spectrum);

struct FusePlan {
  unsigned* counters;  // [1 + nfs]: ticket, done[]
  int cap;             // entries allocated
};

void pd_fused_free(pd_handle* h) {
  FusePlan* fpn = reinterpret_cast<FusePlan*>(h->fuse_plan);
  if (!fpn) return;
  if (fpn->counters) cudaFree(fpn->counters);
  delete fpn;
  h->fuse_plan = nullptr;
}

bool pd_fused_supported(const pd_handle* h) {
  const int N = h->cfg.N_t;
  if (h->fft_kind != 1 || N < 1024 || N > 8192) return false;      // K must be a multiple of the CTA width
  if (h->kcount != N || h->nloc != h->n) return false;              // frequency- / line-sharded handles: stage API
  if (pd_solve_nchunks(h) < 2) return false;
  return true;
}

template <int R0, int R1, int R2, int R3>
static int launch_fused(pd_handle* h, const cplx* x, cplx* w, cudaStream_t st, cplx* lastl) {
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int T = N / 16;
  constexpr int THREADS = T < 256 ? 256 : T;
  constexpr int LPB = THREADS / T;
  SolveParams sp; Levels lv; SlabPtrs sl;
  pd_solve_fill_params(h, sp, lv, sl, 0);
  FuseParams fp;
  fp.n = h->n;
  fp.nch = sp.rows[1] + 1;
  fp.per = h->fuse_chunks;
  fp.R = 17 * fp.per;
  fp.nfs = (fp.n + fp.R - 1) / fp.R;
  fp.nas = (fp.nch + fp.per - 1) / fp.per;
  fp.nf = (2 * fp.R + LPB - 1) / LPB;
  fp.kb = (sp.K + THREADS - 1) / THREADS;
  fp.cpb = h->fuse_cpb;
  fp.na = fp.kb * ((fp.per + fp.cpb - 1) / fp.cpb);
  fp.lag = h->fuse_lag;
  FusePlan* fpn = reinterpret_cast<FusePlan*>(h->fuse_plan);
  if (!fpn) {
    fpn = new FusePlan();
    fpn->counters = nullptr;
    fpn->cap = 0;
    h->fuse_plan = fpn;
  }
  if (fpn->cap < 1 + fp.nfs) {
    if (fpn->counters) cudaFree(fpn->counters);
    PD_CUDA(cudaMalloc(&fpn->counters, sizeof(unsigned) * (size_t)(1 + fp.nfs)));
    fpn->cap = 1 + fp.nfs;
  }
  fp.ticket = fpn->counters;
  fp.done = fpn->counters + 1;
  PD_CUDA(cudaMemsetAsync(fpn->counters, 0, sizeof(unsigned) * (size_t)(1 + fp.nfs), st));
  const int ngroups = fp.nfs > fp.nas + fp.lag ? fp.nfs : fp.nas + fp.lag;
  const size_t smem = (size_t)LPB * (N + N / 16) * sizeof(cplx);
  static_assert((size_t)PD_L * THREADS * sizeof(cplx) <= (size_t)LPB * (N + N / 16) * sizeof(cplx),
                "the pivot table of the pass-A role must fit into the FFT role's shared memory");
  const double scale = 1.0 / (double)N;
  const unsigned grid = (unsigned)ngroups * (unsigned)(fp.nf + fp.na);
  if (sp.al) {
    auto k = pd_fused_ifft_passA_kernel<R0, R1, R2, R3, true>;
    PD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, THREADS, smem, st>>>(x, w, h->twiddle, scale, lv.F[0], lv.R[1], lastl, sp, fp);
  } else {
    auto k = pd_fused_ifft_passA_kernel<R0, R1, R2, R3, false>;
    PD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, THREADS, smem, st>>>(x, w, h->twiddle, scale, lv.F[0], lv.R[1], lastl, sp, fp);
  }
  PD_CHECK_LAUNCH();
  h->launches++;
  return PD_OK;
}

// x (2, n, N_t) -> w = ifft_t(x) and the level-0 reduce of w (F[0], R[1] of the solve plan) in one launch
int pd_fused_ifft_passA_launch(pd_handle* h, const cplx* x, cplx* w, cudaStream_t st, cplx* lastl) {
  switch (h->cfg.N_t) {
    case 1024: return launch_fused<16, 16, 4, 1>(h, x, w, st, lastl);
    case 2048: return launch_fused<16, 16, 8, 1>(h, x, w, st, lastl);
    case 4096: return launch_fused<16, 16, 16, 1>(h, x, w, st, lastl);
    case 8192: return launch_fused<16, 16, 8, 4>(h, x, w, st, lastl);
    default: break;
  }
  pd_set_error("pd_fused_ifft_passA: N_t = %d is not covered", h->cfg.N_t);
  return PD_ERR_UNSUPPORTED;
}
