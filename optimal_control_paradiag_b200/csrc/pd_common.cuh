// Shared device/host helpers for libparadiag (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "paradiag.h"

typedef double2 cplx;  // x = re, y = im; 16-byte aligned -> 128-bit loads/stores

// ---------------------------------------------------------------- complex math
__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cneg(cplx a) { return cmake(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// acc + a*b
__host__ __device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
  return cmake(acc.x + a.x * b.x - a.y * b.y, acc.y + a.x * b.y + a.y * b.x);
}
// acc - a*b
__host__ __device__ __forceinline__ cplx cfms(cplx a, cplx b, cplx acc) {
  return cmake(acc.x - a.x * b.x + a.y * b.y, acc.y - a.x * b.y - a.y * b.x);
}
// multiply by +i / -i
__host__ __device__ __forceinline__ cplx cmuli(cplx a) { return cmake(-a.y, a.x); }
__host__ __device__ __forceinline__ cplx cmulni(cplx a) { return cmake(a.y, -a.x); }
__host__ __device__ __forceinline__ cplx crcp(cplx a) {
  double d = 1.0 / (a.x * a.x + a.y * a.y);
  return cmake(a.x * d, -a.y * d);
}

// ------------------------------------------------------------- error handling
void pd_set_error(const char* fmt, ...);

#define PD_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      pd_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                   cudaGetErrorString(_e));                                        \
      return PD_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

#define PD_CHECK_LAUNCH()                                                          \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      pd_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,        \
                   cudaGetErrorString(_e));                                        \
      return PD_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

// Every extern "C" entry point that takes a handle runs on the handle's device and leaves the caller's
// current device as it found it (one process may hold handles on several GPUs).
struct PdDeviceGuard {
  int prev, dev;
  explicit PdDeviceGuard(int device) : prev(-1), dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~PdDeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  PdDeviceGuard(const PdDeviceGuard&) = delete;
  PdDeviceGuard& operator=(const PdDeviceGuard&) = delete;
};
#define PD_ON_DEVICE(h) PdDeviceGuard pd_device_guard_((h)->cfg.device)

// ------------------------------------------------------------------ programmatic dependent launch
// The kernels of one apply depend on each other in a chain, and at the small and the sharded sizes the launch
// latency and the drain / ramp at every kernel boundary are a visible share of the apply.  Kernels launched through
// PD_KLAUNCH carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may be scheduled as soon as every
// CTA of the preceding kernel has executed `griddepcontrol.launch_dependents` (or exited), i.e. while that kernel
// drains.  Every such kernel starts with pd_pdl_enter(): it releases ITS dependents and then blocks in
// `griddepcontrol.wait` until the preceding grid has completed and its writes are visible -- before its first global
// access other than plan-time tables, so read-after-write and write-after-read across kernels stay ordered
// (transitively along the chain: a kernel cannot complete before its own wait has returned).  Launched without the
// attribute (PD_PDL=0, or any plain <<< >>> launch) both instructions are no-ops.
#if defined(__CUDACC__)
__device__ __forceinline__ void pd_pdl_enter(bool release_dependents_now = true) {
  if (release_dependents_now) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... P, typename... A>
static inline cudaError_t pd_klaunch(int pdl, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
// kernel names with template commas go in parentheses: PD_KLAUNCH((k<a, b>), grid, block, smem, stream, args...)
#define PD_KLAUNCH(kernel, grid, block, smem, st, ...) \
  (void)pd_klaunch(h->pdl, kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
#endif

// ------------------------------------------------------------------ the handle
static const int PD_MAX_FFT_PASSES = 16;

struct pd_handle {
  pd_config cfg;
  int n;         // nodes = N_x + 1
  int m;         // interior nodes = N_x - 1
  int kcount;    // frequencies solved by this handle
  int kbegin;
  int nloc;      // node lines per field transformed by this handle
  int slab_rank, slab_count;  // slab mode (x-slab sharding through the solve); count <= 1: off
  int node_begin;             // global index of local node row 0 (0 unless slab mode)
  double dt, h, c;
  int num_sms;
  size_t ws_bytes;
  int64_t launches;

  // FFT plan
  cplx* twiddle;  // e^{-2 pi i j / N_t}, j < N_t
  cplx* twiddle_half;     // same for N_t / 2 (power-of-two N_t >= 128: real-input path)
  cplx* twiddle_quarter;  // same for N_t / 4 (only N_t = 16384: the 4-CTA cluster kernel)
  int fft16k_l2;          // 1: cluster-free 16k kernel (default), 0: 4-CTA cluster kernel (PD_FFT16K=cluster)
  int fft16k_tma;         // 1: bulk-async-copy pipeline kernel (PD_FFT16K=tma)
  int fft16k_clusters;    // co-resident 4-CTA clusters (grid of the persistent 16k kernel)
  double* gamma_tab;      // alpha != 1 only: a^j (N_t entries) followed by a^-j, a = alpha^(1/N_t)
  int fft_kind;   // 0 generic smem Stockham, 1 power-of-two register kernel
  int npass;
  int radix[PD_MAX_FFT_PASSES];

  // solve-stage partition
  int L;       // chunk length
  int P;       // separators (interface unknowns per system)
  int Llast;   // rows in the last chunk
  void* solve_plan;  // SolvePlan of pd_solve.cu: level sizes and interface workspaces
  int iface_thomas_max;  // interface systems up to this many rows use the one-launch sequential kernel

  // work vector (2, n, N_t) for the single-GPU apply
  cplx* work;

  // node-slab interleaved schedule of the single-GPU apply (pd_capi.cu): aux stream + fork/join events
  int sched_chunks;        // level-0 chunks per slab; 0: plain schedule (3 whole-array stages)
  int sched_streams;       // 1 or 2
  cudaStream_t sched_aux;
  cudaEvent_t sched_ev[4];

  // fused inverse FFT + pass A (pd_fused.cu): one launch, pass A fed out of L2
  int fuse_on;             // 1: pd_pc_apply / pd_slab_apply use it
  int fuse_chunks;         // level-0 chunks per node slab
  int fuse_cpb;            // chunks per pass-A CTA
  int fuse_lag;            // pass A trails the FFT by this many slabs
  void* fuse_plan;

  // host staging for the *_host entry points
  cplx* stage_x;
  cplx* stage_y;
  void* pinned;
  size_t pinned_bytes;

  // Krylov workspace (lazily allocated)
  cplx* kry_V;      // basis vectors, allocated in blocks
  int kry_cap;
  int kry_real;     // 1 while a real-vector solve runs (reductions zero the pair-wise imaginary parts)
  cplx* kry_w;
  cplx* kry_t;
  cplx* kry_partial;
  cplx* kry_d;      // (A - P) v of the residual-correction mode: zero except on <= 3 time levels per field
  int kry_d_mode;   // 0: not initialised, 1: complex layout, 2: float64 layout
  int opt_gmres_correction;  // pd_set_option "gmres_residual_correction"
  int opt_slab_no_overlap;   // pd_set_option "slab_overlap" 0: the slab apply stays on the caller's stream
  int opt_kry_real;          // pd_set_option "krylov_real_vectors"
  int pdl;                   // 1: the apply's kernels are launched with programmatic stream serialisation (PD_KLAUNCH)
  int pdl_early;             // which kernels also release their dependents at their first instruction (else implicitly
                             // at exit): 1 time transforms, 2 pass A, 4 interface / functionals, 8 pass B, 16 separators
  int opt_host_register;     // pd_set_option "host_register": page-lock host buffers of pd_pc_apply_host once
  cplx* kry_h;
  double* kry_host;
  cudaStream_t own_stream;
};

// stage launchers implemented in the .cu files
int pd_fft_plan(pd_handle* h);
int pd_fft_launch(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse,
                  cudaStream_t st, int with_gamma = 0);
bool pd_fft_gamma_fused(const pd_handle* h);
bool pd_fft_segments_supported(const pd_handle* h);
int pd_fft_launch_segments(pd_handle* h, const cplx* in, cplx* out, int64_t seg_lines, int nseg, int64_t seg_stride,
                           int inverse, cudaStream_t st);
int pd_solve_nchunks(const pd_handle* h);
int pd_solve_passA_range(pd_handle* h, cplx* w, int c0, int c1, cudaStream_t st);
int pd_solve_interface(pd_handle* h, cudaStream_t st);
int pd_solve_passB_range(pd_handle* h, cplx* w, int c0, int c1, cudaStream_t st);
int pd_gamma_launch(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse, cudaStream_t st);
bool pd_rfft_supported(const pd_handle* h);
// with_gamma (alpha != 1): Gamma on the samples read (to_freq) / Gamma^-1 on the samples written (!to_freq)
int pd_rfft_launch(pd_handle* h, const void* in, void* out, int64_t nlines, int to_freq, cudaStream_t st,
                   int with_gamma = 0);
int pd_rfft_pair_launch(pd_handle* h, const void* in, void* out, int64_t nnodes, int to_freq, cudaStream_t st,
                        int with_gamma = 0);
int pd_solve_plan(pd_handle* h);
int pd_solve_launch(pd_handle* h, cplx* w, cudaStream_t st, cudaEvent_t* ev = nullptr, int half_spectrum = 0);
int pd_slab_reduce_launch(pd_handle* h, cplx* w, cplx* out, cudaStream_t st, int half_spectrum = 0,
                          cudaEvent_t* ev = nullptr, int passA_done = 0, int koff = 0, int kend = 0);
int pd_slab_epoch_bump_launch(pd_handle* h, cudaStream_t st);
cplx* pd_slab_lastl(pd_handle* h);
// fused inverse FFT + pass A (pd_fused.cu)
bool pd_fused_supported(const pd_handle* h);
int pd_fused_ifft_passA_launch(pd_handle* h, const cplx* x, cplx* w, cudaStream_t st, cplx* lastl);
void pd_fused_free(pd_handle* h);
int pd_slab_finish_launch(pd_handle* h, cplx* w, const cplx* gathered, cudaStream_t st, int half_spectrum = 0,
                          cudaEvent_t* ev = nullptr, int koff = 0, int kend = 0, int bump_epoch = 0);
bool pd_slab_half_supported(const pd_handle* h);
int pd_slab_comm_create_impl(pd_handle* h, void* ipc_handle_out, void** base_out);
int pd_slab_comm_connect_impl(pd_handle* h, const void* peers, int mode, const int* peer_devices);
int pd_slab_comm_status_impl(pd_handle* h, int* timed_out, unsigned long long* epoch);
bool pd_slab_comm_ready(const pd_handle* h);
int pd_matvec_launch(pd_handle* h, const cplx* x, cplx* y, cudaStream_t st, int circulant,
                     const cplx* halo_lo = nullptr, const cplx* halo_hi = nullptr, int real_vectors = 0);
int pd_rhs_launch(pd_handle* h, cplx* b, cudaStream_t st, int real_vectors = 0);
int pd_delta_launch(pd_handle* h, const cplx* x, cplx* d, cudaStream_t st, const cplx* halo_lo = nullptr,
                    const cplx* halo_hi = nullptr, int real_vectors = 0);
