// C ABI of libparadiag.so: handle lifetime and the DiagFFTPC entry points.
// See include/paradiag.h for the contract and the upstream lines each call replaces.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "pd_common.cuh"

static thread_local char g_err[512] = "";

void pd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void pd_krylov_free(pd_handle* h);
static void hostreg_release(pd_handle* h);
void pd_solve_free(pd_handle* h);

extern "C" const char* pd_last_error(void) { return g_err; }
extern "C" int pd_abi_version(void) { return PD_ABI_VERSION; }
extern "C" size_t pd_workspace_bytes(const pd_handle* h) { return h ? h->ws_bytes : 0; }
extern "C" int64_t pd_launch_count(const pd_handle* h) { return h ? h->launches : 0; }

extern "C" int pd_destroy(pd_handle* h) {
  if (!h) return PD_OK;
  PD_ON_DEVICE(h);
  pd_krylov_free(h);
  if (h->twiddle) cudaFree(h->twiddle);
  if (h->twiddle_half) cudaFree(h->twiddle_half);
  if (h->twiddle_quarter) cudaFree(h->twiddle_quarter);
  if (h->gamma_tab) cudaFree(h->gamma_tab);
  pd_solve_free(h);
  pd_fused_free(h);

  if (h->work) cudaFree(h->work);
  if (h->stage_x) cudaFree(h->stage_x);
  if (h->stage_y) cudaFree(h->stage_y);
  hostreg_release(h);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->sched_aux) cudaStreamDestroy(h->sched_aux);
  for (cudaEvent_t e : h->sched_ev)
    if (e) cudaEventDestroy(e);
  delete h;
  return PD_OK;
}

extern "C" int pd_create(const pd_config* cfg, pd_handle** out) {
  if (!cfg || !out) {
    pd_set_error("pd_create: null argument");
    return PD_ERR_INVALID;
  }
  *out = nullptr;
  if (cfg->abi_version != PD_ABI_VERSION) {
    pd_set_error("pd_create: ABI version mismatch (caller %d, library %d)", cfg->abi_version, PD_ABI_VERSION);
    return PD_ERR_INVALID;
  }
  if (cfg->N_x < 2 || cfg->N_t < 3) {
    pd_set_error("pd_create: need N_x >= 2 and N_t >= 3 (got %d, %d)", cfg->N_x, cfg->N_t);
    return PD_ERR_INVALID;
  }
  if (!(cfg->T > 0.0) || !(cfg->gamma > 0.0)) {
    pd_set_error("pd_create: T and gamma must be positive");
    return PD_ERR_INVALID;
  }
  // alpha != 1 is an extension (the upstream preconditioner is the alpha = 1 block circulant and has no alpha;
  // definition in oracle/pc_alpha.py): the single-GPU and slab-mode applies (complex and real-input) and GMRES,
  // not the frequency-sharded stage handles of the all-to-all mode
  if (!(cfg->alpha > 0.0) || cfg->alpha > 1.0) {
    pd_set_error("pd_create: alpha = %g outside (0, 1]", cfg->alpha);
    return PD_ERR_INVALID;
  }
  if (cfg->alpha != 1.0 && (cfg->k_count > 0 || cfg->n_local > 0)) {
    pd_set_error("pd_create: alpha = %g on a frequency- / node-sharded stage handle is not supported (alpha != 1: "
                 "single-GPU or slab mode)", cfg->alpha);
    return PD_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    pd_set_error("pd_create: no CUDA device available (%s); libparadiag has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return PD_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    pd_set_error("pd_create: device %d out of range (%d devices)", cfg->device, ndev);
    return PD_ERR_INVALID;
  }
  PdDeviceGuard pd_device_guard_(cfg->device);  // the caller's current device is restored on return
  pd_handle* h = new pd_handle();
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->n = cfg->N_x + 1;
  h->m = cfg->N_x - 1;
  h->kbegin = cfg->k_count > 0 ? cfg->k_begin : 0;
  h->kcount = cfg->k_count > 0 ? cfg->k_count : cfg->N_t;
  h->nloc = cfg->n_local > 0 ? cfg->n_local : h->n;
  h->slab_rank = cfg->slab_rank;
  h->slab_count = cfg->slab_count > 1 ? cfg->slab_count : 1;
  if (h->slab_count > 1) {
    // this rank's node rows: row 0 is the Dirichlet node (rank 0) or the separator this slab owns;
    // the body follows; the last rank ends with the Dirichlet node
    if (cfg->slab_rank < 0 || cfg->slab_rank >= h->slab_count || cfg->k_count > 0) {
      pd_set_error("pd_create: bad slab_rank %d of %d (or k_count set together with slab mode)", cfg->slab_rank,
                   h->slab_count);
      delete h;
      return PD_ERR_INVALID;
    }
    const int ntot = cfg->N_x + 1, G = h->slab_count, r = h->slab_rank;
    const int cnt = ntot / G + (r < ntot % G ? 1 : 0);
    h->n = cnt;
    h->m = cnt - 1 - (r == G - 1 ? 1 : 0);
    h->nloc = cnt;
    h->node_begin = r * (ntot / G) + (r < ntot % G ? r : ntot % G);
  }
  if (h->kbegin < 0 || h->kbegin + h->kcount > cfg->N_t) {
    pd_set_error("pd_create: frequency shard [%d, %d) outside [0, %d)", h->kbegin, h->kbegin + h->kcount,
                 cfg->N_t);
    delete h;
    return PD_ERR_INVALID;
  }
  h->dt = cfg->T / cfg->N_t;
  h->h = 1.0 / cfg->N_x;
  h->c = h->dt * h->dt / sqrt(cfg->gamma);
  cudaDeviceProp prop;
  PD_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  h->num_sms = prop.multiProcessorCount;
  {
    // Programmatic dependent launch of the apply's kernels (PD_KLAUNCH / pd_pdl_enter in pd_common.cuh).  Measured on
    // B200: the launch attribute alone gives cfg5 0.674 -> 0.665 ms, 2048 x 4096 360 -> 350 us, cfg3 unchanged.
    // Letting kernels ALSO release their dependents at their first instruction helps only where the grids are below
    // one wave (cfg1 46.8 -> 45.7 us, cfg2 80.0 -> 76.2 us).  At the large sizes it is neutral in the time transforms
    // and in pass B, and harmful in the interface kernel (cfg3 2.53 -> 2.74 ms: the next kernel's CTAs become resident
    // beside a kernel that runs one warp per scheduler on its own latency chain) and in pass A of a short slab
    // (0.388 -> 0.409 ms per rank at cfg3 / 8).  So: attribute everywhere; early release (a bit mask per kernel group,
    // pd_handle::pdl_early) everywhere for vectors <= 64 MiB, nowhere above.  PD_PDL=0 turns the attribute off,
    // PD_PDL_EARLY=<mask> forces the release mask.
    const char* e = getenv("PD_PDL");
    const char* t = getenv("PD_PDL_EARLY");
    const double vec_bytes = 32.0 * (double)h->n * (double)cfg->N_t;
    h->pdl = !(e && e[0] == '0');
    h->pdl_early = t ? atoi(t) : (vec_bytes <= 64.0 * 1024 * 1024 ? 31 : 0);  // bit mask, see pd_common.cuh
  }
  int rc = pd_fft_plan(h);
  if (rc == PD_OK) rc = pd_solve_plan(h);
  if (rc != PD_OK) {
    pd_destroy(h);
    return rc;
  }
  // Schedule of the single-GPU apply.  PD_SCHED=interleave runs it NODE SLAB BY NODE SLAB as separate launches:
  // inverse FFT of a slab of node rows, then pass A of the level-0 chunks in it while the slab is still in L2;
  // later pass B of a slab, then its forward FFT (see apply_interleaved).  Bit-identical results, but MEASURED
  // SLOWER on B200 at every slab size (cfg3: 3.2 ms with 28 MB slabs on two streams, 2.9 ms with 221 MB slabs,
  // against 2.73 ms plain; the same under a CUDA graph -- the cost is the drain / refill of ~280 short kernels,
  // not the host), so it stays opt-in.  PD_SCHED_CHUNKS = level-0 chunks per slab; PD_SCHED_STREAMS = 1|2.
  {
    const char* mode = getenv("PD_SCHED");
    const bool on = mode && mode[0] == 'i';
    const bool sharded = cfg->k_count > 0 || cfg->n_local > 0 || h->slab_count > 1;
    if (on && !sharded && pd_fft_segments_supported(h) && pd_solve_nchunks(h) > 1) {
      // default slab: ~2 resident waves of FFT lines (444 one-line CTAs fit on 148 SMs at N_t = 4096)
      const char* ce = getenv("PD_SCHED_CHUNKS");
      int per = ce ? atoi(ce) : 0;
      if (per <= 0) {
        const int64_t target_bytes = (int64_t)28 << 20;  // per slab, both fields
        per = (int)(target_bytes / ((int64_t)sizeof(cplx) * 2 * 17 * cfg->N_t));
        if (per < 1) per = 1;
      }
      h->sched_chunks = per;
      const char* se = getenv("PD_SCHED_STREAMS");
      h->sched_streams = se ? atoi(se) : 2;
      if (h->sched_streams > 1) {
        PD_CUDA(cudaStreamCreateWithFlags(&h->sched_aux, cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i) PD_CUDA(cudaEventCreateWithFlags(&h->sched_ev[i], cudaEventDisableTiming));
      }
    }
  }
  // Fused inverse FFT + pass A (pd_fused.cu; power-of-two N_t in [1024, 8192], unsharded or x-slab handles):
  // OPT-IN with PD_FUSE=1.  Bit-identical results and pass A really is fed out of L2, but MEASURED SLOWER on B200:
  // 1.40-1.50 ms for the fused launch against 0.68 + 0.38 ms for the two kernels at cfg3 (tools/fuse_probe.py, 24
  // settings of slab size / chunks per CTA / lag).  Both roles are occupancy-bound at two CTAs per SM: an FFT CTA
  // needs every resident slot to keep HBM busy, so slots given to pass-A CTAs slow the FFT by the same amount
  // (DESIGN.md section 8).  PD_FUSE_CHUNKS / PD_FUSE_CPB / PD_FUSE_LAG tune it.
  {
    const char* e = getenv("PD_FUSE");
    h->fuse_on = pd_fused_supported(h) && e && e[0] == '1';
    const char* c = getenv("PD_FUSE_CHUNKS");
    h->fuse_chunks = c && atoi(c) > 0 ? atoi(c) : 8;
    const char* b = getenv("PD_FUSE_CPB");
    h->fuse_cpb = b && atoi(b) > 0 ? atoi(b) : 2;
    if (h->fuse_cpb > h->fuse_chunks) h->fuse_cpb = h->fuse_chunks;
    const char* l = getenv("PD_FUSE_LAG");
    h->fuse_lag = l && atoi(l) > 0 ? atoi(l) : 1;
  }
  PD_CUDA(cudaDeviceSynchronize());
  *out = h;
  return PD_OK;
}

// The apply as node slabs (alpha = 1, power-of-two N_t <= 8192, vectors >> L2).  A slab = `sched_chunks`
// consecutive level-0 chunks of the partition solve = a contiguous range of node rows of both fields:
//   first half :  for every slab   inverse FFT of its rows (x -> work)  ->  pass A of its chunks
//   interface  :  the small levels (whole x-range, ~6 % of the data)
//   second half:  for every slab   pass B of its chunks (in place)      ->  forward FFT of its rows (work -> y)
// Pass A reads what the FFT has just written and the FFT reads what pass B has just written: those two sweeps
// are served by L2 (126 MB) instead of HBM.  Alternate slabs go to a second stream so that the ramp-down of one
// small kernel overlaps the ramp-up of the next; fork/join by events, capturable in a CUDA graph.
static int apply_interleaved(pd_handle* h, const cplx* x, cplx* y, cudaStream_t st) {
  const int nch = pd_solve_nchunks(h), per = h->sched_chunks, n = h->n;
  const int nsl = (nch + per - 1) / per;
  const bool two = h->sched_streams > 1 && nsl > 1 && h->sched_aux;
  cudaStream_t q[2] = {st, two ? h->sched_aux : st};
  auto rows = [&](int s, int& r0, int& r1, int& c0, int& c1) {
    c0 = s * per;
    c1 = c0 + per < nch ? c0 + per : nch;
    r0 = c0 == 0 ? 0 : c0 * 17 + 1;   // chunk c = rows 17 c + 1 .. 17 c + 16 and the separator 17 c + 17
    r1 = c1 == nch ? n : c1 * 17 + 1;
  };
  int rc;
  if (two) {
    PD_CUDA(cudaEventRecord(h->sched_ev[0], st));
    PD_CUDA(cudaStreamWaitEvent(h->sched_aux, h->sched_ev[0], 0));
  }
  for (int s = 0; s < nsl; ++s) {
    int r0, r1, c0, c1;
    rows(s, r0, r1, c0, c1);
    cudaStream_t qs = q[s & 1];
    const int64_t off = (int64_t)r0 * h->cfg.N_t;
    if ((rc = pd_fft_launch_segments(h, x + off, h->work + off, r1 - r0, 2, n, 1, qs))) return rc;
    if ((rc = pd_solve_passA_range(h, h->work, c0, c1, qs))) return rc;
  }
  if (two) {
    PD_CUDA(cudaEventRecord(h->sched_ev[1], h->sched_aux));
    PD_CUDA(cudaStreamWaitEvent(st, h->sched_ev[1], 0));
  }
  if ((rc = pd_solve_interface(h, st))) return rc;
  if (two) {
    PD_CUDA(cudaEventRecord(h->sched_ev[2], st));
    PD_CUDA(cudaStreamWaitEvent(h->sched_aux, h->sched_ev[2], 0));
  }
  // last slabs first: their rows are the most recently written ones, part of them is still in L2
  for (int s = nsl - 1; s >= 0; --s) {
    int r0, r1, c0, c1;
    rows(s, r0, r1, c0, c1);
    cudaStream_t qs = q[s & 1];
    const int64_t off = (int64_t)r0 * h->cfg.N_t;
    if ((rc = pd_solve_passB_range(h, h->work, c0, c1, qs))) return rc;
    if ((rc = pd_fft_launch_segments(h, h->work + off, y + off, r1 - r0, 2, n, 0, qs))) return rc;
  }
  if (two) {
    PD_CUDA(cudaEventRecord(h->sched_ev[3], h->sched_aux));
    PD_CUDA(cudaStreamWaitEvent(st, h->sched_ev[3], 0));
  }
  return PD_OK;
}

static int ensure_work(pd_handle* h) {
  if (!h->work) {
    size_t bytes = sizeof(cplx) * 2 * (size_t)h->n * h->cfg.N_t;
    cudaError_t e = cudaMalloc(&h->work, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      pd_set_error("workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      return PD_ERR_NOMEM;
    }
    h->ws_bytes += bytes;
  }
  return PD_OK;
}

extern "C" int pd_stage_fft(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines, int inverse,
                            void* stream) {
  if (!h || !in_dev || !out_dev || nlines < 0) {
    pd_set_error("pd_stage_fft: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_fft_launch(h, (const cplx*)in_dev, (cplx*)out_dev, nlines, inverse, (cudaStream_t)stream);
}

extern "C" int pd_stage_gamma(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines, int inverse,
                              void* stream) {
  if (!h || !in_dev || !out_dev || nlines < 0) {
    pd_set_error("pd_stage_gamma: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_gamma_launch(h, (const cplx*)in_dev, (cplx*)out_dev, nlines, inverse, (cudaStream_t)stream);
}

extern "C" int pd_stage_solve(pd_handle* h, void* w_dev, void* stream) {
  if (!h || !w_dev) {
    pd_set_error("pd_stage_solve: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_solve_launch(h, (cplx*)w_dev, (cudaStream_t)stream);
}

extern "C" int pd_slab_reduce(pd_handle* h, void* w_dev, void* out_dev, void* stream) {
  if (!h || !w_dev || !out_dev || h->slab_count <= 1) {
    pd_set_error("pd_slab_reduce: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_slab_reduce_launch(h, (cplx*)w_dev, (cplx*)out_dev, (cudaStream_t)stream);
}

extern "C" int pd_slab_finish(pd_handle* h, void* w_dev, const void* gathered_dev, void* stream) {
  if (!h || !w_dev || !gathered_dev || h->slab_count <= 1) {
    pd_set_error("pd_slab_finish: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_slab_finish_launch(h, (cplx*)w_dev, (const cplx*)gathered_dev, (cudaStream_t)stream);
}

// slab mode on the half spectrum of the real-input path: w = (2, n_r, Kp), out (6, Kp), gathered (G, 6, Kp)
extern "C" int pd_slab_reduce_half(pd_handle* h, void* w_dev, void* out_dev, void* stream) {
  if (!h || !w_dev || !out_dev || h->slab_count <= 1) {
    pd_set_error("pd_slab_reduce_half: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (!pd_slab_half_supported(h)) {
    pd_set_error("pd_slab_reduce_half: needs N_t >= 8 (got %d)", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  return pd_slab_reduce_launch(h, (cplx*)w_dev, (cplx*)out_dev, (cudaStream_t)stream, 1);
}

extern "C" int pd_slab_finish_half(pd_handle* h, void* w_dev, const void* gathered_dev, void* stream) {
  if (!h || !w_dev || !gathered_dev || h->slab_count <= 1) {
    pd_set_error("pd_slab_finish_half: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (!pd_slab_half_supported(h)) {
    pd_set_error("pd_slab_finish_half: needs N_t >= 8 (got %d)", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  return pd_slab_finish_launch(h, (cplx*)w_dev, (const cplx*)gathered_dev, (cudaStream_t)stream, 1);
}

// ---- slab mode with the peer-store exchange: the whole distributed apply behind one call
extern "C" int pd_slab_comm_create(pd_handle* h, void* ipc_handle_out, void** base_out) {
  if (!h || h->slab_count <= 1) {
    pd_set_error("pd_slab_comm_create: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_slab_comm_create_impl(h, ipc_handle_out, base_out);
}

extern "C" int pd_slab_comm_connect(pd_handle* h, const void* peers, int mode, const int* peer_devices) {
  if (!h || !peers || h->slab_count <= 1 || (mode != 0 && mode != 1)) {
    pd_set_error("pd_slab_comm_connect: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_slab_comm_connect_impl(h, peers, mode, peer_devices);
}

extern "C" int pd_slab_comm_status(pd_handle* h, int* timed_out, uint64_t* epoch) {
  if (!h || h->slab_count <= 1) {
    pd_set_error("pd_slab_comm_status: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  unsigned long long ep = 0;
  int rc = pd_slab_comm_status_impl(h, timed_out, &ep);
  if (epoch) *epoch = ep;
  return rc;
}

static int slab_apply_check(pd_handle* h, const void* p, const char* who, int real_input) {
  if (!h || !p || h->slab_count <= 1) {
    pd_set_error("%s: invalid argument or handle not in slab mode", who);
    return PD_ERR_INVALID;
  }
  if (!pd_slab_comm_ready(h)) {
    pd_set_error("%s: the peer-store exchange is not connected (pd_slab_comm_create / pd_slab_comm_connect)", who);
    return PD_ERR_INVALID;
  }
  if (real_input && !pd_slab_half_supported(h)) {
    pd_set_error("%s: the real-input path needs N_t >= 8 (got %d)", who, h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  return PD_OK;
}

// aux stream + fork/join events (shared with the opt-in interleaved schedule), created on first use
static int ensure_aux(pd_handle* h) {
  if (!h->sched_aux) {
    PD_CUDA(cudaStreamCreateWithFlags(&h->sched_aux, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i)
      if (!h->sched_ev[i]) PD_CUDA(cudaEventCreateWithFlags(&h->sched_ev[i], cudaEventDisableTiming));
  }
  return PD_OK;
}

// The time transforms of the apply paths.  alpha != 1 (an extension, no upstream counterpart): the inverse transform
// also applies Gamma_alpha to its input and the forward one Gamma_alpha^-1 to its output, inside the kernels where
// they can (every kernel but the opt-in N_t = 16384 cluster variant), else as a separate elementwise launch.
static int apply_fft(pd_handle* h, const cplx* in, cplx* out, int64_t nlines, int inverse, cudaStream_t st) {
  if (h->cfg.alpha == 1.0) return pd_fft_launch(h, in, out, nlines, inverse, st);
  if (pd_fft_gamma_fused(h)) return pd_fft_launch(h, in, out, nlines, inverse, st, 1);
  int rc;
  if (inverse) {
    if ((rc = pd_gamma_launch(h, in, out, nlines, 0, st))) return rc;
    return pd_fft_launch(h, out, out, nlines, 1, st);
  }
  if ((rc = pd_fft_launch(h, in, out, nlines, 0, st))) return rc;
  return pd_gamma_launch(h, out, out, nlines, 1, st);
}
// real lines <-> half spectra of `nnodes` nodes (both fields), Gamma_alpha fused when alpha != 1
static int apply_rfft_pair(pd_handle* h, const void* in, void* out, int64_t nnodes, int to_freq, cudaStream_t st) {
  const int g = h->cfg.alpha != 1.0;
  int rc = pd_rfft_pair_launch(h, in, out, nnodes, to_freq, st, g);
  if (rc == -100) rc = pd_rfft_launch(h, in, out, 2 * nnodes, to_freq, st, g);
  return rc;
}

// The per-frequency part of the slab apply (pass A, interface + peer stores | wait + separator solve, pass B) is run
// as TWO FREQUENCY HALVES ON TWO STREAMS: the interface and separator kernels are short, latency-bound launches
// (one warp per scheduler, a system-scope fence, a wait for the peers) that leave most SMs idle -- with the halves
// staggered they overlap the streaming pass A / pass B of the other half.  Frequencies are independent and the
// exchange flags are per group of 4 frequencies, so the halves share nothing.  MEASURED: no gain (pass A / pass B
// occupy every SM's register file), so it is OFF unless PD_SLAB_OVERLAP=1; it must stay off when one process
// drives several ranks on ONE GPU (LocalSlabGroup; option "slab_overlap" 0, or 2 = split on one stream): there
// the strict order "every first half before any second half" is what keeps a waiting kernel from being scheduled
// ahead of its producer, and a second stream would break it.  The per-stage profile always runs unsplit (its
// stage times are the un-overlapped costs).
static bool slab_overlap(pd_handle* h, int real_input, int* split) {
  static int env = -1;
  if (env < 0) {  // measured: no gain (0.440 vs 0.436 ms at cfg3 on 8 GPUs) -> off unless asked for
    const char* e = getenv("PD_SLAB_OVERLAP");
    env = (e && e[0] == '1') ? 1 : 0;
  }
  const int K = real_input ? ((h->cfg.N_t / 2 + 1 + 7) & ~7) : h->cfg.N_t;
  if ((!env && h->opt_slab_no_overlap != 2) || h->opt_slab_no_overlap == 1 || K < 1024 || h->fuse_on) return false;
  *split = ((K / 2 + 127) / 128) * 128;
  return true;
}

// first half: time transform of this rank's lines, local elimination, functionals pushed to every rank
static int slab_begin(pd_handle* h, const void* x, cudaStream_t st, int real_input, cudaEvent_t* ev, bool overlap) {
  int rc = ensure_work(h);
  if (rc) return rc;
  int fused = 0;
  if (real_input) {
    rc = apply_rfft_pair(h, x, h->work, h->n, 1, st);
  } else if (h->fuse_on && h->cfg.alpha == 1.0) {
    rc = pd_fused_ifft_passA_launch(h, (const cplx*)x, h->work, st, pd_slab_lastl(h));
    fused = 1;
  } else {
    rc = apply_fft(h, (const cplx*)x, h->work, 2 * (int64_t)h->n, 1, st);
  }
  if (rc) return rc;
  if (ev) cudaEventRecord(ev[0], st);
  int split = 0;
  if (overlap && !ev && slab_overlap(h, real_input, &split)) {
    cudaStream_t sb = st;  // "slab_overlap" 2: the two halves one after the other on the caller's stream (tests)
    if (h->opt_slab_no_overlap != 2) {
      if ((rc = ensure_aux(h))) return rc;
      sb = h->sched_aux;
      PD_CUDA(cudaEventRecord(h->sched_ev[0], st));
      PD_CUDA(cudaStreamWaitEvent(sb, h->sched_ev[0], 0));
    }
    if ((rc = pd_slab_reduce_launch(h, h->work, nullptr, st, real_input, nullptr, fused, 0, split))) return rc;
    return pd_slab_reduce_launch(h, h->work, nullptr, sb, real_input, nullptr, fused, split, 1 << 30);
  }
  return pd_slab_reduce_launch(h, h->work, nullptr, st, real_input, ev ? ev + 1 : nullptr, fused);
}

// second half: wait for the peers' functionals, separator solve, back-substitution, time transform
static int slab_end(pd_handle* h, void* y, cudaStream_t st, int real_input, cudaEvent_t* ev, bool overlap) {
  int rc, split = 0;
  if (overlap && !ev && slab_overlap(h, real_input, &split)) {
    cudaStream_t sb = h->opt_slab_no_overlap != 2 ? h->sched_aux : st;
    if ((rc = pd_slab_finish_launch(h, h->work, nullptr, st, real_input, nullptr, 0, split))) return rc;
    if ((rc = pd_slab_finish_launch(h, h->work, nullptr, sb, real_input, nullptr, split, 1 << 30))) return rc;
    if (sb != st) {
      PD_CUDA(cudaEventRecord(h->sched_ev[1], sb));
      PD_CUDA(cudaStreamWaitEvent(st, h->sched_ev[1], 0));
    }
    if ((rc = pd_slab_epoch_bump_launch(h, st))) return rc;  // after the join: the apply's exchange is complete
  } else {
    if ((rc = pd_slab_finish_launch(h, h->work, nullptr, st, real_input, ev, 0, 0, 1))) return rc;  // pass B bumps
  }
  if (ev) cudaEventRecord(ev[1], st);
  if (real_input) return apply_rfft_pair(h, h->work, y, h->n, 0, st);
  return apply_fft(h, h->work, (cplx*)y, 2 * (int64_t)h->n, 0, st);
}

extern "C" int pd_slab_apply_begin(pd_handle* h, const void* x_dev, void* stream, int real_input) {
  int rc = slab_apply_check(h, x_dev, "pd_slab_apply_begin", real_input);
  if (rc) return rc;
  PD_ON_DEVICE(h);
  return slab_begin(h, x_dev, (cudaStream_t)stream, real_input, nullptr, true);
}

extern "C" int pd_slab_apply_end(pd_handle* h, void* y_dev, void* stream, int real_input) {
  int rc = slab_apply_check(h, y_dev, "pd_slab_apply_end", real_input);
  if (rc) return rc;
  PD_ON_DEVICE(h);
  return slab_end(h, y_dev, (cudaStream_t)stream, real_input, nullptr, true);
}

extern "C" int pd_slab_apply(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  int rc = slab_apply_check(h, x_dev, "pd_slab_apply", 0);
  if (rc) return rc;
  if (!y_dev) {
    pd_set_error("pd_slab_apply: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if ((rc = slab_begin(h, x_dev, (cudaStream_t)stream, 0, nullptr, true))) return rc;
  return slab_end(h, y_dev, (cudaStream_t)stream, 0, nullptr, true);
}

extern "C" int pd_slab_apply_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  int rc = slab_apply_check(h, x_dev, "pd_slab_apply_real", 1);
  if (rc) return rc;
  if (!y_dev) {
    pd_set_error("pd_slab_apply_real: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if ((rc = slab_begin(h, x_dev, (cudaStream_t)stream, 1, nullptr, true))) return rc;
  return slab_end(h, y_dev, (cudaStream_t)stream, 1, nullptr, true);
}

// one distributed apply with CUDA events between its stages: ms[0..6] = {inverse FFT, pass A, interface levels,
// functionals + peer stores, wait for the peers + separator solve, pass B, forward FFT}.  Synchronises.
extern "C" int pd_slab_apply_profile(pd_handle* h, const void* x_dev, void* y_dev, void* stream, float* ms, int nms) {
  int rc = slab_apply_check(h, x_dev, "pd_slab_apply_profile", 0);
  if (rc) return rc;
  if (!y_dev || !ms || nms < 7) {
    pd_set_error("pd_slab_apply_profile: invalid argument (need ms[7])");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  struct Events {
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Events() {
      for (cudaEvent_t e : ev)
        if (e) cudaEventDestroy(e);
    }
  } E;
  for (int i = 0; i < 8; ++i) PD_CUDA(cudaEventCreate(&E.ev[i]));
  PD_CUDA(cudaEventRecord(E.ev[0], st));
  if ((rc = slab_begin(h, x_dev, st, 0, &E.ev[1], false))) return rc;   // ev[1] ifft, ev[2] pass A, ev[3] interface
  PD_CUDA(cudaEventRecord(E.ev[4], st));                          // functionals + peer stores
  if ((rc = slab_end(h, y_dev, st, 0, &E.ev[5], false))) return rc;      // ev[5] separator solve, ev[6] pass B
  PD_CUDA(cudaEventRecord(E.ev[7], st));
  PD_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 7; ++i) PD_CUDA(cudaEventElapsedTime(&ms[i], E.ev[i], E.ev[i + 1]));
  return PD_OK;
}

// both fields of `nnodes` node lines at once: (2, nnodes, N_t) float64 <-> (2, nnodes, Kp) complex half spectra
// (one complex N_t-point transform per node where the pair kernel covers N_t, else the per-line kernel)
extern "C" int pd_stage_rfft_pair(pd_handle* h, const void* in_dev, void* out_dev, int64_t nnodes, int to_freq,
                                  void* stream) {
  if (!h || !in_dev || !out_dev || nnodes < 0) {
    pd_set_error("pd_stage_rfft_pair: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (!pd_rfft_supported(h)) {
    pd_set_error("pd_stage_rfft_pair: needs N_t >= 8 (got %d)", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  int rc = pd_rfft_pair_launch(h, in_dev, out_dev, nnodes, to_freq, (cudaStream_t)stream);
  if (rc == -100) rc = pd_rfft_launch(h, in_dev, out_dev, 2 * nnodes, to_freq, (cudaStream_t)stream);
  return rc;
}

extern "C" int pd_pc_apply(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev) {
    pd_set_error("pd_pc_apply: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (h->kcount != h->cfg.N_t || h->nloc != h->n || h->slab_count > 1) {
    pd_set_error("pd_pc_apply: handle is sharded (k_count/n_local/slab set); use the stage API");
    return PD_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_work(h);
  if (rc) return rc;
  const int64_t nlines = 2 * (int64_t)h->n;
  if (h->cfg.alpha != 1.0) {
    // extension: Gamma fused into the loads of the inverse FFT, alpha-shifted per-frequency stage, Gamma^-1 fused
    // into the stores of the forward FFT -- the same three launch groups as alpha = 1
    if ((rc = apply_fft(h, (const cplx*)x_dev, h->work, nlines, 1, st))) return rc;
    if ((rc = pd_solve_launch(h, h->work, st))) return rc;
    return apply_fft(h, h->work, (cplx*)y_dev, nlines, 0, st);
  }
  if (h->sched_chunks > 0) return apply_interleaved(h, (const cplx*)x_dev, (cplx*)y_dev, st);
  // :500-501 ifft along time, :445-540 per-frequency stage, :547-548 fft along time
  if (h->fuse_on) {
    // inverse FFT and pass A in one launch (pass A reads its rows out of L2), then the interface and pass B
    if ((rc = pd_fused_ifft_passA_launch(h, (const cplx*)x_dev, h->work, st, nullptr))) return rc;
    if ((rc = pd_solve_interface(h, st))) return rc;
    if ((rc = pd_solve_passB_range(h, h->work, 0, pd_solve_nchunks(h), st))) return rc;
    return pd_fft_launch(h, h->work, (cplx*)y_dev, nlines, 0, st);
  }
  if ((rc = pd_fft_launch(h, (const cplx*)x_dev, h->work, nlines, 1, st))) return rc;
  if ((rc = pd_solve_launch(h, h->work, st))) return rc;
  if ((rc = pd_fft_launch(h, h->work, (cplx*)y_dev, nlines, 0, st))) return rc;
  return PD_OK;
}

extern "C" int pd_pc_apply_profile(pd_handle* h, const void* x_dev, void* y_dev, void* stream, float* ms,
                                   int nms) {
  if (!h || !x_dev || !y_dev || !ms || nms < 5) {
    pd_set_error("pd_pc_apply_profile: invalid argument (need ms[5])");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (h->kcount != h->cfg.N_t || h->nloc != h->n || h->slab_count > 1) {
    pd_set_error("pd_pc_apply_profile: handle is sharded");
    return PD_ERR_INVALID;
  }
  if (h->cfg.alpha != 1.0) {
    pd_set_error("pd_pc_apply_profile: alpha != 1 is not covered by the per-stage profile");
    return PD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_work(h);
  if (rc) return rc;
  // events are destroyed on every path out of this function
  struct Events {
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Events() {
      for (cudaEvent_t e : ev)
        if (e) cudaEventDestroy(e);
    }
  } E;
  cudaEvent_t* ev = E.ev;
  for (int i = 0; i < 6; ++i) PD_CUDA(cudaEventCreate(&ev[i]));
  const int64_t nlines = 2 * (int64_t)h->n;
  PD_CUDA(cudaEventRecord(ev[0], st));
  if (h->fuse_on && pd_solve_nchunks(h) > 0) {
    // the path pd_pc_apply takes: ms[0] = fused inverse FFT + pass A (one launch), ms[1] = 0
    rc = pd_fused_ifft_passA_launch(h, (const cplx*)x_dev, h->work, st, nullptr);
    PD_CUDA(cudaEventRecord(ev[1], st));
    PD_CUDA(cudaEventRecord(ev[2], st));
    if (!rc) rc = pd_solve_interface(h, st);
    PD_CUDA(cudaEventRecord(ev[3], st));
    if (!rc) rc = pd_solve_passB_range(h, h->work, 0, pd_solve_nchunks(h), st);
  } else {
    rc = pd_fft_launch(h, (const cplx*)x_dev, h->work, nlines, 1, st);
    PD_CUDA(cudaEventRecord(ev[1], st));
    if (!rc) rc = pd_solve_launch(h, h->work, st, &ev[2]);  // records ev[2] after pass A, ev[3] after PCR
  }
  PD_CUDA(cudaEventRecord(ev[4], st));
  if (!rc) rc = pd_fft_launch(h, h->work, (cplx*)y_dev, nlines, 0, st);
  PD_CUDA(cudaEventRecord(ev[5], st));
  PD_CUDA(cudaStreamSynchronize(st));
  if (!rc)
    for (int i = 0; i < 5; ++i) PD_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  return rc;
}

extern "C" int pd_stage_rfft(pd_handle* h, const void* in_dev, void* out_dev, int64_t nlines, int to_freq,
                             void* stream) {
  if (!h || !in_dev || !out_dev || nlines < 0) {
    pd_set_error("pd_stage_rfft: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_rfft_launch(h, in_dev, out_dev, nlines, to_freq, (cudaStream_t)stream);
}

extern "C" int pd_stage_solve_half(pd_handle* h, void* w_dev, void* stream) {
  if (!h || !w_dev || h->kcount != h->cfg.N_t || h->slab_count > 1) {
    pd_set_error("pd_stage_solve_half: invalid argument or sharded handle");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (!pd_rfft_supported(h)) {  // the interface workspaces are sized for N_t columns; Kp > N_t for N_t < 8
    pd_set_error("pd_stage_solve_half: needs N_t >= 8 (got %d)", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  return pd_solve_launch(h, (cplx*)w_dev, (cudaStream_t)stream, nullptr, 1);
}

extern "C" int pd_pc_apply_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev) {
    pd_set_error("pd_pc_apply_real: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (h->kcount != h->cfg.N_t || h->nloc != h->n || h->slab_count > 1) {
    pd_set_error("pd_pc_apply_real: handle is sharded; the real-input path is single-GPU");
    return PD_ERR_INVALID;
  }
  if (!pd_rfft_supported(h)) {
    pd_set_error("pd_pc_apply_real: needs N_t >= 8 (got %d); use pd_pc_apply", h->cfg.N_t);
    return PD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_work(h);
  if (rc) return rc;
  // real lines -> half spectra (k = 0..N_t/2), the same per-frequency stage on half as many columns, back
  // u- and p-line of a node go through ONE complex N_t-point transform (pd_rfft_pair_kernel); sizes it does
  // not cover use the per-line packed kernel
  // (alpha != 1: Gamma is real, so Gamma x stays real and the alpha-shifted symbols keep lambda(N_t - k) =
  // conj lambda(k): the half spectrum still suffices; Gamma / Gamma^-1 ride on the loads / stores of the transforms)
  if ((rc = apply_rfft_pair(h, x_dev, h->work, h->n, 1, st))) return rc;
  if ((rc = pd_solve_launch(h, h->work, st, nullptr, 1))) return rc;
  return apply_rfft_pair(h, h->work, y_dev, h->n, 0, st);
}

extern "C" int pd_pc_apply_transpose(pd_handle*, const void*, void*, void*) {
  pd_set_error("applyTranspose is not implemented (the upstream PC raises NotImplementedError)");
  return PD_ERR_UNSUPPORTED;
}

// OPT-IN (pd_set_option "host_register" 1, or PD_HOST_REGISTER=1): host buffers handed to pd_pc_apply_host are
// page-locked ONCE per (pointer, size) with cudaHostRegister and remembered.  A PETSc Vec array is pageable, and a
// pageable cudaMemcpyAsync runs at about half the PCIe rate (staged through the driver's bounce buffer); KSP work
// vectors keep their arrays for the life of the solve, so the registration cost (~0.3 s per GB, once) is paid on
// the first apply only.  Up to PD_HOSTREG_SLOTS buffers are kept (oldest evicted).  Memory that is already
// page-locked (cudaHostAlloc, torch pin_memory) is detected and left alone.
// It is opt-in because the library does not own these buffers: if the caller frees one while it is registered and
// the allocator hands the same address out again, the stale registration would be used for the new buffer.  The
// host must either keep its vectors alive as long as the handle (KSP work vectors) or call
// pd_host_unregister_all before freeing them.
#define PD_HOSTREG_SLOTS 16
struct HostReg {
  const void* ptr[PD_HOSTREG_SLOTS];
  size_t bytes[PD_HOSTREG_SLOTS];
  bool ours[PD_HOSTREG_SLOTS];  // registered by us (to be unregistered)
  int next;
  int enabled;  // -1 unknown
};

static HostReg* hostreg_of(pd_handle* h) {
  if (!h->pinned) {
    HostReg* r = new HostReg();
    memset(r, 0, sizeof(*r));
    r->enabled = -1;
    h->pinned = r;
  }
  return reinterpret_cast<HostReg*>(h->pinned);
}

static void hostreg_release(pd_handle* h) {
  if (!h->pinned) return;
  HostReg* r = reinterpret_cast<HostReg*>(h->pinned);
  for (int i = 0; i < PD_HOSTREG_SLOTS; ++i)
    if (r->ptr[i] && r->ours[i]) cudaHostUnregister(const_cast<void*>(r->ptr[i]));
  cudaGetLastError();
  delete r;
  h->pinned = nullptr;
}

static void hostreg_pin(pd_handle* h, const void* p, size_t bytes) {
  HostReg* r = hostreg_of(h);
  if (r->enabled < 0) {
    const char* e = getenv("PD_HOST_REGISTER");
    r->enabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (!r->enabled && !h->opt_host_register) return;
  for (int i = 0; i < PD_HOSTREG_SLOTS; ++i)
    if (r->ptr[i] == p && r->bytes[i] >= bytes) return;
  cudaPointerAttributes at;
  bool already = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  bool ours = false;
  if (!already) {
    ours = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();  // failure (e.g. overlapping registration, locked-memory limit) just means a pageable copy
  }
  const int slot = r->next;
  r->next = (r->next + 1) % PD_HOSTREG_SLOTS;
  if (r->ptr[slot] && r->ours[slot]) {
    cudaHostUnregister(const_cast<void*>(r->ptr[slot]));
    cudaGetLastError();
  }
  r->ptr[slot] = p;
  r->bytes[slot] = bytes;
  r->ours[slot] = ours;
}

extern "C" int pd_host_unregister_all(pd_handle* h) {
  if (!h) return PD_OK;
  PD_ON_DEVICE(h);
  hostreg_release(h);
  return PD_OK;
}

extern "C" int pd_pc_apply_host(pd_handle* h, const void* x_host, void* y_host) {
  if (!h || !x_host || !y_host) {
    pd_set_error("pd_pc_apply_host: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  const size_t bytes = sizeof(cplx) * 2 * (size_t)h->n * h->cfg.N_t;
  if (!h->stage_x) {
    PD_CUDA(cudaMalloc(&h->stage_x, bytes));
    h->ws_bytes += bytes;
  }
  if (!h->own_stream) PD_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  cudaStream_t st = h->own_stream;
  hostreg_pin(h, x_host, bytes);
  if (y_host != x_host) hostreg_pin(h, y_host, bytes);
  PD_CUDA(cudaMemcpyAsync(h->stage_x, x_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = pd_pc_apply(h, h->stage_x, h->stage_x, st);
  if (rc) return rc;
  PD_CUDA(cudaMemcpyAsync(y_host, h->stage_x, bytes, cudaMemcpyDeviceToHost, st));
  PD_CUDA(cudaStreamSynchronize(st));
  return PD_OK;
}

// The same for float64 host vectors (a real-scalar PETSc build, or numpy float64 arrays): the real-input path of
// pd_pc_apply_real between the two copies -- half the PCIe bytes of pd_pc_apply_host, which is what bounds it.
extern "C" int pd_pc_apply_real_host(pd_handle* h, const void* x_host, void* y_host) {
  if (!h || !x_host || !y_host) {
    pd_set_error("pd_pc_apply_real_host: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  const size_t bytes = sizeof(double) * 2 * (size_t)h->n * h->cfg.N_t;
  if (!h->stage_x) {  // sized for the complex entry point, which shares it
    PD_CUDA(cudaMalloc(&h->stage_x, 2 * bytes));
    h->ws_bytes += 2 * bytes;
  }
  if (!h->own_stream) PD_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  cudaStream_t st = h->own_stream;
  hostreg_pin(h, x_host, bytes);
  if (y_host != x_host) hostreg_pin(h, y_host, bytes);
  PD_CUDA(cudaMemcpyAsync(h->stage_x, x_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = pd_pc_apply_real(h, h->stage_x, h->stage_x, st);
  if (rc) return rc;
  PD_CUDA(cudaMemcpyAsync(y_host, h->stage_x, bytes, cudaMemcpyDeviceToHost, st));
  PD_CUDA(cudaStreamSynchronize(st));
  return PD_OK;
}

extern "C" int pd_matvec(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev || x_dev == y_dev) {
    pd_set_error("pd_matvec: invalid argument (x and y must be distinct device vectors)");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (h->slab_count > 1) {
    pd_set_error("pd_matvec: handle is in slab mode; use pd_matvec_slab");
    return PD_ERR_INVALID;
  }
  return pd_matvec_launch(h, (const cplx*)x_dev, (cplx*)y_dev, (cudaStream_t)stream, 0);
}

extern "C" int pd_matvec_slab(pd_handle* h, const void* x_dev, const void* halo_lo_dev, const void* halo_hi_dev,
                              void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev || x_dev == y_dev || h->slab_count <= 1) {
    pd_set_error("pd_matvec_slab: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_matvec_launch(h, (const cplx*)x_dev, (cplx*)y_dev, (cudaStream_t)stream, 0, (const cplx*)halo_lo_dev,
                          (const cplx*)halo_hi_dev);
}

extern "C" int pd_matvec_slab_real(pd_handle* h, const void* x_dev, const void* halo_lo_dev, const void* halo_hi_dev,
                                   void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev || x_dev == y_dev || h->slab_count <= 1) {
    pd_set_error("pd_matvec_slab_real: invalid argument or handle not in slab mode");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_matvec_launch(h, (const cplx*)x_dev, (cplx*)y_dev, (cudaStream_t)stream, 0, (const cplx*)halo_lo_dev,
                          (const cplx*)halo_hi_dev, 1);
}

extern "C" int pd_pc_matvec(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev || x_dev == y_dev) {
    pd_set_error("pd_pc_matvec: invalid argument (x and y must be distinct device vectors)");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  if (h->cfg.alpha != 1.0) {
    pd_set_error("pd_pc_matvec: only the alpha = 1 block circulant has a matvec kernel");
    return PD_ERR_UNSUPPORTED;
  }
  return pd_matvec_launch(h, (const cplx*)x_dev, (cplx*)y_dev, (cudaStream_t)stream, 1);
}

extern "C" int pd_matvec_real(pd_handle* h, const void* x_dev, void* y_dev, void* stream) {
  if (!h || !x_dev || !y_dev || x_dev == y_dev || h->slab_count > 1) {
    pd_set_error("pd_matvec_real: invalid argument (distinct float64 device vectors, unsharded handle)");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_matvec_launch(h, (const cplx*)x_dev, (cplx*)y_dev, (cudaStream_t)stream, 0, nullptr, nullptr, 1);
}

extern "C" int pd_build_rhs_real(pd_handle* h, void* b_dev, void* stream) {
  if (!h || !b_dev) {
    pd_set_error("pd_build_rhs_real: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_rhs_launch(h, (cplx*)b_dev, (cudaStream_t)stream, 1);
}

extern "C" int pd_build_rhs(pd_handle* h, void* b_dev, void* stream) {
  if (!h || !b_dev) {
    pd_set_error("pd_build_rhs: invalid argument");
    return PD_ERR_INVALID;
  }
  PD_ON_DEVICE(h);
  return pd_rhs_launch(h, (cplx*)b_dev, (cudaStream_t)stream);
}
