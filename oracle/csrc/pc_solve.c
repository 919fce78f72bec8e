/*
 * CPU restatement of the frequency-domain solve stage of DiagFFTPC -- ORACLE CODE
 * (test infrastructure / timed CPU baseline; never linked into the product).
 *
 * Follows Code/Control_Wave_PC.py:460-484 and :512 of the upstream repository: per
 * frequency k two shifted tridiagonal systems (Sigma_i(k) M + dt^2/2 K) w = rhs with
 * homogeneous Dirichlet rows, which upstream hands to MUMPS as one monolithic LU.
 * On the uniform 1-D P1 mesh each system is Toeplitz tridiagonal, tridiag(a_k, b_k, a_k)
 * (see oracle/pc_fast.py for the closed forms), and LU without pivoting is the Thomas
 * recurrence below.  Same arithmetic order as oracle/pc_fast.py:thomas_toeplitz.
 *
 * Layout: rhs[(row)*ld + k], row = interior node (m rows), k fastest.
 * Threads (pthreads; this image has no OpenMP runtime) split the frequency axis in
 * blocks of KB columns, each block sweeps all rows.
 */
#include <complex.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <unistd.h>

typedef double complex cplx;

#define KB 16

typedef struct {
  int m, K, conj_mode, nblk;
  long ld;
  const cplx *a, *b;
  cplx* rhs;
  atomic_int next;
  atomic_int err;
} job_t;

static void solve_block(const job_t* J, int blk, cplx* cp) {
  const int m = J->m, K = J->K;
  const long ld = J->ld;
  int k0 = blk * KB, kn = (k0 + KB <= K) ? KB : K - k0;
  cplx aa[KB], bb[KB];
  for (int t = 0; t < kn; ++t) {
    aa[t] = J->conj_mode ? conj(J->a[k0 + t]) : J->a[k0 + t];
    bb[t] = J->conj_mode ? conj(J->b[k0 + t]) : J->b[k0 + t];
  }
  cplx* d = J->rhs + k0;
  for (int t = 0; t < kn; ++t) {
    cplx inv = 1.0 / bb[t];
    cp[t] = aa[t] * inv;
    d[t] = d[t] * inv;
  }
  for (int i = 1; i < m; ++i) {
    cplx* di = d + (long)i * ld;
    const cplx* dm = di - ld;
    cplx* cpi = cp + (size_t)i * KB;
    const cplx* cpm = cpi - KB;
    for (int t = 0; t < kn; ++t) {
      cplx inv = 1.0 / (bb[t] - aa[t] * cpm[t]);
      cpi[t] = aa[t] * inv;
      di[t] = (di[t] - aa[t] * dm[t]) * inv;
    }
  }
  for (int i = m - 2; i >= 0; --i) {
    cplx* di = d + (long)i * ld;
    const cplx* dp = di + ld;
    const cplx* cpi = cp + (size_t)i * KB;
    for (int t = 0; t < kn; ++t) di[t] -= cpi[t] * dp[t];
  }
}

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  cplx* cp = (cplx*)malloc(sizeof(cplx) * (size_t)J->m * KB);
  if (!cp) {
    atomic_store(&J->err, 1);
    return NULL;
  }
  for (;;) {
    int blk = atomic_fetch_add(&J->next, 1);
    if (blk >= J->nblk) break;
    solve_block(J, blk, cp);
  }
  free(cp);
  return NULL;
}

static int g_threads = 0;

int oracle_num_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

void oracle_set_num_threads(int n) { g_threads = n; }

/* conj_mode != 0 solves with conj(a), conj(b) (the zeta_- systems). */
int oracle_thomas_toeplitz(int m, int K, long ld, const cplx* a, const cplx* b, cplx* rhs, int conj_mode) {
  job_t J;
  J.m = m; J.K = K; J.ld = ld; J.a = a; J.b = b; J.rhs = rhs; J.conj_mode = conj_mode;
  J.nblk = (K + KB - 1) / KB;
  atomic_init(&J.next, 0);
  atomic_init(&J.err, 0);
  int nt = oracle_num_threads();
  if (nt > J.nblk) nt = J.nblk;
  if (nt < 1) nt = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  if (!th) return 1;
  int started = 0;
  for (int t = 1; t < nt; ++t)
    if (pthread_create(&th[started], NULL, worker, &J) == 0) ++started;
  worker(&J);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(th);
  return atomic_load(&J.err);
}

/* ------------------------------------------------------------------------------------------
 * Whole per-frequency stage of DiagFFTPC.apply, fused and threaded (the timed CPU baseline):
 * on xh = ifft_t(x) (Control_Wave_PC.py:500-503), layout [field][node][k], in place
 *   :445-457  rho_+ = (uh / z + i sigma ph) / 2,  rho_- = (uh / z - i sigma ph) / 2
 *   :460-484, :512  Tt zeta_+ = rho_+,  conj(Tt) zeta_- = rho_-  on the interior nodes (Thomas)
 *   :516-540  wh_u = zeta_+ + zeta_-,  wh_p = -i sigma z (zeta_+ - zeta_-);  Dirichlet rows -> 0
 * Same arithmetic as oracle/pc_fast.py (forward_stage / solve_stage / backward_stage without the
 * FFTs), which stays the semantic definition; tests compare the two.
 */
typedef struct {
  int n, K, nblk;
  const cplx *a, *b, *z;
  const double* sigma;
  cplx *u, *p;
  atomic_int next;
  atomic_int err;
} stage_t;

static void stage_block(const stage_t* J, int blk, cplx* cp) {
  const int n = J->n, K = J->K, m = n - 2;
  int k0 = blk * KB, kn = (k0 + KB <= K) ? KB : K - k0;
  cplx aa[KB], bb[KB], zc[KB], isg[KB], osg[KB];
  for (int t = 0; t < kn; ++t) {
    aa[t] = J->a[k0 + t];
    bb[t] = J->b[k0 + t];
    zc[t] = conj(J->z[k0 + t]);
    isg[t] = I * J->sigma[k0 + t];                    /* i sigma */
    osg[t] = -I * J->sigma[k0 + t] * J->z[k0 + t];    /* -i sigma z */
  }
  cplx* U = J->u + k0;
  cplx* P = J->p + k0;
  /* rotate in: U <- rho_+, P <- conj(rho_-)  (conj(Tt) y = r  <=>  Tt conj(y) = conj(r)) */
  for (int j = 1; j <= m; ++j) {
    cplx* uj = U + (long)j * K;
    cplx* pj = P + (long)j * K;
    for (int t = 0; t < kn; ++t) {
      cplx uz = uj[t] * zc[t], ip = isg[t] * pj[t];
      uj[t] = (uz + ip) / 2;
      pj[t] = conj((uz - ip) / 2);
    }
  }
  /* forward elimination, both right-hand sides */
  {
    cplx* u1 = U + K;
    cplx* p1 = P + K;
    for (int t = 0; t < kn; ++t) {
      cplx inv = 1.0 / bb[t];
      cp[t] = aa[t] * inv;
      u1[t] *= inv;
      p1[t] *= inv;
    }
  }
  for (int i = 1; i < m; ++i) {
    cplx* ui = U + (long)(i + 1) * K;
    cplx* pi = P + (long)(i + 1) * K;
    const cplx* um = ui - K;
    const cplx* pm = pi - K;
    cplx* cpi = cp + (size_t)i * KB;
    const cplx* cpm = cpi - KB;
    for (int t = 0; t < kn; ++t) {
      cplx inv = 1.0 / (bb[t] - aa[t] * cpm[t]);
      cpi[t] = aa[t] * inv;
      ui[t] = (ui[t] - aa[t] * um[t]) * inv;
      pi[t] = (pi[t] - aa[t] * pm[t]) * inv;
    }
  }
  /* plain back substitution, then a separate rotation sweep (keeps the recurrence simple) */
  for (int i = m - 2; i >= 0; --i) {
    cplx* ui = U + (long)(i + 1) * K;
    cplx* pi = P + (long)(i + 1) * K;
    const cplx* un = ui + K;
    const cplx* pn = pi + K;
    const cplx* cpi = cp + (size_t)i * KB;
    for (int t = 0; t < kn; ++t) {
      ui[t] -= cpi[t] * un[t];
      pi[t] -= cpi[t] * pn[t];
    }
  }
  for (int j = 1; j <= m; ++j) {
    cplx* uj = U + (long)j * K;
    cplx* pj = P + (long)j * K;
    for (int t = 0; t < kn; ++t) {
      cplx zp = uj[t], zm = conj(pj[t]);
      uj[t] = zp + zm;
      pj[t] = osg[t] * (zp - zm);
    }
  }
  for (int t = 0; t < kn; ++t) {
    U[t] = 0;
    P[t] = 0;
    U[(long)(n - 1) * K + t] = 0;
    P[(long)(n - 1) * K + t] = 0;
  }
}

static void* stage_worker(void* arg) {
  stage_t* J = (stage_t*)arg;
  cplx* cp = (cplx*)malloc(sizeof(cplx) * (size_t)J->n * KB);
  if (!cp) {
    atomic_store(&J->err, 1);
    return NULL;
  }
  for (;;) {
    int blk = atomic_fetch_add(&J->next, 1);
    if (blk >= J->nblk) break;
    stage_block(J, blk, cp);
  }
  free(cp);
  return NULL;
}

/* xh: (2, n, K) complex128, frequency fastest, overwritten by wh.  a, b, z: K complex; sigma: K doubles. */
int oracle_pc_stage(int n, int K, const cplx* a, const cplx* b, const cplx* z, const double* sigma, cplx* xh) {
  stage_t J;
  J.n = n; J.K = K; J.a = a; J.b = b; J.z = z; J.sigma = sigma;
  J.u = xh; J.p = xh + (size_t)n * K;
  J.nblk = (K + KB - 1) / KB;
  atomic_init(&J.next, 0);
  atomic_init(&J.err, 0);
  if (n < 3) return 2;
  int nt = oracle_num_threads();
  if (nt > J.nblk) nt = J.nblk;
  if (nt < 1) nt = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  if (!th) return 1;
  int started = 0;
  for (int t = 1; t < nt; ++t)
    if (pthread_create(&th[started], NULL, stage_worker, &J) == 0) ++started;
  stage_worker(&J);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(th);
  return atomic_load(&J.err);
}
