/* TEST INFRASTRUCTURE ONLY -- plain C restatement of the reference's algebraic-distance
 * relaxation, used as the checker at sizes the Python port is too slow for and as the CPU
 * baseline of bench.py (`cpu_baseline`, `--impl reference`).  Never linked into the product.
 *
 * Follows hypergraph_embedding/algebraic_distance.py of the reference, in f64 like numpy:
 *   _update_alg_dist   :34-51   new_a = (a + sum_b(e_b * w_b) / sum_b(w_b)) / 2,
 *                               w_b = 1 / nnz(B2A[b]), neighbours in ascending order
 *   _helper_update_embeddings :54-91   node half from the OLD edge rows, then the edge half
 *                               from the NEW node rows (Jacobi inside a half)
 *   _helper_scale_embeddings  :97-123  joint per-column min / max, x -= min, x /= delta
 * Rows are independent inside a half-sweep (the reference farms them out to a process pool,
 * :59-72); here they are pthreads pulling blocks of rows.  Compiled without FMA contraction so that the
 * arithmetic is the same sequence of IEEE f64 operations numpy performs.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- a minimal fork-join over row blocks (this image has no libgomp) ----------------------- */
typedef void (*block_fn)(int64_t r0, int64_t r1, void* arg);

typedef struct {
  block_fn fn;
  void* arg;
  int64_t rows, block;
  atomic_llong next;
} job_t;

static void* job_worker(void* p) {
  job_t* j = (job_t*)p;
  for (;;) {
    const int64_t r0 = (int64_t)atomic_fetch_add(&j->next, (long long)j->block);
    if (r0 >= j->rows) break;
    const int64_t r1 = r0 + j->block < j->rows ? r0 + j->block : j->rows;
    j->fn(r0, r1, j->arg);
  }
  return NULL;
}

static void parallel_rows(int threads, int64_t rows, int64_t block, block_fn fn, void* arg) {
  job_t j;
  j.fn = fn;
  j.arg = arg;
  j.rows = rows;
  j.block = block;
  atomic_init(&j.next, 0);
  if (threads <= 1) {
    job_worker(&j);
    return;
  }
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  int started = 0;
  for (int t = 0; t < threads - 1; ++t)
    if (pthread_create(&tid[started], NULL, job_worker, &j) == 0) ++started;
  job_worker(&j);
  for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
  free(tid);
}

typedef struct {
  const int64_t* a_ptr;
  const int32_t* a_idx;
  const int64_t* b_ptr;
  int R;
  const double* xa;
  const double* xb;
  double* out;
} half_args;

static void half_block(int64_t r0, int64_t r1, void* p) {
  const half_args* h = (const half_args*)p;
  const int R = h->R;
  double* acc = (double*)malloc(sizeof(double) * (size_t)R);
  for (int64_t a = r0; a < r1; ++a) {
    double wsum = 0.0;
    for (int c = 0; c < R; ++c) acc[c] = 0.0;
    for (int64_t q = h->a_ptr[a]; q < h->a_ptr[a + 1]; ++q) {
      const int32_t b = h->a_idx[q];
      const double w = 1.0 / (double)(h->b_ptr[b + 1] - h->b_ptr[b]);
      const double* e = h->xb + (size_t)b * R;
      for (int c = 0; c < R; ++c) acc[c] = acc[c] + e[c] * w;
      wsum = wsum + w;
    }
    const double* own = h->xa + (size_t)a * R;
    double* o = h->out + (size_t)a * R;
    for (int c = 0; c < R; ++c) o[c] = (own[c] + acc[c] / wsum) / 2.0;
  }
  free(acc);
}

static void half_sweep(int threads, int64_t rows, const int64_t* a_ptr, const int32_t* a_idx,
                       const int64_t* b_ptr, int R, const double* xa, const double* xb,
                       double* out) {
  half_args h = {a_ptr, a_idx, b_ptr, R, xa, xb, out};
  parallel_rows(threads, rows, 256, half_block, &h);
}

static void column_minmax(int64_t rows, int R, const double* x, double* lo, double* hi) {
  for (int64_t r = 0; r < rows; ++r)
    for (int c = 0; c < R; ++c) {
      const double v = x[(size_t)r * R + c];
      if (v < lo[c]) lo[c] = v;
      if (v > hi[c]) hi[c] = v;
    }
}

typedef struct {
  int R;
  double* x;
  const double* lo;
  const double* delta;
} scale_args;

static void scale_block(int64_t r0, int64_t r1, void* p) {
  const scale_args* s = (const scale_args*)p;
  for (int64_t r = r0; r < r1; ++r)
    for (int c = 0; c < s->R; ++c) {
      double v = s->x[(size_t)r * s->R + c];
      v = v - s->lo[c];
      v = v / s->delta[c];
      s->x[(size_t)r * s->R + c] = v;
    }
}

static void rescale(int threads, int64_t rows, int R, double* x, const double* lo,
                    const double* delta) {
  scale_args s = {R, x, lo, delta};
  parallel_rows(threads, rows, 4096, scale_block, &s);
}

int oracle_max_threads(void) {
  const long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* xn [N, R], xe [E, R] in place.  threads <= 0: all online cores.  Returns 0, or -1 on
 * allocation failure. */
int oracle_algdist(int64_t N, int64_t E, const int64_t* n2e_ptr, const int32_t* n2e_idx,
                   const int64_t* e2n_ptr, const int32_t* e2n_idx, int R, int iterations,
                   double* xn, double* xe, int threads) {
  if (threads <= 0) threads = oracle_max_threads();
  double* new_n = (double*)malloc(sizeof(double) * (size_t)N * R);
  double* new_e = (double*)malloc(sizeof(double) * (size_t)E * R);
  double* lo = (double*)malloc(sizeof(double) * 3 * (size_t)R);
  if (!new_n || !new_e || !lo) {
    free(new_n);
    free(new_e);
    free(lo);
    return -1;
  }
  double* hi = lo + R;
  double* delta = hi + R;
  for (int it = 0; it < iterations; ++it) {
    half_sweep(threads, N, n2e_ptr, n2e_idx, e2n_ptr, R, xn, xe, new_n);
    half_sweep(threads, E, e2n_ptr, e2n_idx, n2e_ptr, R, xe, new_n, new_e);
    for (int c = 0; c < R; ++c) {
      lo[c] = INFINITY;
      hi[c] = -INFINITY;
    }
    column_minmax(N, R, new_n, lo, hi);
    column_minmax(E, R, new_e, lo, hi);
    for (int c = 0; c < R; ++c) delta[c] = hi[c] - lo[c];
    rescale(threads, N, R, new_n, lo, delta);
    rescale(threads, E, R, new_e, lo, delta);
    memcpy(xn, new_n, sizeof(double) * (size_t)N * R);
    memcpy(xe, new_e, sizeof(double) * (size_t)E * R);
  }
  free(new_n);
  free(new_e);
  free(lo);
  return 0;
}
