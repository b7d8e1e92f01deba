// hypergraph2vec training on sm_100a: the consumer of the sample columns
// (hg2v_model.py:51-203 BooleanModel / UnweightedFloatModel, fit loop embedding.py:269-305).
//
//   node_node = act(<N[ln], N[rn]>)            edge_edge = act(<E[le], E[re]>)
//   node_edge = mean_i act(<N[nn_i], N[ln]>) * mean_i act(<E[ne_i], E[re]>)
//   loss = sum over the three outputs of the batch mean (KL divergence or squared error),
//   Adagrad (lr 0.01, eps 1e-7, accumulators from 0) on the two embedding tables, whose row 0 is
//   the (trained) padding row every absent index points at.
//
// Mini-batch SGD is sequential in the batches and a batch of 256 samples is ~0.5 MB of gathers,
// so the whole epoch is ONE kernel launch of ONE thread-block cluster (8 CTAs x 32 warps = 256
// warps, a sample per warp) that walks the batches with two hardware cluster barriers each:
//   phase 1  forward + backward of the warp's samples; gradients of duplicated rows are summed
//            with float atomics into dense gradient tables (row 0, which every sample touches
//            several times, is first reduced in shared memory per CTA);
//   phase 2  every warp re-visits the rows its samples touched; the first to claim a row
//            (atomicExch on a per-row batch stamp) applies the Adagrad update and clears the
//            row's gradient.
// Tables, accumulators and gradients are read with ld.global.cg: other SMs rewrite them every
// batch and L1 is not coherent.  A launch per batch would cost ~3 launches x 3 000 batches per
// epoch on the 755 K-sample fixture run; here the per-batch cost is the two barriers.
//
// Batches of more than 256 samples: the same kernel is launched (cooperatively, so that every
// block is resident) as ceil(batch / 256) clusters, up to what the device holds at once (16-18 on
// B200), a sample per warp again, and the two per-batch barriers become cluster barrier ->
// one arrival per cluster on a global counter -> cluster barrier.  The one-cluster launch keeps
// the pure hardware barrier.
#include <cooperative_groups.h>

#include <algorithm>
#include <new>

#include "hge_common.cuh"
#include "hge_staged.cuh"

namespace cg = cooperative_groups;

struct hge_hg2v_model {
  hge_ctx* ctx = nullptr;
  int32_t node_rows = 0, edge_rows = 0;
  int dim = 0, k = 0, activation = 0, loss = 0;
  float *N = nullptr, *E = nullptr, *accN = nullptr, *accE = nullptr, *gN = nullptr, *gE = nullptr;
  int32_t *claimN = nullptr, *claimE = nullptr;
  int32_t* feat = nullptr;   // [4 + 2k][M]
  float* target = nullptr;   // [3][M]
  int32_t* order = nullptr;  // [M]
  double* d_loss = nullptr;
  int64_t M = 0;
  int32_t next_batch_id = 0;
  unsigned int* grid_bar = nullptr;   // arrival counter of the inter-cluster batch barrier
  int max_clusters = 0;               // co-resident clusters of k_hg2v_epoch (0: not asked yet)
  int last_clusters = 0;              // clusters of the last epoch's launch
};

namespace {

constexpr int kCluster = 8;
constexpr int kThreads = 1024;
constexpr int kWarpsPerCta = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxDim = 256;

struct TrainArgs {
  float *N, *E, *accN, *accE, *gN, *gE;
  int32_t *claimN, *claimE;
  const int32_t* feat;
  const float* target;
  const int32_t* order;
  double* loss_sum;
  int64_t M;
  int dim, k, batch, activation, loss, batch_id0;
  float lr, eps;
  int clusters;             // clusters of this launch
  unsigned int* grid_bar;   // zero at launch
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

__device__ __forceinline__ float activate(float z, int kind) {
  return kind == 0 ? 1.0f / (1.0f + __expf(-z)) : fmaxf(z, 0.0f);
}
__device__ __forceinline__ float activate_grad(float z, float p, int kind) {
  return kind == 0 ? p * (1.0f - p) : (z > 0.0f ? 1.0f : 0.0f);
}
// per-sample loss and d loss / d p (hge_hg2v: 0 = KL divergence with clipping, 1 = squared error)
__device__ __forceinline__ float loss_grad(float p, float y, int kind, float* loss) {
  if (kind == 0) {
    const float e = 1e-7f;
    const float yc = fminf(fmaxf(y, e), 1.0f), pc = fminf(fmaxf(p, e), 1.0f);
    *loss = yc * __logf(yc / pc);
    return (p >= e && p <= 1.0f) ? -yc / pc : 0.0f;
  }
  const float d = p - y;
  *loss = d * d;
  return 2.0f * d;
}

// gradient of one table row: row 0 goes to the CTA's shared accumulator, the rest to global
template <int DPL>
__device__ __forceinline__ void add_grad(float* __restrict__ g, float* s_g0, int32_t row, int dim,
                                         int lane, float coef, const float (&v)[DPL]) {
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int c = lane + 32 * j;
    if (c < dim) {
      if (row == 0) atomicAdd(s_g0 + c, coef * v[j]);
      else atomicAdd(g + (size_t)row * dim + c, coef * v[j]);
    }
  }
}

template <int DPL>
__device__ __forceinline__ void load_row(const float* __restrict__ t, int32_t row, int dim, int lane,
                                         float (&v)[DPL]) {
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < dim ? __ldcg(t + (size_t)row * dim + c) : 0.0f;
  }
}

template <int DPL>
__device__ __forceinline__ float dot(const float (&a)[DPL], const float (&b)[DPL]) {
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < DPL; ++j) s = fmaf(a[j], b[j], s);
  return warp_sum(s);
}

template <int DPL>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    k_hg2v_epoch(const TrainArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_g0[2][kMaxDim];   // gradient of the padding rows N[0], E[0] of this CTA
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int num_warps = a.clusters * kCluster * kWarpsPerCta;
  // batch barrier: everything every block of the launch wrote (gradient atomics, table updates)
  // is visible to every block after it
  unsigned int bar_target = 0;
  auto batch_barrier = [&]() {
    if (a.clusters == 1) {
      // barrier.cluster arrive.release / wait.acquire: every CTA that touches the tables is in
      // this cluster, so cluster scope orders the gradient atomics before the updates
      cluster.sync();
      return;
    }
    __threadfence();
    cluster.sync();
    bar_target += (unsigned int)a.clusters;
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.grid_bar, 1u);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.grid_bar) : "memory");
      } while (seen < bar_target);
      __threadfence();
    }
    cluster.sync();
  };
  const int dim = a.dim, k = a.k, cols = 4 + 2 * k;
  for (int c = threadIdx.x; c < 2 * kMaxDim; c += kThreads) (&s_g0[0][0])[c] = 0.0f;
  __syncthreads();
  double loss_total = 0.0;
  const int64_t num_batches = (a.M + a.batch - 1) / a.batch;
  // the indices / targets of this warp's first sample of a batch are fetched one batch ahead
  // (during phase 2 of the previous one) and kept for phase 2: two dependent global loads
  // (order -> feature columns) less on the critical path of every phase
  auto fetch = [&](int64_t first, int bs, int s, int32_t* my, float* my_t) {
    *my = 0;
    *my_t = 0.0f;
    if (s < bs) {
      const int64_t sample = a.order[first + s];
      if (lane < cols) *my = a.feat[(size_t)lane * a.M + sample];
      if (lane < 3) *my_t = a.target[(size_t)lane * a.M + sample];
    }
  };
  int32_t my_first, my_next = 0;
  float my_t_first, my_t_next = 0.0f;
  fetch(0, (int)min((int64_t)a.batch, a.M), warp, &my_first, &my_t_first);
  for (int64_t b = 0; b < num_batches; ++b) {
    const int64_t first = b * a.batch;
    const int bs = (int)min((int64_t)a.batch, a.M - first);
    const float inv_bs = 1.0f / (float)bs;
    // ---- phase 1: forward + backward ----------------------------------------------------
    for (int s = warp; s < bs; s += num_warps) {
      int32_t my = my_first;
      float my_t = my_t_first;
      if (s != warp) fetch(first, bs, s, &my, &my_t);
      const int32_t ln = __shfl_sync(kFull, my, 0), le = __shfl_sync(kFull, my, 1);
      const int32_t rn = __shfl_sync(kFull, my, 2), re = __shfl_sync(kFull, my, 3);
      float Ln[DPL], Rn[DPL], Le[DPL], Re[DPL], X[DPL];
      load_row<DPL>(a.N, ln, dim, lane, Ln);
      load_row<DPL>(a.N, rn, dim, lane, Rn);
      load_row<DPL>(a.E, le, dim, lane, Le);
      load_row<DPL>(a.E, re, dim, lane, Re);
      const float z_nn = dot<DPL>(Ln, Rn), z_ee = dot<DPL>(Le, Re);
      const float p_nn = activate(z_nn, a.activation), p_ee = activate(z_ee, a.activation);
      // neighbourhood terms: lane i keeps the pre-activation of neighbour i of either kind.
      // With dim <= 32 and k <= 8 all 2k neighbour rows are loaded up front (one memory latency
      // instead of 2k dependent ones) and stay in registers for the gradient pass.
      constexpr int KM = 8;
      const bool preload = DPL == 1 && k <= KM;
      float nbN[KM], nbE[KM];
      float za_mine = 0.0f, zb_mine = 0.0f;
      if (preload) {
#pragma unroll
        for (int i = 0; i < KM; ++i) {
          const int32_t xn = __shfl_sync(kFull, my, (4 + i) & 31), xe = __shfl_sync(kFull, my, (4 + k + i) & 31);
          nbN[i] = (i < k && lane < dim) ? __ldcg(a.N + (size_t)xn * dim + lane) : 0.0f;
          nbE[i] = (i < k && lane < dim) ? __ldcg(a.E + (size_t)xe * dim + lane) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < KM; ++i) {
          if (i < k) {
            const float za = warp_sum(nbN[i] * Ln[0]), zb = warp_sum(nbE[i] * Re[0]);
            if (lane == i) {
              za_mine = za;
              zb_mine = zb;
            }
          }
        }
      } else {
        for (int i = 0; i < k; ++i) {
          load_row<DPL>(a.N, __shfl_sync(kFull, my, 4 + i), dim, lane, X);
          const float za = dot<DPL>(X, Ln);
          load_row<DPL>(a.E, __shfl_sync(kFull, my, 4 + k + i), dim, lane, X);
          const float zb = dot<DPL>(X, Re);
          if (lane == i) {
            za_mine = za;
            zb_mine = zb;
          }
        }
      }
      const float a_mine = lane < k ? activate(za_mine, a.activation) : 0.0f;
      const float b_mine = lane < k ? activate(zb_mine, a.activation) : 0.0f;
      const float inv_k = k ? 1.0f / (float)k : 0.0f;
      const float A = warp_sum(a_mine) * inv_k, B = warp_sum(b_mine) * inv_k;
      const float p_ne = A * B;
      float l_nn, l_ee, l_ne;
      const float g_nn = loss_grad(p_nn, __shfl_sync(kFull, my_t, 0), a.loss, &l_nn);
      const float g_ee = loss_grad(p_ee, __shfl_sync(kFull, my_t, 1), a.loss, &l_ee);
      const float g_ne = loss_grad(p_ne, __shfl_sync(kFull, my_t, 2), a.loss, &l_ne);
      if (lane == 0) loss_total += (double)l_nn + (double)l_ee + (double)l_ne;
      const float d_nn = g_nn * inv_bs * activate_grad(z_nn, p_nn, a.activation);
      const float d_ee = g_ee * inv_bs * activate_grad(z_ee, p_ee, a.activation);
      const float da_mine = g_ne * inv_bs * B * inv_k * activate_grad(za_mine, a_mine, a.activation);
      const float db_mine = g_ne * inv_bs * A * inv_k * activate_grad(zb_mine, b_mine, a.activation);
      // gradient of the left node / right edge row: the direct term plus every neighbour term
      float gLn[DPL], gRe[DPL];
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        gLn[j] = d_nn * Rn[j];
        gRe[j] = d_ee * Le[j];
      }
      if (preload) {
#pragma unroll
        for (int i = 0; i < KM; ++i) {
          if (i < k) {
            const float da = __shfl_sync(kFull, da_mine, i), db = __shfl_sync(kFull, db_mine, i);
            const int32_t xn = __shfl_sync(kFull, my, 4 + i), xe = __shfl_sync(kFull, my, 4 + k + i);
            if (da != 0.0f) {
              gLn[0] = fmaf(da, nbN[i], gLn[0]);
              add_grad<DPL>(a.gN, s_g0[0], xn, dim, lane, da, Ln);
            }
            if (db != 0.0f) {
              gRe[0] = fmaf(db, nbE[i], gRe[0]);
              add_grad<DPL>(a.gE, s_g0[1], xe, dim, lane, db, Re);
            }
          }
        }
      } else {
        for (int i = 0; i < k; ++i) {
          const float da = __shfl_sync(kFull, da_mine, i), db = __shfl_sync(kFull, db_mine, i);
          const int32_t xn = __shfl_sync(kFull, my, 4 + i), xe = __shfl_sync(kFull, my, 4 + k + i);
          if (da != 0.0f) {
            load_row<DPL>(a.N, xn, dim, lane, X);
#pragma unroll
            for (int j = 0; j < DPL; ++j) gLn[j] = fmaf(da, X[j], gLn[j]);
            add_grad<DPL>(a.gN, s_g0[0], xn, dim, lane, da, Ln);
          }
          if (db != 0.0f) {
            load_row<DPL>(a.E, xe, dim, lane, X);
#pragma unroll
            for (int j = 0; j < DPL; ++j) gRe[j] = fmaf(db, X[j], gRe[j]);
            add_grad<DPL>(a.gE, s_g0[1], xe, dim, lane, db, Re);
          }
        }
      }
      add_grad<DPL>(a.gN, s_g0[0], ln, dim, lane, 1.0f, gLn);
      add_grad<DPL>(a.gN, s_g0[0], rn, dim, lane, d_nn, Ln);
      add_grad<DPL>(a.gE, s_g0[1], re, dim, lane, 1.0f, gRe);
      add_grad<DPL>(a.gE, s_g0[1], le, dim, lane, d_ee, Re);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * dim; c += kThreads) {
      const int t = c / dim, cc = c - t * dim;
      const float v = s_g0[t][cc];
      if (v != 0.0f) atomicAdd((t ? a.gE : a.gN) + cc, v);
      s_g0[t][cc] = 0.0f;
    }
    batch_barrier();
    if (b + 1 < num_batches)
      fetch(first + a.batch, (int)min((int64_t)a.batch, a.M - first - a.batch), warp, &my_next,
            &my_t_next);
    // ---- phase 2: Adagrad on the rows this warp's samples touched ------------------------
    const int32_t stamp = a.batch_id0 + (int32_t)b;
    for (int s = warp; s < bs; s += num_warps) {
      int32_t my = my_first;
      if (s != warp) {
        float unused;
        fetch(first, bs, s, &my, &unused);
      }
      // column c of the sample indexes the node table for c in {0, 2, 4 .. 4+k-1}
      int32_t old = stamp;
      if (lane < cols) {
        const bool node = lane == 0 || lane == 2 || (lane >= 4 && lane < 4 + k);
        old = atomicExch((node ? a.claimN : a.claimE) + my, stamp);
      }
      // the rows this warp owns, four at a time: all loads of a group first, then the updates
      unsigned owned = __ballot_sync(kFull, lane < cols && old != stamp);
      while (owned) {
        int32_t rw[4] = {0, 0, 0, 0};
        unsigned is_node = 0;
        int nr = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (owned) {
            const int c = __ffs(owned) - 1;
            owned &= owned - 1;
            rw[j] = __shfl_sync(kFull, my, c);
            if (c == 0 || c == 2 || (c >= 4 && c < 4 + k)) is_node |= 1u << j;
            nr = j + 1;
          }
        }
        for (int c = lane; c < dim; c += 32) {
          float gv[4], av[4], pv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const size_t at = (size_t)rw[j] * dim + c;
            const bool node = (is_node >> j) & 1;
            gv[j] = j < nr ? __ldcg((node ? a.gN : a.gE) + at) : 0.0f;
            av[j] = j < nr ? __ldcg((node ? a.accN : a.accE) + at) : 0.0f;
            pv[j] = j < nr ? __ldcg((node ? a.N : a.E) + at) : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nr && gv[j] != 0.0f) {
              const size_t at = (size_t)rw[j] * dim + c;
              const bool node = (is_node >> j) & 1;
              const float acc = av[j] + gv[j] * gv[j];
              (node ? a.accN : a.accE)[at] = acc;
              (node ? a.N : a.E)[at] = pv[j] - a.lr * gv[j] / (sqrtf(acc) + a.eps);
              (node ? a.gN : a.gE)[at] = 0.0f;
            }
          }
        }
      }
    }
    batch_barrier();
    my_first = my_next;
    my_t_first = my_t_next;
  }
  if (lane == 0 && loss_total != 0.0) atomicAdd(a.loss_sum, loss_total);
}

template <typename T>
int upload(hge_ctx* ctx, T** dst, const T* src, size_t count, int mem) {
  HGE_TRY(hge_dev_alloc(ctx, dst, count));
  if (count)
    HGE_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T),
                             mem == HGE_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                             ctx->stream));
  return HGE_OK;
}

}  // namespace

extern "C" {

int hge_hg2v_destroy(hge_hg2v_model* m) {
  if (!m) return HGE_OK;
  const hge_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  hge_dev_free(ctx, m->N);
  hge_dev_free(ctx, m->E);
  hge_dev_free(ctx, m->accN);
  hge_dev_free(ctx, m->accE);
  hge_dev_free(ctx, m->gN);
  hge_dev_free(ctx, m->gE);
  hge_dev_free(ctx, m->claimN);
  hge_dev_free(ctx, m->claimE);
  hge_dev_free(ctx, m->feat);
  hge_dev_free(ctx, m->target);
  hge_dev_free(ctx, m->order);
  hge_dev_free(ctx, m->d_loss);
  hge_dev_free(ctx, m->grid_bar);
  delete m;
  return HGE_OK;
}

int hge_hg2v_create(hge_ctx* ctx, int32_t node_rows, int32_t edge_rows, int dim, int num_neighbors,
                    int activation, int loss, const float* node_init, const float* edge_init,
                    int mem, hge_hg2v_model** out) {
  HGE_REQUIRE(ctx && out && node_init && edge_init, "hge_hg2v_create: NULL argument");
  *out = nullptr;
  HGE_REQUIRE(node_rows >= 1 && edge_rows >= 1, "hge_hg2v_create: empty embedding table");
  HGE_REQUIRE(dim >= 1 && dim <= kMaxDim, "hge_hg2v_create: dimension %d not in [1, %d]", dim, kMaxDim);
  HGE_REQUIRE(num_neighbors >= 0 && 4 + 2 * num_neighbors <= 32,
              "hge_hg2v_create: num_neighbors %d not in [0, 14]", num_neighbors);
  HGE_REQUIRE((activation == 0 || activation == 1) && (loss == 0 || loss == 1),
              "hge_hg2v_create: activation / loss must be 0 or 1");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_hg2v_create: bad mem %d", mem);
  HGE_CUDA(cudaSetDevice(ctx->device));
  hge_hg2v_model* m = new (std::nothrow) hge_hg2v_model();
  if (!m) return HGE_ERR_NOMEM;
  m->ctx = ctx;
  m->node_rows = node_rows;
  m->edge_rows = edge_rows;
  m->dim = dim;
  m->k = num_neighbors;
  m->activation = activation;
  m->loss = loss;
  const size_t nn = (size_t)node_rows * dim, ne = (size_t)edge_rows * dim;
  int rc = upload(ctx, &m->N, node_init, nn, mem);
  if (rc == HGE_OK) rc = upload(ctx, &m->E, edge_init, ne, mem);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->accN, nn);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->accE, ne);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->gN, nn);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->gE, ne);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->claimN, (size_t)node_rows);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->claimE, (size_t)edge_rows);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->d_loss, 1);
  if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &m->grid_bar, 1);
  if (rc == HGE_OK) {
    cudaError_t e = cudaMemsetAsync(m->accN, 0, nn * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->accE, 0, ne * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->gN, 0, nn * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->gE, 0, ne * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->claimN, 0xff, (size_t)node_rows * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->claimE, 0xff, (size_t)edge_rows * 4, ctx->stream);
    if (e != cudaSuccess) {
      hge_set_error("hge_hg2v_create: memset failed: %s", cudaGetErrorString(e));
      rc = HGE_ERR_CUDA;
    }
  }
  if (rc != HGE_OK) {
    hge_hg2v_destroy(m);
    return rc;
  }
  *out = m;
  return HGE_OK;
}

int hge_hg2v_set_samples(hge_hg2v_model* m, const int32_t* features, const float* targets,
                         int64_t num_samples, int mem) {
  HGE_REQUIRE(m && num_samples >= 0 && (num_samples == 0 || (features && targets)),
              "hge_hg2v_set_samples: bad argument");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_hg2v_set_samples: bad mem %d", mem);
  hge_ctx* ctx = m->ctx;
  HGE_CUDA(cudaSetDevice(ctx->device));
  hge_dev_free(ctx, m->feat);
  hge_dev_free(ctx, m->target);
  hge_dev_free(ctx, m->order);
  m->M = num_samples;
  const size_t cols = 4 + 2 * (size_t)m->k;
  HGE_TRY(upload(ctx, &m->feat, features, cols * (size_t)num_samples, mem));
  HGE_TRY(upload(ctx, &m->target, targets, 3 * (size_t)num_samples, mem));
  HGE_TRY(hge_dev_alloc(ctx, &m->order, (size_t)num_samples));
  // every index must address a table row: check once here instead of in the hot loop
  if (mem == HGE_MEM_HOST) {
    for (size_t c = 0; c < cols; ++c) {
      const bool node = c == 0 || c == 2 || (c >= 4 && c < 4 + (size_t)m->k);
      const int32_t limit = node ? m->node_rows : m->edge_rows;
      const int32_t* col = features + c * (size_t)num_samples;
      for (int64_t i = 0; i < num_samples; ++i)
        HGE_REQUIRE(col[i] >= 0 && col[i] < limit,
                    "hge_hg2v_set_samples: feature column %zu, sample %lld: index %d outside its "
                    "table of %d rows", c, (long long)i, col[i], limit);
    }
  }
  return HGE_OK;
}

int hge_hg2v_fit_epoch(hge_hg2v_model* m, const int32_t* order, int batch_size, int mem,
                       double* epoch_loss) {
  HGE_REQUIRE(m && epoch_loss && batch_size >= 1, "hge_hg2v_fit_epoch: bad argument");
  HGE_REQUIRE(m->M > 0 && order, "hge_hg2v_fit_epoch: no samples (hge_hg2v_set_samples) or no order");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_hg2v_fit_epoch: bad mem %d", mem);
  hge_ctx* ctx = m->ctx;
  HGE_CUDA(cudaSetDevice(ctx->device));
  if (mem == HGE_MEM_HOST)
    for (int64_t i = 0; i < m->M; ++i)
      HGE_REQUIRE(order[i] >= 0 && order[i] < m->M, "hge_hg2v_fit_epoch: order[%lld] out of range",
                  (long long)i);
  HGE_CUDA(cudaMemcpyAsync(m->order, order, (size_t)m->M * 4,
                           mem == HGE_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                           ctx->stream));
  HGE_CUDA(cudaMemsetAsync(m->d_loss, 0, sizeof(double), ctx->stream));
  const int64_t num_batches = (m->M + batch_size - 1) / batch_size;
  HGE_REQUIRE((int64_t)m->next_batch_id + num_batches < INT32_MAX, "hge_hg2v_fit_epoch: batch counter overflow");
  TrainArgs a;
  a.N = m->N; a.E = m->E; a.accN = m->accN; a.accE = m->accE; a.gN = m->gN; a.gE = m->gE;
  a.claimN = m->claimN; a.claimE = m->claimE;
  a.feat = m->feat; a.target = m->target; a.order = m->order; a.loss_sum = m->d_loss;
  a.M = m->M; a.dim = m->dim; a.k = m->k; a.batch = batch_size;
  a.activation = m->activation; a.loss = m->loss; a.batch_id0 = m->next_batch_id;
  a.lr = 0.01f; a.eps = 1e-7f;   // keras.optimizers.Adagrad defaults
  m->next_batch_id += (int32_t)num_batches;
  const int dpl = (m->dim + 31) / 32;
  void (*kernel)(const TrainArgs) = dpl <= 1 ? k_hg2v_epoch<1> : dpl <= 2 ? k_hg2v_epoch<2>
                                    : dpl <= 4 ? k_hg2v_epoch<4> : k_hg2v_epoch<8>;
  // clusters that can be resident at once (they spin on each other): asked once per model
  if (m->max_clusters == 0) {
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3(kCluster * 32, 1, 1);
    probe.blockDim = dim3(kThreads, 1, 1);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &probe) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 1;
    }
    m->max_clusters = n;
  }
  const int want = (int)std::min<int64_t>((std::min<int64_t>(batch_size, m->M) + kCluster * kWarpsPerCta - 1) /
                                              (kCluster * kWarpsPerCta), m->max_clusters);
  int clusters = std::max(1, std::min(want, ctx->trainer_max_clusters > 0 ? ctx->trainer_max_clusters : want));
  a.grid_bar = m->grid_bar;
  for (;;) {
    a.clusters = clusters;
    cudaError_t e = cudaSuccess;
    if (clusters == 1) {
      kernel<<<kCluster, kThreads, 0, ctx->stream>>>(a);
      e = cudaGetLastError();
    } else {
      e = cudaMemsetAsync(m->grid_bar, 0, sizeof(unsigned int), ctx->stream);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kCluster * clusters, 1, 1);
      cfg.blockDim = dim3(kThreads, 1, 1);
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;   // every block resident, or the launch fails
      attr[0].val.cooperative = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, kernel, a);
    }
    if (e == cudaSuccess) break;
    cudaGetLastError();
    if (clusters == 1) {
      hge_set_error("hge_hg2v_fit_epoch: kernel launch failed: %s", cudaGetErrorString(e));
      return HGE_ERR_CUDA;
    }
    clusters = clusters > 2 ? clusters / 2 : 1;   // fewer clusters fit: same result, more rounds per batch
    m->max_clusters = clusters;
  }
  m->last_clusters = clusters;
  ctx->launches++;
  double total = 0.0;
  HGE_CUDA(cudaMemcpyAsync(&total, m->d_loss, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  *epoch_loss = total / (double)m->M;
  return HGE_OK;
}

int hge_hg2v_last_clusters(const hge_hg2v_model* m) { return m ? m->last_clusters : 0; }

int hge_hg2v_get_weights(hge_hg2v_model* m, float* node, float* edge, int mem) {
  HGE_REQUIRE(m && node && edge, "hge_hg2v_get_weights: NULL argument");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_hg2v_get_weights: bad mem %d", mem);
  hge_ctx* ctx = m->ctx;
  HGE_CUDA(cudaSetDevice(ctx->device));
  const cudaMemcpyKind kind = mem == HGE_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  HGE_CUDA(cudaMemcpyAsync(node, m->N, (size_t)m->node_rows * m->dim * 4, kind, ctx->stream));
  HGE_CUDA(cudaMemcpyAsync(edge, m->E, (size_t)m->edge_rows * m->dim * 4, kind, ctx->stream));
  HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  return HGE_OK;
}

}  // extern "C"
