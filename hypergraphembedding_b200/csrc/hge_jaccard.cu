// Sparse weighted Jaccard similarities of WeightedJaccardSamples (hg2v_sample.py:250-395) on
// sm_100a, in the reference's own summation order.
//
//   J(x, y) = sum_c min(x_c, y_c) / sum_c max(x_c, y_c)   over the union of the non-zeros,
//   0 when the denominator is 0 (SparseWeightedJaccard, :250-275).
//
// The reference walks the union of the two rows' columns in ascending order and adds the smaller
// value to the numerator and the larger to the denominator ONE BY ONE, in the dtype of the
// features (fp32 for every weighting scheme of hg2v_weighting.py).  On rows with thousands of
// non-zeros that sum is only good to a few 1e-6, so any other order (a warp reduction, the
// identity sum max = sum x + sum y - sum min) misses the 1e-5 bar on a fraction of the records
// even though it is closer to the exact value.  Here a pair is one thread that merges the two
// sorted rows and carries the two fp32 sums in that order: the result is the reference's, bit
// for bit; the parallelism is across the (hundreds of thousands of) sampled pairs.
//
//   same type   (node, node) / (edge, edge): x and y are rows of one feature matrix
//               (SameTypeJaccardSample, :323-340);
//   diff type   (node, edge): x is a feature row, y the CENTROID of a group of feature rows,
//               y = (sum_{t in group} F[t]) / |group|  (CentroidFromRows, :284-299,
//               DiffTypeJaccardSample, :343-392).  The reference materialises the centroid rows
//               with scipy (column sums of F[rows]: the members are added in ascending order,
//               fp32, then one division); so does this file, for the groups the pairs name:
//               expand (group, column, value) triples in member order, stable radix sort by
//               (group, column), one sequential sum per run -- the classic expand / sort /
//               compress product, batched so that the expansion stays inside a memory budget.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "hge_common.cuh"
#include "hge_staged.cuh"

namespace {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;
// entries of one expansion batch (12 bytes each, twice for the sort's double buffer)
constexpr int64_t kBatchEntries = 192ll << 20;

// one step of SparseWeightedJaccard's loop (hg2v_sample.py:264-271)
__device__ __forceinline__ void jaccard_step(float x, float y, float& num, float& den) {
  if (x < y) {
    num = __fadd_rn(num, x);
    den = __fadd_rn(den, y);
  } else {
    num = __fadd_rn(num, y);
    den = __fadd_rn(den, x);
  }
}

// out[p] = J(A[pa[p]], B[pb[p]]), one thread per pair, union walked in ascending column order.
// skip_empty_b: a pair whose B row is marked empty (bptr[r] < 0 never happens; see b_count) gives 0.
__global__ void __launch_bounds__(kBlock) k_jaccard_merge(
    int64_t num_pairs, const int64_t* __restrict__ aptr, const int32_t* __restrict__ aidx,
    const float* __restrict__ aval, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bidx,
    const float* __restrict__ bval, const int32_t* __restrict__ pa, const int32_t* __restrict__ pb,
    const int64_t* __restrict__ b_members, float* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < num_pairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t ra = pa[p], rb = pb[p];
    if (b_members && b_members[rb + 1] == b_members[rb]) {   // a group without members has no centroid
      out[p] = 0.f;
      continue;
    }
    int64_t i = aptr[ra], j = bptr[rb];
    const int64_t ie = aptr[ra + 1], je = bptr[rb + 1];
    float num = 0.f, den = 0.f;
    int32_t ci = i < ie ? aidx[i] : INT32_MAX, cj = j < je ? bidx[j] : INT32_MAX;
    while (i < ie || j < je) {
      float x = 0.f, y = 0.f;
      const bool take_a = ci <= cj, take_b = cj <= ci;
      if (take_a) {
        x = aval[i];
        ++i;
      }
      if (take_b) {
        y = bval[j];
        ++j;
      }
      if (take_a) ci = i < ie ? aidx[i] : INT32_MAX;
      if (take_b) cj = j < je ? bidx[j] : INT32_MAX;
      jaccard_step(x, y, num, den);
    }
    out[p] = den == 0.f ? 0.f : __fdiv_rn(num, den);
  }
}

// ---- centroid rows -----------------------------------------------------------------------------

__global__ void k_flag_groups(int64_t num_pairs, const int32_t* __restrict__ pg, int32_t* __restrict__ flag) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < num_pairs;
       p += (int64_t)gridDim.x * blockDim.x)
    flag[pg[p]] = 1;
}

// w[t] = non-zeros of F[gidx[t]] when t's group is named by a pair, else 0 (one warp per group)
__global__ void __launch_bounds__(kBlock) k_member_weights(
    int64_t groups, const int64_t* __restrict__ gptr, const int32_t* __restrict__ gidx,
    const int32_t* __restrict__ flag, const int64_t* __restrict__ fptr, int64_t* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t g = blockIdx.x * (int64_t)(kBlock / 32) + (threadIdx.x >> 5); g < groups; g += nw) {
    const bool on = flag[g] != 0;
    for (int64_t t = gptr[g] + lane; t < gptr[g + 1]; t += 32) {
      const int32_t m = gidx[t];
      w[t] = on ? fptr[m + 1] - fptr[m] : 0;
    }
  }
}

// goff[g] = off[gptr[g]]: first expanded entry of group g (goff[groups] = total)
__global__ void k_group_offsets(int64_t groups, int64_t gnnz, int64_t total, const int64_t* __restrict__ gptr,
                                const int64_t* __restrict__ off, int64_t* __restrict__ goff) {
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g <= groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = gptr[g];
    goff[g] = t < gnnz ? off[t] : total;
  }
}

// expansion of the member positions [t0, t1) (groups [g0, g1)): one warp per member position,
// entries in member order; key = (group - g0) << 32 | column
__global__ void __launch_bounds__(kBlock) k_expand(
    int64_t t0, int64_t t1, int64_t g0, int64_t g1, int64_t base, const int64_t* __restrict__ gptr,
    const int32_t* __restrict__ gidx, const int64_t* __restrict__ off, const int64_t* __restrict__ w,
    const int64_t* __restrict__ fptr, const int32_t* __restrict__ fidx, const float* __restrict__ fval,
    uint64_t* __restrict__ keys, float* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t t = t0 + blockIdx.x * (int64_t)(kBlock / 32) + (threadIdx.x >> 5); t < t1; t += nw) {
    const int64_t n = w[t];
    if (n == 0) continue;
    // group of t: last g in [g0, g1) with gptr[g] <= t
    int64_t lo = g0, hi = g1;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (gptr[mid] <= t) lo = mid; else hi = mid;
    }
    const uint64_t gkey = (uint64_t)(lo - g0) << 32;
    const int64_t src = fptr[gidx[t]], dst = off[t] - base;
    for (int64_t k = lane; k < n; k += 32) {
      keys[dst + k] = gkey | (uint32_t)fidx[src + k];
      vals[dst + k] = fval[src + k];
    }
  }
}

__global__ void k_run_heads(int64_t n, const uint64_t* __restrict__ keys, int64_t* __restrict__ head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// head of a run: the members' values added in member order, then one division by the group size
// (targets2features[rows].sum(axis=0) / len(rows), hg2v_sample.py:293)
__global__ void k_run_sums(int64_t n, int64_t g0, const uint64_t* __restrict__ keys,
                           const float* __restrict__ vals, const int64_t* __restrict__ pos,
                           const int64_t* __restrict__ gptr, int32_t* __restrict__ cidx,
                           float* __restrict__ cval) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    if (i > 0 && keys[i - 1] == key) continue;
    float s = 0.f;
    for (int64_t k = i; k < n && keys[k] == key; ++k) s = __fadd_rn(s, vals[k]);
    const int64_t g = g0 + (int64_t)(key >> 32);
    const int64_t o = pos[i];
    cidx[o] = (int32_t)(uint32_t)key;
    cval[o] = __fdiv_rn(s, (float)(gptr[g + 1] - gptr[g]));
  }
}

// cptr[g] for g in [g0, g1]: compressed position of the first run of group g (+ rows before this batch)
__global__ void k_centroid_ptr(int64_t g0, int64_t g1, int64_t n, int64_t n_runs, int64_t rows_before,
                               const uint64_t* __restrict__ keys, const int64_t* __restrict__ pos,
                               int64_t* __restrict__ cptr, bool last_batch) {
  const int64_t count = g1 - g0 + (last_batch ? 1 : 0);
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count;
       k += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t want = (uint64_t)k << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    cptr[g0 + k] = rows_before + (lo < n ? pos[lo] : n_runs);
  }
}

int grid_for(const hge_ctx* ctx, int64_t work, int per_block = kBlock) {
  const int64_t want = std::max<int64_t>(1, (work + per_block - 1) / per_block);
  return (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * 16);
}

struct FeatureMatrix {
  Staged<int64_t> ptr;
  Staged<int32_t> idx;
  Staged<float> val;
  int init(hge_ctx* c, const int64_t* p, const int32_t* i, const float* v, int64_t rows, int64_t nnz, int mem) {
    HGE_TRY(ptr.init(c, p, (size_t)rows + 1, mem, true, false));
    HGE_TRY(idx.init(c, i, (size_t)nnz, mem, true, false));
    HGE_TRY(val.init(c, v, (size_t)nnz, mem, true, false));
    return HGE_OK;
  }
};

template <typename T>
int read_back(hge_ctx* ctx, const T* dev, T* host, size_t n) {
  HGE_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  return HGE_OK;
}

// Centroid rows of the groups named by pg[0 .. num_pairs): CSR over ALL groups (rows of the
// groups no pair names are empty).  Everything on ctx->stream; the arrays are the caller's to free.
struct CentroidRows {
  int64_t* ptr = nullptr;
  int32_t* idx = nullptr;
  float* val = nullptr;
};

int build_centroids(hge_ctx* ctx, const int64_t* gptr, const int32_t* gidx, int64_t groups, int64_t gnnz,
                    const FeatureMatrix& f, const int32_t* pg, int64_t num_pairs, CentroidRows* out) {
  int32_t* flag = nullptr;
  int64_t *w = nullptr, *off = nullptr, *goff = nullptr;
  HGE_TRY(hge_dev_alloc(ctx, &flag, (size_t)groups));
  HGE_TRY(hge_dev_alloc(ctx, &w, (size_t)gnnz + 1));
  HGE_TRY(hge_dev_alloc(ctx, &off, (size_t)gnnz + 1));
  HGE_TRY(hge_dev_alloc(ctx, &goff, (size_t)groups + 1));
  HGE_TRY(hge_dev_alloc(ctx, &out->ptr, (size_t)groups + 1));
  HGE_CUDA(cudaMemsetAsync(flag, 0, (size_t)groups * sizeof(int32_t), ctx->stream));
  HGE_CUDA(cudaMemsetAsync(w, 0, ((size_t)gnnz + 1) * sizeof(int64_t), ctx->stream));
  k_flag_groups<<<grid_for(ctx, num_pairs), kBlock, 0, ctx->stream>>>(num_pairs, pg, flag);
  HGE_CHECK_LAUNCH(ctx);
  k_member_weights<<<grid_for(ctx, groups, kBlock / 32), kBlock, 0, ctx->stream>>>(groups, gptr, gidx, flag,
                                                                                 f.ptr.dev, w);
  HGE_CHECK_LAUNCH(ctx);
  size_t scan_bytes = 0;
  HGE_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, w, off, gnnz + 1, ctx->stream));
  char* scan_temp = nullptr;
  HGE_TRY(hge_dev_alloc(ctx, &scan_temp, scan_bytes));
  HGE_CUDA(cub::DeviceScan::ExclusiveSum(scan_temp, scan_bytes, w, off, gnnz + 1, ctx->stream));
  ctx->launches++;
  hge_dev_free(ctx, scan_temp);
  int64_t total = 0;
  HGE_TRY(read_back(ctx, off + gnnz, &total, 1));
  k_group_offsets<<<grid_for(ctx, groups + 1), kBlock, 0, ctx->stream>>>(groups, gnnz, total, gptr, off, goff);
  HGE_CHECK_LAUNCH(ctx);
  std::vector<int64_t> h_goff((size_t)groups + 1), h_gptr((size_t)groups + 1);
  HGE_TRY(read_back(ctx, goff, h_goff.data(), (size_t)groups + 1));
  HGE_TRY(read_back(ctx, gptr, h_gptr.data(), (size_t)groups + 1));

  // batches of consecutive groups whose expansion fits the budget (a single group larger than
  // the budget is its own batch)
  struct Part {
    int32_t* idx;
    float* val;
    int64_t n;
  };
  std::vector<Part> parts;
  int64_t rows_before = 0;
  int rc = HGE_OK;
  for (int64_t g0 = 0; g0 < groups && rc == HGE_OK;) {
    int64_t g1 = g0 + 1;
    while (g1 < groups && h_goff[(size_t)g1 + 1] - h_goff[(size_t)g0] <= kBatchEntries && g1 - g0 < (1ll << 31) - 1)
      ++g1;
    const int64_t n = h_goff[(size_t)g1] - h_goff[(size_t)g0];
    const bool last = g1 == groups;
    if (n == 0) {
      // no pair names any of these groups: empty rows
      k_centroid_ptr<<<grid_for(ctx, g1 - g0 + 1), kBlock, 0, ctx->stream>>>(g0, g1, 0, 0, rows_before, nullptr,
                                                                           nullptr, out->ptr, last);
      HGE_CHECK_LAUNCH(ctx);
      g0 = g1;
      continue;
    }
    uint64_t *keys = nullptr, *keys2 = nullptr;
    float *vals = nullptr, *vals2 = nullptr;
    int64_t *head = nullptr, *pos = nullptr;
    char* temp = nullptr;
    Part part = {nullptr, nullptr, 0};
    auto batch = [&]() -> int {
      HGE_TRY(hge_dev_alloc(ctx, &keys, (size_t)n));
      HGE_TRY(hge_dev_alloc(ctx, &keys2, (size_t)n));
      HGE_TRY(hge_dev_alloc(ctx, &vals, (size_t)n));
      HGE_TRY(hge_dev_alloc(ctx, &vals2, (size_t)n));
      k_expand<<<grid_for(ctx, h_gptr[(size_t)g1] - h_gptr[(size_t)g0], kBlock / 32), kBlock, 0, ctx->stream>>>(
          h_gptr[(size_t)g0], h_gptr[(size_t)g1], g0, g1, h_goff[(size_t)g0], gptr, gidx, off, w, f.ptr.dev,
          f.idx.dev, f.val.dev, keys, vals);
      HGE_CHECK_LAUNCH(ctx);
      int group_bits = 1;
      while ((1ll << group_bits) < g1 - g0) ++group_bits;
      size_t bytes = 0;
      HGE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys2, vals, vals2, n, 0, 32 + group_bits,
                                               ctx->stream));
      HGE_TRY(hge_dev_alloc(ctx, &temp, bytes));
      // radix sort is stable: inside a (group, column) run the members stay in ascending order
      HGE_CUDA(cub::DeviceRadixSort::SortPairs(temp, bytes, keys, keys2, vals, vals2, n, 0, 32 + group_bits,
                                               ctx->stream));
      ctx->launches++;
      hge_dev_free(ctx, temp);
      HGE_TRY(hge_dev_alloc(ctx, &head, (size_t)n + 1));
      HGE_TRY(hge_dev_alloc(ctx, &pos, (size_t)n + 1));
      k_run_heads<<<grid_for(ctx, n), kBlock, 0, ctx->stream>>>(n, keys2, head);
      HGE_CHECK_LAUNCH(ctx);
      HGE_CUDA(cudaMemsetAsync(head + n, 0, sizeof(int64_t), ctx->stream));
      HGE_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, head, pos, n + 1, ctx->stream));
      HGE_TRY(hge_dev_alloc(ctx, &temp, bytes));
      HGE_CUDA(cub::DeviceScan::ExclusiveSum(temp, bytes, head, pos, n + 1, ctx->stream));
      ctx->launches++;
      HGE_TRY(read_back(ctx, pos + n, &part.n, 1));
      HGE_TRY(hge_dev_alloc(ctx, &part.idx, (size_t)std::max<int64_t>(1, part.n)));
      HGE_TRY(hge_dev_alloc(ctx, &part.val, (size_t)std::max<int64_t>(1, part.n)));
      k_run_sums<<<grid_for(ctx, n), kBlock, 0, ctx->stream>>>(n, g0, keys2, vals2, pos, gptr, part.idx, part.val);
      HGE_CHECK_LAUNCH(ctx);
      k_centroid_ptr<<<grid_for(ctx, g1 - g0 + 1), kBlock, 0, ctx->stream>>>(g0, g1, n, part.n, rows_before, keys2,
                                                                           pos, out->ptr, last);
      HGE_CHECK_LAUNCH(ctx);
      return HGE_OK;
    };
    rc = batch();
    hge_dev_free(ctx, keys);
    hge_dev_free(ctx, keys2);
    hge_dev_free(ctx, vals);
    hge_dev_free(ctx, vals2);
    hge_dev_free(ctx, head);
    hge_dev_free(ctx, pos);
    hge_dev_free(ctx, temp);
    if (rc == HGE_OK) {
      parts.push_back(part);
      rows_before += part.n;
    } else {
      hge_dev_free(ctx, part.idx);
      hge_dev_free(ctx, part.val);
    }
    g0 = g1;
  }
  hge_dev_free(ctx, flag);
  hge_dev_free(ctx, w);
  hge_dev_free(ctx, off);
  hge_dev_free(ctx, goff);
  if (rc == HGE_OK) {
    if (parts.size() == 1) {
      out->idx = parts[0].idx;
      out->val = parts[0].val;
      parts.clear();
    } else {
      rc = hge_dev_alloc(ctx, &out->idx, (size_t)std::max<int64_t>(1, rows_before));
      if (rc == HGE_OK) rc = hge_dev_alloc(ctx, &out->val, (size_t)std::max<int64_t>(1, rows_before));
      int64_t at = 0;
      for (const Part& p : parts) {
        if (rc == HGE_OK && p.n) {
          cudaMemcpyAsync(out->idx + at, p.idx, (size_t)p.n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream);
          cudaMemcpyAsync(out->val + at, p.val, (size_t)p.n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        }
        at += p.n;
      }
    }
  }
  for (Part& p : parts) {
    hge_dev_free(ctx, p.idx);
    hge_dev_free(ctx, p.val);
  }
  if (rc != HGE_OK) {
    hge_dev_free(ctx, out->ptr);
    hge_dev_free(ctx, out->idx);
    hge_dev_free(ctx, out->val);
  }
  return rc;
}

}  // namespace

extern "C" {

int hge_jaccard_rows(hge_ctx* ctx, const int64_t* ptr, const int32_t* idx, const float* val,
                     int64_t rows, int64_t nnz, const int32_t* pi, const int32_t* pj,
                     int64_t num_pairs, float* out, int mem) {
  HGE_REQUIRE(ctx && ptr && (idx || nnz == 0) && (val || nnz == 0) && rows >= 0 && nnz >= 0,
              "hge_jaccard_rows: bad feature matrix");
  HGE_REQUIRE(num_pairs >= 0 && (num_pairs == 0 || (pi && pj && out)), "hge_jaccard_rows: bad pairs");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_jaccard_rows: bad mem %d", mem);
  HGE_CUDA(cudaSetDevice(ctx->device));
  if (num_pairs == 0) return HGE_OK;
  FeatureMatrix f;
  Staged<int32_t> s_i, s_j;
  Staged<float> s_out;
  HGE_TRY(f.init(ctx, ptr, idx, val, rows, nnz, mem));
  HGE_TRY(s_i.init(ctx, pi, (size_t)num_pairs, mem, true, false));
  HGE_TRY(s_j.init(ctx, pj, (size_t)num_pairs, mem, true, false));
  HGE_TRY(s_out.init(ctx, out, (size_t)num_pairs, mem, false, true));
  k_jaccard_merge<<<grid_for(ctx, num_pairs), kBlock, 0, ctx->stream>>>(
      num_pairs, f.ptr.dev, f.idx.dev, f.val.dev, f.ptr.dev, f.idx.dev, f.val.dev, s_i.dev, s_j.dev, nullptr,
      s_out.dev);
  HGE_CHECK_LAUNCH(ctx);
  return s_out.finish();
}

int hge_jaccard_centroid(hge_ctx* ctx, const int64_t* xptr, const int32_t* xidx, const float* xval,
                         int64_t xrows, int64_t xnnz, const int64_t* gptr, const int32_t* gidx,
                         int64_t grows, int64_t gnnz, const int64_t* fptr, const int32_t* fidx,
                         const float* fval, int64_t frows, int64_t fnnz, const int32_t* px,
                         const int32_t* pg, int64_t num_pairs, float* out, int mem) {
  HGE_REQUIRE(ctx && xptr && gptr && fptr && xrows >= 0 && grows >= 0 && frows >= 0 && xnnz >= 0 &&
                  gnnz >= 0 && fnnz >= 0, "hge_jaccard_centroid: bad matrix");
  HGE_REQUIRE(num_pairs >= 0 && (num_pairs == 0 || (px && pg && out)), "hge_jaccard_centroid: bad pairs");
  HGE_REQUIRE(mem == HGE_MEM_HOST || mem == HGE_MEM_DEVICE, "hge_jaccard_centroid: bad mem %d", mem);
  HGE_CUDA(cudaSetDevice(ctx->device));
  if (num_pairs == 0) return HGE_OK;
  FeatureMatrix x, f;
  const bool same = xptr == fptr && xidx == fidx && xval == fval;   // usual case: X is F
  Staged<int64_t> s_gptr;
  Staged<int32_t> s_gidx, s_px, s_pg;
  Staged<float> s_out;
  HGE_TRY(f.init(ctx, fptr, fidx, fval, frows, fnnz, mem));
  if (!same) HGE_TRY(x.init(ctx, xptr, xidx, xval, xrows, xnnz, mem));
  HGE_TRY(s_gptr.init(ctx, gptr, (size_t)grows + 1, mem, true, false));
  HGE_TRY(s_gidx.init(ctx, gidx, (size_t)gnnz, mem, true, false));
  HGE_TRY(s_px.init(ctx, px, (size_t)num_pairs, mem, true, false));
  HGE_TRY(s_pg.init(ctx, pg, (size_t)num_pairs, mem, true, false));
  HGE_TRY(s_out.init(ctx, out, (size_t)num_pairs, mem, false, true));
  CentroidRows c;
  HGE_TRY(build_centroids(ctx, s_gptr.dev, s_gidx.dev, grows, gnnz, f, s_pg.dev, num_pairs, &c));
  const FeatureMatrix& xr = same ? f : x;
  k_jaccard_merge<<<grid_for(ctx, num_pairs), kBlock, 0, ctx->stream>>>(
      num_pairs, xr.ptr.dev, xr.idx.dev, xr.val.dev, c.ptr, c.idx, c.val, s_px.dev, s_pg.dev, s_gptr.dev,
      s_out.dev);
  ctx->launches++;
  const bool launched = cudaGetLastError() == cudaSuccess;
  hge_dev_free(ctx, c.ptr);
  hge_dev_free(ctx, c.idx);
  hge_dev_free(ctx, c.val);
  if (!launched) {
    hge_set_error("hge_jaccard_centroid: kernel launch failed");
    return HGE_ERR_CUDA;
  }
  return s_out.finish();
}

}  // extern "C"
