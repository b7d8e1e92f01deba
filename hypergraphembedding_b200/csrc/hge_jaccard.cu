// Sparse weighted Jaccard similarities of WeightedJaccardSamples (hg2v_sample.py:250-395) on
// sm_100a.
//
//   J(x, y) = sum_c min(x_c, y_c) / sum_c max(x_c, y_c)   over the union of the non-zeros,
//   0 when the denominator is 0 (SparseWeightedJaccard, :250-275).
//
// Feature rows are CSR rows with fp32 values and sorted column ids.  For non-negative features
// (the only kind the reference's weighting schemes produce; the entry points check it)
// sum max = sum x + sum y - sum min, and min(x_c, y_c) is non-zero only on the intersection, so
// a pair costs one pass over the shorter row with a binary search in the longer one.
//
//   same type   (node, node) / (edge, edge): x and y are rows of one feature matrix
//               (SameTypeJaccardSample, :323-340);
//   diff type   (node, edge): x is a feature row, y the CENTROID of a group of feature rows,
//               y = mean_{t in group} F[t]  (CentroidFromRows, :284-299, DiffTypeJaccardSample,
//               :343-392).  The reference materialises every centroid as a sparse matrix with
//               the pattern of A * A^T (5.7 M entries on the 4.5 K-incidence fixture); here
//               y_c is evaluated on the fly for the columns of x only -- members of the group
//               in ascending order, fp32, which is also scipy's accumulation order -- and
//               sum y comes from the row sums of F.
// One warp per pair.
#include <algorithm>

#include "hge_common.cuh"
#include "hge_staged.cuh"

namespace {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

// value of column c in the sorted row [b, e) of (idx, val), 0 when absent
__device__ __forceinline__ float lookup(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                        int64_t b, int64_t e, int32_t c) {
  int64_t lo = b, hi = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(idx + mid) < c) lo = mid + 1; else hi = mid;
  }
  return (lo < e && __ldg(idx + lo) == c) ? __ldg(val + lo) : 0.f;
}

// rowsum[r] = sum of the row's values; *negative is set when any value is < 0
__global__ void __launch_bounds__(kBlock) k_row_sums(int64_t rows, const int64_t* __restrict__ ptr,
                                                     const float* __restrict__ val,
                                                     float* __restrict__ rowsum, int* negative) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (kBlock / 32);
  bool neg = false;
  for (int64_t r = gw; r < rows; r += nw) {
    float s = 0.f;
    for (int64_t p = ptr[r] + lane; p < ptr[r + 1]; p += 32) {
      const float v = val[p];
      neg |= v < 0.f;
      s += v;
    }
    s = warp_sum(s);
    if (lane == 0) rowsum[r] = s;
  }
  if (__any_sync(kFull, neg) && lane == 0) *negative = 1;
}

__global__ void __launch_bounds__(kBlock) k_jaccard_rows(
    int64_t num_pairs, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
    const float* __restrict__ val, const float* __restrict__ rowsum, const int32_t* __restrict__ pi,
    const int32_t* __restrict__ pj, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t p = gw; p < num_pairs; p += nw) {
    int32_t i = pi[p], j = pj[p];
    if (ptr[i + 1] - ptr[i] > ptr[j + 1] - ptr[j]) {   // walk the shorter row
      const int32_t t = i;
      i = j;
      j = t;
    }
    const int64_t bj = ptr[j], ej = ptr[j + 1];
    float smin = 0.f;
    for (int64_t q = ptr[i] + lane; q < ptr[i + 1]; q += 32)
      smin += fminf(__ldg(val + q), lookup(idx, val, bj, ej, __ldg(idx + q)));
    smin = warp_sum(smin);
    if (lane == 0) {
      const float den = rowsum[i] + rowsum[j] - smin;
      out[p] = den == 0.f ? 0.f : smin / den;
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_jaccard_centroid(
    int64_t num_pairs, const int64_t* __restrict__ xptr, const int32_t* __restrict__ xidx,
    const float* __restrict__ xval, const float* __restrict__ xsum,
    const int64_t* __restrict__ gptr, const int32_t* __restrict__ gidx,
    const int64_t* __restrict__ fptr, const int32_t* __restrict__ fidx,
    const float* __restrict__ fval, const float* __restrict__ fsum,
    const int32_t* __restrict__ px, const int32_t* __restrict__ pg, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t p = gw; p < num_pairs; p += nw) {
    const int32_t r = px[p], g = pg[p];
    const int64_t gb = gptr[g], ge = gptr[g + 1];
    const float count = (float)(ge - gb);
    float sy = 0.f;   // sum of the centroid = mean of the members' row sums
    for (int64_t t = gb + lane; t < ge; t += 32) sy += fsum[gidx[t]];
    sy = warp_sum(sy) / count;
    float smin = 0.f;
    for (int64_t q = xptr[r] + lane; q < xptr[r + 1]; q += 32) {
      const int32_t c = __ldg(xidx + q);
      float y = 0.f;   // members in ascending order, fp32: scipy's own accumulation order
      for (int64_t t = gb; t < ge; ++t) {
        const int32_t m = __ldg(gidx + t);
        y += lookup(fidx, fval, fptr[m], fptr[m + 1], c);
      }
      smin += fminf(__ldg(xval + q), y / count);
    }
    smin = warp_sum(smin);
    if (lane == 0) {
      // an empty group has no centroid (the reference divides by len(rows) == 0)
      const float den = xsum[r] + sy - smin;
      out[p] = (ge == gb || den == 0.f) ? 0.f : smin / den;
    }
  }
}

int grid_for(const hge_ctx* ctx, int64_t warps) {
  const int64_t want = std::max<int64_t>(1, (warps + kBlock / 32 - 1) / (kBlock / 32));
  return (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * 16);
}

struct FeatureMatrix {
  Staged<int64_t> ptr;
  Staged<int32_t> idx;
  Staged<float> val;
  float* rowsum = nullptr;
  const hge_ctx* ctx = nullptr;

  int init(hge_ctx* c, const int64_t* p, const int32_t* i, const float* v, int64_t rows, int64_t nnz,
           int mem, int* d_negative) {
    ctx = c;
    HGE_TRY(ptr.init(c, p, (size_t)rows + 1, mem, true, false));
    HGE_TRY(idx.init(c, i, (size_t)nnz, mem, true, false));
    HGE_TRY(val.init(c, v, (size_t)nnz, mem, true, false));
    HGE_TRY(hge_dev_alloc(c, &rowsum, (size_t)rows));
    k_row_sums<<<grid_for(c, rows), kBlock, 0, c->stream>>>(rows, ptr.dev, val.dev, rowsum, d_negative);
    HGE_CHECK_LAUNCH(c);
    return HGE_OK;
  }
  ~FeatureMatrix() {
    if (ctx) hge_dev_free(ctx, rowsum);
  }
};

int check_non_negative(hge_ctx* ctx, int* d_negative, const char* fn) {
  int negative = 0;
  HGE_CUDA(cudaMemcpyAsync(&negative, d_negative, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (negative) {
    hge_set_error("%s: negative feature values are not supported (sum max = sum x + sum y - sum min "
                  "needs non-negative features)", fn);
    return HGE_ERR_UNSUPPORTED;
  }
  return HGE_OK;
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
  int* d_negative = nullptr;
  HGE_TRY(hge_dev_alloc(ctx, &d_negative, 1));
  HGE_CUDA(cudaMemsetAsync(d_negative, 0, sizeof(int), ctx->stream));
  int rc;
  {
    FeatureMatrix f;
    Staged<int32_t> s_i, s_j;
    Staged<float> s_out;
    rc = f.init(ctx, ptr, idx, val, rows, nnz, mem, d_negative);
    if (rc == HGE_OK) rc = s_i.init(ctx, pi, (size_t)num_pairs, mem, true, false);
    if (rc == HGE_OK) rc = s_j.init(ctx, pj, (size_t)num_pairs, mem, true, false);
    if (rc == HGE_OK) rc = s_out.init(ctx, out, (size_t)num_pairs, mem, false, true);
    if (rc == HGE_OK) {
      k_jaccard_rows<<<grid_for(ctx, num_pairs), kBlock, 0, ctx->stream>>>(
          num_pairs, f.ptr.dev, f.idx.dev, f.val.dev, f.rowsum, s_i.dev, s_j.dev, s_out.dev);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) {
        hge_set_error("hge_jaccard_rows: kernel launch failed");
        rc = HGE_ERR_CUDA;
      }
    }
    if (rc == HGE_OK) rc = check_non_negative(ctx, d_negative, "hge_jaccard_rows");
    if (rc == HGE_OK) rc = s_out.finish();
  }
  hge_dev_free(ctx, d_negative);
  return rc;
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
  int* d_negative = nullptr;
  HGE_TRY(hge_dev_alloc(ctx, &d_negative, 1));
  HGE_CUDA(cudaMemsetAsync(d_negative, 0, sizeof(int), ctx->stream));
  int rc;
  {
    FeatureMatrix x, f;
    const bool same = xptr == fptr && xidx == fidx && xval == fval;   // usual case: X is F
    Staged<int64_t> s_gptr;
    Staged<int32_t> s_gidx, s_px, s_pg;
    Staged<float> s_out;
    rc = f.init(ctx, fptr, fidx, fval, frows, fnnz, mem, d_negative);
    if (rc == HGE_OK && !same) rc = x.init(ctx, xptr, xidx, xval, xrows, xnnz, mem, d_negative);
    if (rc == HGE_OK) rc = s_gptr.init(ctx, gptr, (size_t)grows + 1, mem, true, false);
    if (rc == HGE_OK) rc = s_gidx.init(ctx, gidx, (size_t)gnnz, mem, true, false);
    if (rc == HGE_OK) rc = s_px.init(ctx, px, (size_t)num_pairs, mem, true, false);
    if (rc == HGE_OK) rc = s_pg.init(ctx, pg, (size_t)num_pairs, mem, true, false);
    if (rc == HGE_OK) rc = s_out.init(ctx, out, (size_t)num_pairs, mem, false, true);
    if (rc == HGE_OK) {
      const FeatureMatrix& xr = same ? f : x;
      k_jaccard_centroid<<<grid_for(ctx, num_pairs), kBlock, 0, ctx->stream>>>(
          num_pairs, xr.ptr.dev, xr.idx.dev, xr.val.dev, xr.rowsum, s_gptr.dev, s_gidx.dev, f.ptr.dev,
          f.idx.dev, f.val.dev, f.rowsum, s_px.dev, s_pg.dev, s_out.dev);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) {
        hge_set_error("hge_jaccard_centroid: kernel launch failed");
        rc = HGE_ERR_CUDA;
      }
    }
    if (rc == HGE_OK) rc = check_non_negative(ctx, d_negative, "hge_jaccard_centroid");
    if (rc == HGE_OK) rc = s_out.finish();
  }
  hge_dev_free(ctx, d_negative);
  return rc;
}

}  // extern "C"
