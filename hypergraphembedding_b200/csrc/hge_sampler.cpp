// Bit-exact candidate rows and sample draws (hg2v_sample.py:49-86) -- host code.
//
// The reference draws every sample from numpy's process-global legacy MT19937, one row after
// the other, with data-dependent rejection; the stream is inherently sequential, so this part
// of the path runs on the host while the probabilities of the
// drawn pairs are computed on the GPU (hge_weighting.cu).  Two third-party behaviours are
// restated here and pinned by tests/test_sampler_host.py against numpy / scipy themselves:
//   * scipy's csr_matmat emits each product row in reverse first-discovery order;
//   * numpy's legacy RandomState: choice(replace=False) == permutation(n)[:k] by Fisher-Yates
//     with rk_interval (masked rejection on 32-bit draws), choice(replace=True) and
//     randint == rk_interval(n - 1) per draw (no draw at all when n == 1).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "hge_common.cuh"

namespace {

// ---- numpy legacy MT19937 -------------------------------------------------------------------
struct Mt19937 {
  uint32_t* key;   // 624 words, borrowed from the caller's state buffer
  int pos;

  void regenerate() {
    const uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; ++i) {
      y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    for (; i < 623; ++i) {
      y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    y = (key[623] & kUpper) | (key[0] & kLower);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    pos = 0;
  }

  inline uint32_t next32() {
    if (pos == 624) regenerate();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }

  // rk_interval: uniform in [0, mx]; mx == 0 consumes nothing.
  inline uint32_t interval(uint32_t mx) {
    if (mx == 0) return 0;
    uint32_t mask = mx;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    do {
      v = next32() & mask;
    } while (v > mx);
    return v;
  }
};

// state625 = key[624] followed by pos
struct StateGuard {
  Mt19937 mt;
  uint32_t* buf;
  explicit StateGuard(uint32_t* state625) : buf(state625) {
    mt.key = state625;
    mt.pos = (int)state625[624];
  }
  ~StateGuard() { buf[624] = (uint32_t)mt.pos; }
};

// ---- candidate rows ---------------------------------------------------------------------------
struct Csr {
  const int64_t* ptr;
  const int32_t* idx;
};

// Row `r` of M1 (kind 0), of M1*M2 (kind 1) or of (M1*M2)*M3 (kind 2) in the order scipy stores
// it: kind 0 as stored; products in reverse first-discovery order (the left row is walked in its
// own stored order, each right row ascending).
class RowBuilder {
 public:
  RowBuilder(int kind, Csr m1, Csr m2, Csr m3, int32_t mid_cols, int32_t out_cols)
      : kind_(kind), m1_(m1), m2_(m2), m3_(m3) {
    if (kind >= 1) stamp_a_.assign((size_t)(kind == 1 ? out_cols : mid_cols), 0);
    if (kind == 2) stamp_b_.assign((size_t)out_cols, 0);
  }

  // Returns the candidate list (valid until the next call).
  const std::vector<int32_t>& row(int32_t r) {
    out_.clear();
    if (kind_ == 0) {
      out_.assign(m1_.idx + m1_.ptr[r], m1_.idx + m1_.ptr[r + 1]);
      return out_;
    }
    ++epoch_;
    if (epoch_ == 0) {   // stamp wrap-around
      std::fill(stamp_a_.begin(), stamp_a_.end(), 0u);
      std::fill(stamp_b_.begin(), stamp_b_.end(), 0u);
      epoch_ = 1;
    }
    mid_.clear();
    for (int64_t p = m1_.ptr[r]; p < m1_.ptr[r + 1]; ++p) {
      const int32_t j = m1_.idx[p];
      for (int64_t q = m2_.ptr[j]; q < m2_.ptr[j + 1]; ++q) {
        const int32_t k = m2_.idx[q];
        if (stamp_a_[(size_t)k] != epoch_) {
          stamp_a_[(size_t)k] = epoch_;
          mid_.push_back(k);
        }
      }
    }
    std::reverse(mid_.begin(), mid_.end());
    if (kind_ == 1) {
      out_.swap(mid_);
      return out_;
    }
    for (size_t t = 0; t < mid_.size(); ++t) {
      const int32_t j = mid_[t];
      for (int64_t q = m3_.ptr[j]; q < m3_.ptr[j + 1]; ++q) {
        const int32_t k = m3_.idx[q];
        if (stamp_b_[(size_t)k] != epoch_) {
          stamp_b_[(size_t)k] = epoch_;
          out_.push_back(k);
        }
      }
    }
    std::reverse(out_.begin(), out_.end());
    return out_;
  }

 private:
  int kind_;
  Csr m1_, m2_, m3_;
  std::vector<uint32_t> stamp_a_, stamp_b_;
  std::vector<int32_t> mid_, out_;
  uint32_t epoch_ = 0;
};

}  // namespace

extern "C" {

int hge_mt19937_random_raw(uint32_t* state625, int64_t n, uint32_t* out) {
  HGE_REQUIRE(state625 && (out || n == 0) && n >= 0, "hge_mt19937_random_raw: bad argument");
  HGE_REQUIRE(state625[624] <= 624, "hge_mt19937_random_raw: pos %u out of range", state625[624]);
  StateGuard g(state625);
  for (int64_t i = 0; i < n; ++i) out[i] = g.mt.next32();
  return HGE_OK;
}

int hge_mt19937_interval(uint32_t* state625, uint32_t max_inclusive, int64_t n, uint32_t* out) {
  HGE_REQUIRE(state625 && (out || n == 0) && n >= 0, "hge_mt19937_interval: bad argument");
  HGE_REQUIRE(state625[624] <= 624, "hge_mt19937_interval: pos %u out of range", state625[624]);
  StateGuard g(state625);
  for (int64_t i = 0; i < n; ++i) out[i] = g.mt.interval(max_inclusive);
  return HGE_OK;
}

int hge_spgemm_rows(int kind, const int64_t* p1, const int32_t* i1, const int64_t* p2,
                    const int32_t* i2, const int64_t* p3, const int32_t* i3, int32_t mid_cols,
                    int32_t out_cols, const int32_t* rows, int64_t num_rows, int sorted,
                    int64_t* out_ptr, int32_t* out_idx, int64_t capacity) {
  HGE_REQUIRE(kind >= 0 && kind <= 2, "hge_spgemm_rows: kind must be 0, 1 or 2");
  HGE_REQUIRE(p1 && i1 && (kind < 1 || (p2 && i2)) && (kind < 2 || (p3 && i3)),
              "hge_spgemm_rows: NULL matrix");
  HGE_REQUIRE(out_ptr && num_rows >= 0 && (rows || num_rows == 0), "hge_spgemm_rows: bad rows");
  RowBuilder rb(kind, Csr{p1, i1}, Csr{p2, i2}, Csr{p3, i3}, mid_cols, out_cols);
  int64_t total = 0;
  out_ptr[0] = 0;
  std::vector<int32_t> tmp;
  for (int64_t t = 0; t < num_rows; ++t) {
    const std::vector<int32_t>& c = rb.row(rows[t]);
    if (out_idx) {
      HGE_REQUIRE(total + (int64_t)c.size() <= capacity, "hge_spgemm_rows: capacity %lld too small",
                  (long long)capacity);
      std::copy(c.begin(), c.end(), out_idx + total);
      if (sorted) std::sort(out_idx + total, out_idx + total + c.size());
    }
    total += (int64_t)c.size();
    out_ptr[t + 1] = total;
  }
  return HGE_OK;
}

int hge_sample_adj_rows(int kind, const int64_t* p1, const int32_t* i1, const int64_t* p2,
                        const int32_t* i2, const int64_t* p3, const int32_t* i3, int32_t mid_cols,
                        int32_t out_cols, const int32_t* rows, int64_t num_rows,
                        const int32_t* samples_per_row, int replace, int negative,
                        uint32_t* state625, int32_t* out_row, int32_t* out_col, int64_t capacity,
                        int64_t* out_count) {
  HGE_REQUIRE(kind >= 0 && kind <= 2, "hge_sample_adj_rows: kind must be 0, 1 or 2");
  HGE_REQUIRE(p1 && i1 && (kind < 1 || (p2 && i2)) && (kind < 2 || (p3 && i3)),
              "hge_sample_adj_rows: NULL matrix");
  HGE_REQUIRE(rows && samples_per_row && state625 && out_count, "hge_sample_adj_rows: NULL argument");
  HGE_REQUIRE(num_rows > 0, "hge_sample_adj_rows: no rows (hg2v_sample.py:67)");
  HGE_REQUIRE(state625[624] <= 624, "hge_sample_adj_rows: RNG pos %u out of range", state625[624]);
  HGE_REQUIRE(out_cols > 0, "hge_sample_adj_rows: out_cols must be positive");
  StateGuard g(state625);
  RowBuilder rb(kind, Csr{p1, i1}, Csr{p2, i2}, Csr{p3, i3}, mid_cols, out_cols);
  std::vector<int32_t> perm;
  int64_t n_out = 0;
  auto emit = [&](int32_t r, int32_t c) -> bool {
    if (n_out >= capacity) return false;
    out_row[n_out] = r;
    out_col[n_out] = c;
    ++n_out;
    return true;
  };
  for (int64_t t = 0; t < num_rows; ++t) {
    const int32_t r = rows[t];
    const int32_t want = samples_per_row[t];
    HGE_REQUIRE(want >= 0, "hge_sample_adj_rows: negative sample count");
    if (negative) {
      // np.random.randint(matrix.shape[1], size=num_samples), hg2v_sample.py:74
      for (int32_t s = 0; s < want; ++s)
        if (!emit(r, (int32_t)g.mt.interval((uint32_t)out_cols - 1))) goto full;
      continue;
    }
    {
      const std::vector<int32_t>& cand = rb.row(r);
      const int64_t n = (int64_t)cand.size();
      if (n == 0) continue;   // hg2v_sample.py:79
      if (!replace) {
        // np.random.choice(cols, min(k, n), replace=False) == cols[permutation(n)[:k]]
        const int64_t k = std::min<int64_t>(want, n);
        perm.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;
        for (int64_t i = n - 1; i >= 1; --i) {
          const uint32_t j = g.mt.interval((uint32_t)i);
          std::swap(perm[(size_t)i], perm[j]);
        }
        for (int64_t s = 0; s < k; ++s)
          if (!emit(r, cand[(size_t)perm[(size_t)s]])) goto full;
      } else {
        for (int32_t s = 0; s < want; ++s)
          if (!emit(r, cand[g.mt.interval((uint32_t)(n - 1))])) goto full;
      }
    }
  }
  *out_count = n_out;
  return HGE_OK;
full:
  hge_set_error("hge_sample_adj_rows: output capacity %lld too small", (long long)capacity);
  return HGE_ERR_INVALID;
}

int hge_sample_neighbors(const int64_t* n2e_ptr, const int32_t* n2e_idx, const int64_t* e2n_ptr,
                         const int32_t* e2n_idx, const int32_t* nodes, const int32_t* edges,
                         int64_t num_samples, int k, uint32_t* state625, int32_t* out_nbr_edges,
                         int32_t* out_nbr_nodes) {
  HGE_REQUIRE(n2e_ptr && n2e_idx && e2n_ptr && e2n_idx && state625, "hge_sample_neighbors: NULL");
  HGE_REQUIRE(num_samples >= 0 && k >= 0, "hge_sample_neighbors: negative count");
  HGE_REQUIRE(num_samples == 0 || (nodes && edges), "hge_sample_neighbors: NULL sample arrays");
  HGE_REQUIRE(k == 0 || num_samples == 0 || (out_nbr_edges && out_nbr_nodes),
              "hge_sample_neighbors: NULL output");
  HGE_REQUIRE(state625[624] <= 624, "hge_sample_neighbors: RNG pos %u out of range", state625[624]);
  StateGuard g(state625);
  for (int64_t s = 0; s < num_samples; ++s) {
    // edges of the node first, then nodes of the edge (hg2v_sample.py:184-187, 604-605)
    const int64_t nb = n2e_ptr[nodes[s]], nd = n2e_ptr[nodes[s] + 1] - nb;
    const int64_t eb = e2n_ptr[edges[s]], ed = e2n_ptr[edges[s] + 1] - eb;
    if (k > 0 && (nd == 0 || ed == 0)) {
      hge_set_error("hge_sample_neighbors: sample %lld has no neighbours to draw from "
                    "(numpy: 'a' cannot be empty unless no samples are taken)", (long long)s);
      return HGE_ERR_INVALID;
    }
    for (int t = 0; t < k; ++t)
      out_nbr_edges[s * k + t] = n2e_idx[nb + g.mt.interval((uint32_t)(nd - 1))];
    for (int t = 0; t < k; ++t)
      out_nbr_nodes[s * k + t] = e2n_idx[eb + g.mt.interval((uint32_t)(ed - 1))];
  }
  return HGE_OK;
}

}  // extern "C"
