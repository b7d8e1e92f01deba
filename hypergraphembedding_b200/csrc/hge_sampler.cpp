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

#include <immintrin.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "hge_common.cuh"

namespace {

// ---- numpy legacy MT19937 -------------------------------------------------------------------
// The generator is run one whole 624-word generation ahead: `tempered` holds the outputs of the
// current generation followed by those of the next one, so a bounded draw can look at the next
// four outputs at once and pick the first accepted one without a data-dependent branch (the
// rejection branch of rk_interval mispredicts ~25 % of the time and dominated the draw loop).
// The state handed back is numpy's own lazy representation: the key of the generation the last
// consumed output belongs to, and pos in [0, 624].
struct Mt19937 {
  uint32_t cur[624], nxt[624];
  uint32_t tempered[1248 + 4];
  int pos;

  static void regenerate(uint32_t* key) {
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
  }

  static void temper(const uint32_t* key, uint32_t* out) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = key[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      out[i] = y;
    }
  }

  // The same two steps 16 words at a time (AVX-512).  A block of the twist reads key[i+1 ..
  // i+16] and key[i+397 ..] (first part) or key[i-227 ..] (second part, already new) before it
  // writes key[i .. i+15], so whole blocks are independent of their own output.
  __attribute__((target("avx512f"))) static void twist16(uint32_t* key, int i, int other) {
    const __m512i upper = _mm512_set1_epi32((int)0x80000000u), lower = _mm512_set1_epi32(0x7fffffff);
    const __m512i matrix = _mm512_set1_epi32((int)0x9908b0dfu), one = _mm512_set1_epi32(1);
    const __m512i a = _mm512_loadu_si512(key + i), b = _mm512_loadu_si512(key + i + 1);
    const __m512i y = _mm512_or_si512(_mm512_and_si512(a, upper), _mm512_and_si512(b, lower));
    const __mmask16 odd = _mm512_test_epi32_mask(y, one);
    __m512i r = _mm512_xor_si512(_mm512_loadu_si512(key + other), _mm512_srli_epi32(y, 1));
    r = _mm512_mask_xor_epi32(r, odd, r, matrix);
    _mm512_storeu_si512(key + i, r);
  }

  __attribute__((target("avx512f"))) static void regenerate_wide(uint32_t* key) {
    const uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
    int i = 0;
    for (; i + 16 <= 624 - 397; i += 16) twist16(key, i, i + 397);
    for (; i < 624 - 397; ++i) {
      const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    for (; i + 16 <= 623; i += 16) twist16(key, i, i + (397 - 624));
    for (; i < 623; ++i) {
      const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    const uint32_t y = (key[623] & kUpper) | (key[0] & kLower);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
  }

  __attribute__((target("avx512f"))) static void temper_wide(const uint32_t* key, uint32_t* out) {
    const __m512i m7 = _mm512_set1_epi32((int)0x9d2c5680u), m15 = _mm512_set1_epi32((int)0xefc60000u);
    for (int i = 0; i < 624; i += 16) {
      __m512i y = _mm512_loadu_si512(key + i);
      y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
      y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), m7));
      y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), m15));
      y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
      _mm512_storeu_si512(out + i, y);
    }
  }

  static void next_generation(uint32_t* key) {
    static const bool wide = __builtin_cpu_supports("avx512f");
    if (wide) regenerate_wide(key); else regenerate(key);
  }
  static void temper_all(const uint32_t* key, uint32_t* out) {
    static const bool wide = __builtin_cpu_supports("avx512f");
    if (wide) temper_wide(key, out); else temper(key, out);
  }

  void init(const uint32_t* key, int pos0) {
    memcpy(cur, key, sizeof(cur));
    memcpy(nxt, key, sizeof(nxt));
    next_generation(nxt);
    temper_all(cur, tempered);
    temper_all(nxt, tempered + 624);
    memset(tempered + 1248, 0, 4 * sizeof(uint32_t));
    pos = pos0;
  }

  // pos >= 624: the next output belongs to the following generation
  void advance_generation() {
    memcpy(cur, nxt, sizeof(cur));
    memcpy(tempered, tempered + 624, 624 * sizeof(uint32_t));
    next_generation(nxt);
    temper_all(nxt, tempered + 624);
    pos -= 624;
  }

  inline uint32_t next32() {
    if (pos >= 624) advance_generation();
    return tempered[pos++];
  }

  static inline uint32_t mask_of(uint32_t mx) {
    uint32_t mask = mx;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    return mask;
  }

  // `count` draws of rk_interval(mx) (uniform in [0, mx]; mx == 0 consumes nothing) into out[]:
  // the raw stream is filtered without a
  // data-dependent branch (every masked output is stored, the write index only advances when
  // it is accepted); the rejection branch of the scalar form mispredicts ~25 % of the time.
  void interval_many(uint32_t mx, int64_t count, uint32_t* out) {
    if (mx == 0) {
      for (int64_t c = 0; c < count; ++c) out[c] = 0;
      return;
    }
    const uint32_t mask = mask_of(mx);
    static const bool wide = __builtin_cpu_supports("avx512f");
    int64_t c = 0;
    while (c < count) {
      if (pos >= 624) advance_generation();
      int p = pos;
      if (wide) {
        // constant bound: 16 words at a time while a whole block still fits into `count`
        while (p + 16 <= 1248 && count - c >= 16) {
          c += accept16(tempered + p, mask, mx, out + c);
          p += 16;
        }
      }
      while (p < 1248 && c < count) {
        const uint32_t v = tempered[p++] & mask;
        out[c] = v;
        c += (v <= mx);
      }
      pos = p;
    }
  }

  __attribute__((target("avx512f"))) static int accept16(const uint32_t* w, uint32_t mask, uint32_t mx,
                                                         uint32_t* out) {
    const __m512i v = _mm512_and_si512(_mm512_loadu_si512(w), _mm512_set1_epi32((int)mask));
    const __mmask16 ok = _mm512_cmple_epu32_mask(v, _mm512_set1_epi32((int)mx));
    // register compress + full 64-byte store (the masked compress-store to memory is slow);
    // callers leave 16 words of slack behind the region being filled
    _mm512_storeu_si512(out, _mm512_maskz_compress_epi32(ok, v));
    return __builtin_popcount((unsigned)ok);
  }

  // The draws of a Fisher-Yates shuffle of n items in the order they are consumed: out[q] =
  // interval(n - 1 - q) for q = 0 .. n - 2, by the same branch-free filter; the bound (and, at
  // powers of two, the mask) shrinks by one with every accepted output.  Where the bound is far
  // from the next power of two, 16 words are decided at once (AVX-512, filter16 below).  Same
  // stream, same results -- only faster on long rows.
  void shuffle_draws(uint32_t n, uint32_t* out) {
    if (n < 2) return;
    static const bool wide = __builtin_cpu_supports("avx512f");
    uint32_t i = n - 1, mask = mask_of(i);
    size_t q = 0;
    while (i >= 1) {
      if (pos >= 624) advance_generation();
      int p = pos;
      if (wide) {
        while (p + 16 <= 1248 && i >= (mask >> 1) + 16) {
          const int c = filter16(tempered + p, mask, i, out + q);
          p += 16;
          q += (size_t)c;
          i -= (uint32_t)c;
          mask = (i <= (mask >> 1)) ? (mask >> 1) : mask;   // a full block can end on the boundary
        }
      }
      const int limit = p + 16 < 1248 ? p + 16 : 1248;
      while (p < limit && i >= 1) {
        const uint32_t v = tempered[p++] & mask;
        out[q] = v;
        const uint32_t acc = v <= i;
        q += acc;
        i -= acc;
        mask = (i <= (mask >> 1)) ? (mask >> 1) : mask;
      }
      pos = p;
    }
  }

  // 16 tempered words against bound i with a constant mask: number accepted (their masked values
  // appended to out).  Word p is accepted for sure if it is <= i - p and rejected for sure if it
  // is > i; the few words in between are resolved in lane order from the accept mask built so
  // far (bound of lane p = i - number of accepted lanes before p).
  __attribute__((target("avx512f"))) static int filter16(const uint32_t* w, uint32_t mask, uint32_t i,
                                                         uint32_t* out) {
    const __m512i v = _mm512_and_si512(_mm512_loadu_si512(w), _mm512_set1_epi32((int)mask));
    const __m512i lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i top = _mm512_set1_epi32((int)i);
    unsigned acc = _mm512_cmple_epu32_mask(v, _mm512_sub_epi32(top, lane));
    unsigned open = (unsigned)_mm512_cmple_epu32_mask(v, top) & ~acc;
    if (open) {
      alignas(64) uint32_t vv[16];
      _mm512_store_si512(vv, v);
      while (open) {
        const int p = __builtin_ctz(open);
        open &= open - 1;
        const uint32_t before = (uint32_t)__builtin_popcount(acc & ((1u << p) - 1u));
        if (vv[p] <= i - before) acc |= 1u << p;
      }
    }
    _mm512_storeu_si512(out, _mm512_maskz_compress_epi32((__mmask16)acc, v));
    return __builtin_popcount(acc);
  }
};

// state625 = key[624] followed by pos
struct StateGuard {
  Mt19937 mt;
  uint32_t* buf;
  explicit StateGuard(uint32_t* state625) : buf(state625) { mt.init(state625, (int)state625[624]); }
  ~StateGuard() {
    while (mt.pos > 624) mt.advance_generation();
    memcpy(buf, mt.cur, 624 * sizeof(uint32_t));
    buf[624] = (uint32_t)mt.pos;
  }
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

// Three-stage pipeline around the one thing that is sequential, the MT19937 stream:
//   build   worker threads build the candidate rows of blocks of consecutive rows (each worker
//           with its own RowBuilder);
//   draw    the single consumer walks the blocks in order and only runs the branch-free draw
//           filters (one bounded draw per candidate for a shuffle), writing the draws next to the
//           candidates and fixing each row's position in the output;
//   apply   worker threads turn draws into samples: Fisher-Yates swaps on a private permutation,
//           first k entries -> (row, column) pairs at the row's output offset.
// A slot is reused for block b + depth once block b has been applied.  The same workers serve
// both stages (apply first: it frees slots).
class SamplePipeline {
 public:
  struct Block {
    std::vector<int64_t> ptr;      // candidate offsets per row (rows + 1)
    std::vector<int32_t> idx;      // candidates
    std::vector<uint32_t> draws;   // draw stage output
    std::vector<int64_t> doff;     // per row: offset of its draws
    std::vector<int64_t> ooff;     // per row: offset of its samples in the output (-1: none)
    std::vector<int32_t> count;    // per row: samples to emit
    int64_t id = -1;
    int state = 0;                 // 0 free, 1 building, 2 built, 3 drawn, 4 applying
  };

  SamplePipeline(int kind, Csr m1, Csr m2, Csr m3, int32_t mid_cols, int32_t out_cols,
                 const int32_t* rows, int64_t num_rows, int threads, bool replace, int32_t* out_row,
                 int32_t* out_col)
      : rows_(rows), num_rows_(num_rows), replace_(replace), out_row_(out_row), out_col_(out_col) {
    num_blocks_ = (num_rows + kBlockRows - 1) / kBlockRows;
    depth_ = std::max<int64_t>(4, 8 * (int64_t)threads);
    slots_.resize((size_t)depth_);
    for (int t = 0; t < threads; ++t)
      workers_.emplace_back([=] { work(kind, m1, m2, m3, mid_cols, out_cols); });
  }

  ~SamplePipeline() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (std::thread& t : workers_) t.join();
  }

  int64_t num_blocks() const { return num_blocks_; }
  static int64_t block_rows() { return kBlockRows; }

  // Blocks must be taken in order 0, 1, 2, ...; the caller fills draws / doff / ooff / count.
  Block* wait_built(int64_t b) {
    std::unique_lock<std::mutex> lk(mu_);
    Block& blk = slots_[(size_t)(b % depth_)];
    cv_built_.wait(lk, [&] { return blk.id == b && blk.state == 2; });
    return &blk;
  }

  void submit_drawn(Block* blk) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      blk->state = 3;
      ++drawn_pending_;
    }
    cv_work_.notify_one();
  }

  // Returns once every submitted block has been applied.
  void finish() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_applied_.wait(lk, [&] { return drawn_pending_ == 0 && applying_ == 0; });
  }

 private:
  static const int64_t kBlockRows = 64;   // small blocks: the consumer starts after 64 rows

  void apply(Block& blk, std::vector<int32_t>& perm) {
    const int64_t r0 = blk.id * kBlockRows;
    const int64_t nrows = (int64_t)blk.ptr.size() - 1;
    for (int64_t k = 0; k < nrows; ++k) {
      const int64_t o = blk.ooff[(size_t)k];
      if (o < 0) continue;
      const int32_t r = rows_[r0 + k];
      const int32_t* cand = blk.idx.data() + blk.ptr[(size_t)k];
      const int64_t n = blk.ptr[(size_t)k + 1] - blk.ptr[(size_t)k];
      const uint32_t* d = blk.draws.data() + blk.doff[(size_t)k];
      const int32_t cnt = blk.count[(size_t)k];
      if (!replace_) {
        // np.random.choice(cols, min(k, n), replace=False) == cols[permutation(n)[:k]]
        perm.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;
        for (int64_t i = n - 1, t = 0; i >= 1; --i, ++t) std::swap(perm[(size_t)i], perm[d[t]]);
        for (int32_t q = 0; q < cnt; ++q) {
          out_row_[o + q] = r;
          out_col_[o + q] = cand[(size_t)perm[(size_t)q]];
        }
      } else {
        for (int32_t q = 0; q < cnt; ++q) {
          out_row_[o + q] = r;
          out_col_[o + q] = cand[d[q]];
        }
      }
    }
  }

  void work(int kind, Csr m1, Csr m2, Csr m3, int32_t mid_cols, int32_t out_cols) {
    RowBuilder rb(kind, m1, m2, m3, mid_cols, out_cols);
    std::vector<int32_t> perm;
    for (;;) {
      Block* todo = nullptr;
      bool is_apply = false;
      {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
          if (stop_) return;
          if (drawn_pending_ > 0) {          // apply first: it is what frees slots
            for (Block& blk : slots_)
              if (blk.state == 3 && (!todo || blk.id < todo->id)) todo = &blk;
            if (todo) {
              todo->state = 4;
              --drawn_pending_;
              ++applying_;
              is_apply = true;
              break;
            }
          }
          if (next_ < num_blocks_ && slots_[(size_t)(next_ % depth_)].state == 0) {
            todo = &slots_[(size_t)(next_ % depth_)];
            todo->state = 1;
            todo->id = next_++;
            break;
          }
          cv_work_.wait(lk);
        }
      }
      if (is_apply) {
        apply(*todo, perm);
        {
          std::lock_guard<std::mutex> lk(mu_);
          todo->state = 0;
          --applying_;
        }
        cv_work_.notify_all();      // a slot is free again
        cv_applied_.notify_all();
      } else {
        const int64_t r0 = todo->id * kBlockRows, r1 = std::min(num_rows_, r0 + kBlockRows);
        todo->ptr.assign(1, 0);
        todo->idx.clear();
        for (int64_t t = r0; t < r1; ++t) {
          const std::vector<int32_t>& c = rb.row(rows_[t]);
          todo->idx.insert(todo->idx.end(), c.begin(), c.end());
          todo->ptr.push_back((int64_t)todo->idx.size());
        }
        {
          std::lock_guard<std::mutex> lk(mu_);
          todo->state = 2;
        }
        cv_built_.notify_all();
      }
    }
  }

  const int32_t* rows_;
  int64_t num_rows_, num_blocks_ = 0, depth_ = 4;
  bool replace_;
  int32_t* out_row_;
  int32_t* out_col_;
  std::vector<Block> slots_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_built_, cv_applied_;
  int64_t next_ = 0, drawn_pending_ = 0, applying_ = 0;
  bool stop_ = false;
};

std::atomic<int> g_sampler_threads{0};   // 0 = automatic

// fn(lo, hi) over [0, n) in contiguous slices, on up to 4 threads when there is enough to do
template <typename F>
void parallel_for(int64_t n, F fn) {
  int threads = g_sampler_threads.load();
  if (threads == 0) threads = (int)std::min<unsigned>(4u, std::max(1u, std::thread::hardware_concurrency()));
  if (n < 65536 || threads <= 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t)
    pool.emplace_back([=] { fn(n * t / threads, n * (t + 1) / threads); });
  fn(0, n / threads);
  for (std::thread& th : pool) th.join();
}

int sampler_threads_for(int kind, int64_t num_rows) {
  int n = g_sampler_threads.load();
  if (n == 0) {
    const char* env = getenv("HGE_SAMPLER_THREADS");
    if (env && *env) n = atoi(env);
  }
  if (n == 0) {
    if (kind == 0 || num_rows < 512) return 1;   // small jobs: thread start-up costs more
    // measured (16-core host, 3.5e8 candidates): 2.42 s inline, 1.11 s with 2 to 12 workers --
    // from two workers on the draw loop (3.1 ns per candidate) is all that is left
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    n = (int)std::min(4u, hw);
    // three-factor rows (the node-edge pairs of HOBE / FOBE: candidates two hops away) cost far
    // more to build than to draw from; on the 100 000-node HOBE case the builders scale to 8+
    // workers (3.9 s with 1, 0.9 s with 4 on an 8-core host)
    if (kind == 2) n = (int)std::min(16u, std::max(4u, hw / 2));
  }
  return std::max(1, n);
}

}  // namespace

extern "C" {

int hge_sampler_set_threads(int threads) {
  HGE_REQUIRE(threads >= 0 && threads <= 256, "hge_sampler_set_threads: 0 (automatic) .. 256");
  g_sampler_threads.store(threads);
  return HGE_OK;
}

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
  g.mt.interval_many(max_inclusive, n, out);
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
  for (int64_t t = 0; t < num_rows; ++t)
    HGE_REQUIRE(samples_per_row[t] >= 0, "hge_sample_adj_rows: negative sample count");
  const int threads = negative ? 1 : sampler_threads_for(kind, num_rows);
  int64_t n_out = 0;

  if (threads > 1) {
    // pipelined: this thread only draws (see SamplePipeline)
    SamplePipeline pipe(kind, Csr{p1, i1}, Csr{p2, i2}, Csr{p3, i3}, mid_cols, out_cols, rows,
                        num_rows, threads, replace != 0, out_row, out_col);
    bool full = false;
    for (int64_t b = 0; b < pipe.num_blocks() && !full; ++b) {
      SamplePipeline::Block* blk = pipe.wait_built(b);
      const int64_t r0 = b * SamplePipeline::block_rows();
      const int64_t nrows = (int64_t)blk->ptr.size() - 1;
      blk->doff.assign((size_t)nrows, 0);
      blk->ooff.assign((size_t)nrows, -1);
      blk->count.assign((size_t)nrows, 0);
      int64_t need = 0;
      for (int64_t k = 0; k < nrows; ++k) {
        const int64_t n = blk->ptr[(size_t)k + 1] - blk->ptr[(size_t)k];
        blk->doff[(size_t)k] = need;
        if (n > 0) need += replace ? samples_per_row[r0 + k] : n;
      }
      blk->draws.resize((size_t)need + 16);   // 16 words of slack for the block filter's stores
      for (int64_t k = 0; k < nrows; ++k) {
        const int64_t n = blk->ptr[(size_t)k + 1] - blk->ptr[(size_t)k];
        if (n == 0) continue;   // hg2v_sample.py:79
        const int32_t want = samples_per_row[r0 + k];
        uint32_t* d = blk->draws.data() + blk->doff[(size_t)k];
        int32_t cnt;
        if (!replace) {
          g.mt.shuffle_draws((uint32_t)n, d);
          cnt = (int32_t)std::min<int64_t>(want, n);
        } else {
          g.mt.interval_many((uint32_t)(n - 1), want, d);
          cnt = want;
        }
        if (n_out + cnt > capacity) {
          full = true;
          break;
        }
        blk->ooff[(size_t)k] = n_out;
        blk->count[(size_t)k] = cnt;
        n_out += cnt;
      }
      pipe.submit_drawn(blk);
    }
    pipe.finish();
    if (full) {
      hge_set_error("hge_sample_adj_rows: output capacity %lld too small", (long long)capacity);
      return HGE_ERR_INVALID;
    }
    *out_count = n_out;
    return HGE_OK;
  }

  RowBuilder rb(kind, Csr{p1, i1}, Csr{p2, i2}, Csr{p3, i3}, mid_cols, out_cols);
  std::vector<int32_t> perm;
  std::vector<uint32_t> draws;
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
    if (negative) {
      // np.random.randint(matrix.shape[1], size=num_samples), hg2v_sample.py:74
      draws.resize((size_t)want);
      g.mt.interval_many((uint32_t)out_cols - 1, want, draws.data());
      for (int32_t s = 0; s < want; ++s)
        if (!emit(r, (int32_t)draws[(size_t)s])) goto full;
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
        draws.resize((size_t)n + 16);
        for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;
        g.mt.shuffle_draws((uint32_t)n, draws.data());
        for (int64_t i = n - 1, t = 0; i >= 1; --i, ++t)
          std::swap(perm[(size_t)i], perm[draws[(size_t)t]]);
        for (int64_t s = 0; s < k; ++s)
          if (!emit(r, cand[(size_t)perm[(size_t)s]])) goto full;
      } else {
        draws.resize((size_t)want);
        g.mt.interval_many((uint32_t)(n - 1), want, draws.data());
        for (int32_t s = 0; s < want; ++s)
          if (!emit(r, cand[draws[(size_t)s]])) goto full;
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
  if (k == 0 || num_samples == 0) return HGE_OK;
  // Three passes, only the middle one sequential: (1) degrees and row starts of every sample's
  // node and edge (random reads of the row pointers, in parallel); (2) the draws, k per side,
  // edges of the node first, then nodes of the edge (hg2v_sample.py:184-187, 604-605), streaming
  // through the arrays of pass 1; (3) draws -> neighbour ids (random reads of the column ids, in
  // parallel).  The draws are parked in the output arrays.
  std::vector<int64_t> base((size_t)num_samples * 2);
  std::vector<uint32_t> bound((size_t)num_samples * 2);
  std::atomic<int64_t> empty_sample{-1};
  parallel_for(num_samples, [&](int64_t lo, int64_t hi) {
    for (int64_t s = lo; s < hi; ++s) {
      const int64_t nb = n2e_ptr[nodes[s]], nd = n2e_ptr[nodes[s] + 1] - nb;
      const int64_t eb = e2n_ptr[edges[s]], ed = e2n_ptr[edges[s] + 1] - eb;
      if (nd <= 0 || ed <= 0) {
        int64_t seen = empty_sample.load();
        while ((seen < 0 || s < seen) && !empty_sample.compare_exchange_weak(seen, s)) {
        }
      }
      base[(size_t)s * 2] = nb;
      base[(size_t)s * 2 + 1] = eb;
      bound[(size_t)s * 2] = (uint32_t)std::max<int64_t>(nd - 1, 0);
      bound[(size_t)s * 2 + 1] = (uint32_t)std::max<int64_t>(ed - 1, 0);
    }
  });
  if (empty_sample.load() >= 0) {
    hge_set_error("hge_sample_neighbors: sample %lld has no neighbours to draw from "
                  "(numpy: 'a' cannot be empty unless no samples are taken)",
                  (long long)empty_sample.load());
    return HGE_ERR_INVALID;
  }
  uint32_t* draw_e = reinterpret_cast<uint32_t*>(out_nbr_edges);
  uint32_t* draw_n = reinterpret_cast<uint32_t*>(out_nbr_nodes);
  for (int64_t s = 0; s < num_samples; ++s) {
    g.mt.interval_many(bound[(size_t)s * 2], k, draw_e + s * k);
    g.mt.interval_many(bound[(size_t)s * 2 + 1], k, draw_n + s * k);
  }
  parallel_for(num_samples, [&](int64_t lo, int64_t hi) {
    for (int64_t s = lo; s < hi; ++s) {
      const int32_t* row_e = n2e_idx + base[(size_t)s * 2];
      const int32_t* row_n = e2n_idx + base[(size_t)s * 2 + 1];
      for (int t = 0; t < k; ++t) {
        out_nbr_edges[s * k + t] = row_e[draw_e[s * k + t]];
        out_nbr_nodes[s * k + t] = row_n[draw_n[s * k + t]];
      }
    }
  });
  return HGE_OK;
}

}  // extern "C"
