// Roof of the access pattern the half-sweep is made of: random gathers of 128-byte rows (one
// float4 per lane, 8 lanes per row, 4 rows per warp-wide load) from a table of a given size,
// U loads in flight per lane, ids read coalesced from a pre-drawn array.  Prints GB/s of
// gathered bytes per table size: the L2-resident and the DRAM-resident ceilings of the pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/roof/gather_roof tools/roof/gather_roof.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ table, const uint32_t* __restrict__ ids,
                                              long long per_warp, float4* out) {
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t* p = ids + warp * per_warp * 4 + g;     // 4 ids per warp-step, one per group
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = 0; i < per_warp; i += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t c = __ldcs(p + (i + u) * 4);
      v[u] = __ldg(table + (size_t)c * 8 + gl);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  if (acc.x == 123.456f) out[0] = acc;
}

// The traffic mix of a half-sweep without any of its bookkeeping: every lane group owns one row
// per round -- 8 gathers from the table, one 128-byte read and one 128-byte write of the owned
// row.  own_random = 0: owned rows are visited in address order, 1: in a random order (what a
// degree-sorted schedule does to them).
__global__ void __launch_bounds__(256, 4) k_mix(const float4* __restrict__ table, const uint32_t* __restrict__ ids,
                                               float4* own, const uint32_t* __restrict__ own_order,
                                               long long rows_per_warp, int own_random) {
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t* p = ids + warp * rows_per_warp * 8 * 4 + g;
  for (long long i = 0; i < rows_per_warp; ++i) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t c = __ldcs(p + (i * 8 + u) * 4);
      v[u] = __ldg(table + (size_t)c * 8 + gl);
    }
    const long long r = (warp * rows_per_warp + i) * 4 + g;
    const size_t orow = own_random ? own_order[r] : (size_t)r;
    float4 acc = __ldcs(own + orow * 8 + gl);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
    own[orow * 8 + gl] = acc;
  }
}

// The same mix with L2 eviction priorities: table rows evict_last, owned rows (read once,
// written once) evict_first.
__device__ __forceinline__ float4 ld_hint(const float4* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__global__ void __launch_bounds__(256, 4) k_mix_hint(const float4* __restrict__ table,
                                                    const uint32_t* __restrict__ ids, float4* own,
                                                    long long rows_per_warp, int own_first) {
  unsigned long long keep, once;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(once));
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t* p = ids + warp * rows_per_warp * 8 * 4 + g;
  for (long long i = 0; i < rows_per_warp; ++i) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t c = __ldcs(p + (i * 8 + u) * 4);
      v[u] = ld_hint(table + (size_t)c * 8 + gl, keep);
    }
    const size_t orow = (size_t)((warp * rows_per_warp + i) * 4 + g);
    float4 acc = own_first ? ld_hint(own + orow * 8 + gl, once) : __ldcs(own + orow * 8 + gl);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
    if (own_first)
      asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(own + orow * 8 + gl),
                   "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w), "l"(once) : "memory");
    else
      own[orow * 8 + gl] = acc;
  }
}

// The mix with config 2's own shape: ROUNDS rounds of 10 gathers per owned row (node half: 1M
// rows x 10 gathers from the 64 MB edge block; edge half: 500K rows x 20 gathers from the 128 MB
// node block), row read, row written.
template <int ROUNDS>
__global__ void __launch_bounds__(256, 4) k_mix_c2(const float4* __restrict__ table,
                                                  const uint32_t* __restrict__ ids, float4* own,
                                                  long long rows_per_warp) {
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t* p = ids + warp * rows_per_warp * (10 * ROUNDS) * 4 + g;
  for (long long i = 0; i < rows_per_warp; ++i) {
    const size_t orow = (size_t)((warp * rows_per_warp + i) * 4 + g);
    float4 acc = __ldcs(own + orow * 8 + gl);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      float4 v[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        const uint32_t c = __ldcs(p + ((i * ROUNDS + r) * 10 + u) * 4);
        v[u] = __ldg(table + (size_t)c * 8 + gl);
      }
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    __stcs(own + orow * 8 + gl, acc);
  }
}

template <int ROUNDS>
float run_c2(const float4* table, const uint32_t* ids, float4* own, long long own_rows, int blocks) {
  const long long rpw = own_rows / 4 / ((long long)blocks * 8);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_mix_c2<ROUNDS><<<blocks, 256>>>(table, ids, own, rpw);
  CK(cudaEventRecord(a));
  for (int r = 0; r < 5; ++r) k_mix_c2<ROUNDS><<<blocks, 256>>>(table, ids, own, rpw);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / 5 * (float)((double)own_rows / ((double)rpw * 4 * blocks * 8));
}

__global__ void k_fill_ids(uint32_t* ids, long long n, uint32_t rows, uint64_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
    ids[i] = (uint32_t)(x % rows);
  }
}

template <int U>
float run(const float4* table, const uint32_t* ids, long long per_warp, int blocks, float4* out) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_gather<U><<<blocks, 256>>>(table, ids, per_warp, out);
  CK(cudaEventRecord(a));
  for (int r = 0; r < 5; ++r) k_gather<U><<<blocks, 256>>>(table, ids, per_warp, out);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / 5;
}

// config 2's own shape, uniform-random ids
static void c2_section(int sms) {
  {
    printf("\nconfig-2 shape (uniform-random ids; the real edge sizes are heavy-tailed, which L1 / L2 like better):\n");
    float4 *tn, *te, *own;
    uint32_t* idc;
    CK(cudaMalloc(&tn, 128ull << 20)); CK(cudaMalloc(&te, 64ull << 20)); CK(cudaMalloc(&own, 128ull << 20));
    CK(cudaMemset(tn, 0, 128ull << 20)); CK(cudaMemset(te, 0, 64ull << 20)); CK(cudaMemset(own, 0, 128ull << 20));
    CK(cudaMalloc(&idc, 10ll * (1 << 20) * 4));
    for (int bps : {4, 8}) {
      k_fill_ids<<<sms * 8, 256>>>(idc, 10ll << 20, (64u << 20) / 128, 777);
      CK(cudaDeviceSynchronize());
      const float node = run_c2<1>(te, idc, own, 1ll << 20, sms * bps);
      k_fill_ids<<<sms * 8, 256>>>(idc, 10ll << 20, (128u << 20) / 128, 778);
      CK(cudaDeviceSynchronize());
      const float edge = run_c2<2>(tn, idc, own, 1ll << 19, sms * bps);
      printf("blocks/SM %d  node half (1M rows x 10 gathers, 64 MB table): %.4f ms   edge half (512K rows x 20 gathers, 128 MB table): %.4f ms\n",
             bps, node, edge);
    }
  }
}

int main(int argc, char** argv) {
  const bool only_c2 = argc > 1;   // any argument: the config-2 section only
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  if (only_c2) {
    c2_section(sms);
    return 0;
  }
  const long long total_gathers = 16ll << 20;            // 16M rows = 2 GB gathered per launch
  uint32_t* ids; CK(cudaMalloc(&ids, total_gathers * 4));
  float4* out; CK(cudaMalloc(&out, 64));
  const size_t sizes_mb[] = {8, 32, 48, 64, 80, 96, 128, 192, 256, 1024, 8192};
  printf("sms %d; gathered GB/s (rows of 128 B) by table size, blocks per SM and loads in flight per lane\n", sms);
  for (size_t mb : sizes_mb) {
    const uint32_t rows = (uint32_t)((mb << 20) / 128);
    float4* table; CK(cudaMalloc(&table, (size_t)rows * 128));
    CK(cudaMemset(table, 0, (size_t)rows * 128));
    k_fill_ids<<<sms * 8, 256>>>(ids, total_gathers, rows, mb * 7919);
    CK(cudaDeviceSynchronize());
    for (int bps : {4, 8}) {
      const int blocks = sms * bps;
      const long long per_warp = total_gathers / 4 / ((long long)blocks * 8) / 16 * 16;
      const double bytes = (double)per_warp * 4 * blocks * 8 * 128;
      const float t8 = run<8>(table, ids, per_warp, blocks, out);
      const float t16 = run<16>(table, ids, per_warp, blocks, out);
      printf("table %5zu MB  blocks/SM %d  U=8: %7.0f GB/s  U=16: %7.0f GB/s\n", mb, bps, bytes / t8 / 1e6,
             bytes / t16 / 1e6);
    }
    CK(cudaFree(table));
  }
  // ---- the mix of a half-sweep -------------------------------------------------------------
  printf("\nmix: per owned row 8 gathers from the table + read and write of the row itself (128 B each);\n"
         "GB/s counts gathered bytes only, ms is for 1M owned rows / 8M gathers (the node half of config 2\n"
         "has 1M rows and 10M gathers from a 64 MB table, the edge half 500K rows and 10M gathers from 128 MB)\n");
  {
    const long long own_rows = 1ll << 20;                 // 128 MB of owned rows
    float4* own; CK(cudaMalloc(&own, own_rows * 128));
    CK(cudaMemset(own, 0, own_rows * 128));
    uint32_t* order; CK(cudaMalloc(&order, own_rows * 4));
    // a random permutation-like order (collisions do not matter for traffic)
    k_fill_ids<<<sms * 8, 256>>>(order, own_rows, (uint32_t)own_rows, 12345);
    for (size_t mb : {(size_t)64, (size_t)128}) {
      const uint32_t rows = (uint32_t)((mb << 20) / 128);
      float4* table; CK(cudaMalloc(&table, (size_t)rows * 128));
      CK(cudaMemset(table, 0, (size_t)rows * 128));
      k_fill_ids<<<sms * 8, 256>>>(ids, own_rows * 8, rows, mb * 104729);
      CK(cudaDeviceSynchronize());
      for (int bps : {4, 8}) {
        const int blocks = sms * bps;
        const long long rpw = own_rows / 4 / ((long long)blocks * 8);
        for (int rnd = 0; rnd < 2; ++rnd) {
          cudaEvent_t a, b;
          CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
          k_mix<<<blocks, 256>>>(table, ids, own, order, rpw, rnd);
          CK(cudaEventRecord(a));
          for (int r = 0; r < 5; ++r) k_mix<<<blocks, 256>>>(table, ids, own, order, rpw, rnd);
          CK(cudaEventRecord(b));
          CK(cudaEventSynchronize(b));
          float ms; CK(cudaEventElapsedTime(&ms, a, b));
          ms /= 5;
          const double done_rows = (double)rpw * 4 * blocks * 8;
          printf("table %4zu MB  blocks/SM %d  owned rows %s: %.4f ms  (%.0f GB/s gathered, %.0f GB/s with the owned rows)\n",
                 mb, bps, rnd ? "random    " : "sequential", ms * (double)own_rows / done_rows,
                 done_rows * 8 * 128 / ms / 1e6, done_rows * 10 * 128 / ms / 1e6);
        }
      }
      for (int own_first = 0; own_first < 2; ++own_first) {
        const int blocks = sms * 4;
        const long long rpw = own_rows / 4 / ((long long)blocks * 8);
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        k_mix_hint<<<blocks, 256>>>(table, ids, own, rpw, own_first);
        CK(cudaEventRecord(a));
        for (int r = 0; r < 5; ++r) k_mix_hint<<<blocks, 256>>>(table, ids, own, rpw, own_first);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        ms /= 5;
        const double done_rows = (double)rpw * 4 * blocks * 8;
        printf("table %4zu MB  blocks/SM 4  table rows evict_last, owned rows %s: %.4f ms  (%.0f GB/s gathered)\n",
               mb, own_first ? "evict_first" : "default    ", ms * (double)own_rows / done_rows,
               done_rows * 8 * 128 / ms / 1e6);
      }
      CK(cudaFree(table));
    }
  }
  c2_section(sms);
  return 0;
}
