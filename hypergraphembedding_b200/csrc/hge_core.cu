// Context, error reporting and versioning of libhge_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "hge_common.cuh"

static thread_local char g_err[1024] = "";

void hge_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void* hge_ctx_pinned(hge_ctx* ctx, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  ctx->pinned_in_flight = true;
  // first chunk at or after the current one with room
  for (int k = ctx->pinned_cur; k < ctx->pinned_chunks; ++k) {
    const size_t off = (k == ctx->pinned_cur) ? ctx->pinned_off : 0;
    if (off + bytes <= ctx->pinned_size[k]) {
      ctx->pinned_cur = k;
      ctx->pinned_off = off + bytes;
      return static_cast<char*>(ctx->pinned_chunk[k]) + off;
    }
  }
  if (ctx->pinned_chunks == 32) {
    hge_set_error("pinned staging arena exhausted");
    return nullptr;
  }
  const size_t size = bytes > ((size_t)32 << 20) ? bytes : ((size_t)32 << 20);
  void* p = nullptr;
  if (cudaMallocHost(&p, size) != cudaSuccess) {
    cudaGetLastError();
    hge_set_error("cudaMallocHost of %zu bytes failed", size);
    return nullptr;
  }
  const int k = ctx->pinned_chunks++;
  ctx->pinned_chunk[k] = p;
  ctx->pinned_size[k] = size;
  ctx->pinned_cur = k;
  ctx->pinned_off = bytes;
  return p;
}

void hge_ctx_pinned_reset(hge_ctx* ctx) {
  if (ctx->pinned_in_flight) cudaStreamSynchronize(ctx->stream);
  ctx->pinned_in_flight = false;
  ctx->pinned_cur = 0;
  ctx->pinned_off = 0;
}

extern "C" {

int hge_version(void) { return 100; }

const char* hge_last_error(void) { return g_err; }

int hge_ctx_create(int device, void* stream, hge_ctx** out) {
  HGE_REQUIRE(out != nullptr, "hge_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    hge_set_error("hge_ctx_create: no CUDA device is usable (%s); this library has no CPU path",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return HGE_ERR_CUDA;
  }
  HGE_REQUIRE(device >= 0 && device < count, "hge_ctx_create: device %d out of range [0, %d)",
              device, count);
  HGE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HGE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    hge_set_error("hge_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                  device, prop.major, prop.minor);
    return HGE_ERR_UNSUPPORTED;
  }
  hge_ctx* ctx = new (std::nothrow) hge_ctx();
  if (!ctx) return HGE_ERR_NOMEM;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->own_stream = false;
  ctx->light_max_deg = 64;
  ctx->chunk = 512;
  ctx->blocks_per_sm = 0;  // 0: ask the occupancy calculator
  // measured slower than the register gather (profiles/r1_bulk_copy_experiment.md): opt-in
  ctx->use_bulk = 0;
  if (const char* env = getenv("HGE_BULK")) ctx->use_bulk = atoi(env) != 0;
  ctx->launches = 0;
  ctx->pinned_chunks = 0;
  ctx->pinned_cur = 0;
  ctx->pinned_off = 0;
  ctx->pinned_in_flight = false;
  // NULL selects the legacy default stream, which is also torch's default stream, so work
  // queued by the caller on that stream is ordered with ours.
  ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  // keep freed blocks in the stream-ordered pool: the per-call workspaces are re-used
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  cudaGetLastError();
  *out = ctx;
  return HGE_OK;
}

int hge_ctx_destroy(hge_ctx* ctx) {
  if (!ctx) return HGE_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < ctx->pinned_chunks; ++k) cudaFreeHost(ctx->pinned_chunk[k]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return HGE_OK;
}

int hge_ctx_set_stream(hge_ctx* ctx, void* stream) {
  HGE_REQUIRE(ctx != nullptr, "hge_ctx_set_stream: ctx is NULL");
  cudaStream_t next = reinterpret_cast<cudaStream_t>(stream);
  if (next != ctx->stream) {
    // allocations and frees are ordered on the context's stream: drain it before switching
    HGE_CUDA(cudaSetDevice(ctx->device));
    HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  ctx->stream = next;
  return HGE_OK;
}

int hge_ctx_sync(hge_ctx* ctx) {
  HGE_REQUIRE(ctx != nullptr, "hge_ctx_sync: ctx is NULL");
  HGE_CUDA(cudaSetDevice(ctx->device));
  HGE_CUDA(cudaStreamSynchronize(ctx->stream));
  return HGE_OK;
}

int hge_ctx_set_tuning(hge_ctx* ctx, int light_max_deg, int chunk, int blocks_per_sm) {
  HGE_REQUIRE(ctx != nullptr, "hge_ctx_set_tuning: ctx is NULL");
  HGE_REQUIRE(light_max_deg >= 0 && light_max_deg <= 255,
              "hge_ctx_set_tuning: light_max_deg %d not in [0, 255]", light_max_deg);
  HGE_REQUIRE(chunk >= 0 && (chunk == 0 || chunk % 32 == 0),
              "hge_ctx_set_tuning: chunk %d must be a multiple of 32", chunk);
  HGE_REQUIRE(blocks_per_sm >= 0 && blocks_per_sm <= 32,
              "hge_ctx_set_tuning: blocks_per_sm %d not in [0, 32]", blocks_per_sm);
  if (light_max_deg) ctx->light_max_deg = light_max_deg;
  if (chunk) ctx->chunk = chunk;
  ctx->blocks_per_sm = blocks_per_sm;
  return HGE_OK;
}

int hge_ctx_set_bulk(hge_ctx* ctx, int enabled) {
  HGE_REQUIRE(ctx != nullptr, "hge_ctx_set_bulk: ctx is NULL");
  ctx->use_bulk = enabled != 0;
  return HGE_OK;
}

int64_t hge_ctx_launch_count(const hge_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
