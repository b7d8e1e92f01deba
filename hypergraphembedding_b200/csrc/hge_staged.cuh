// Host <-> device staging for the HGE_MEM_HOST flavour of the entry points.
#pragma once

#include "hge_common.cuh"

template <typename T>
struct Staged {
  const hge_ctx* ctx = nullptr;
  T* dev = nullptr;
  T* host = nullptr;
  size_t count = 0;
  bool owned = false;
  bool copy_out = false;

  int init(const hge_ctx* c, const T* p, size_t n, int mem, bool in, bool out) {
    ctx = c;
    count = n;
    copy_out = out && mem == HGE_MEM_HOST;
    host = const_cast<T*>(p);
    if (mem == HGE_MEM_DEVICE || p == nullptr) {
      dev = const_cast<T*>(p);
      return HGE_OK;
    }
    owned = true;
    HGE_TRY(hge_dev_alloc(ctx, &dev, n));
    if (in && n)
      HGE_CUDA(cudaMemcpyAsync(dev, p, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return HGE_OK;
  }
  int finish() {
    if (copy_out && count) {
      HGE_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
      HGE_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return HGE_OK;
  }
  ~Staged() {
    if (owned) hge_dev_free(ctx, dev);
  }
};

