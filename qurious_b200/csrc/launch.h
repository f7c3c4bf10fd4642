// Kernel launch helper shared by every .cu file: counts launches (bench.py: gpu_launches) and, when
// profiling is enabled on the context (qgpu_profile_enable), brackets each launch with a pair of CUDA
// events on the launching stream so that bench.py can report per-kernel device time measured live.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

struct KernelScope {
  Ctx* ctx;
  int slot = -1;
  KernelScope(Ctx* c, const char* name) : ctx(c) {
    c->launches++;
    if (c->profiling) slot = c->prof_begin(name);
  }
  ~KernelScope() {
    if (slot >= 0) ctx->prof_end(slot);
  }
};

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                    \
  do {                                                                 \
    ::qgpu::KernelScope _ks((ctx), #kernel);                           \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
    CUDA_CHECK(cudaGetLastError());                                    \
  } while (0)

static inline int grid_for(Ctx* ctx, int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx->sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace qgpu
