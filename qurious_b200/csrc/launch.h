// Kernel launch helper shared by every .cu file: counts launches (bench.py: gpu_launches) and, when
// profiling is enabled on the context (qgpu_profile_enable), brackets each launch with a pair of CUDA
// events on the launching stream so that bench.py can report per-kernel device time measured live.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

static inline long long grid_blocks(long long g) { return g; }
static inline long long grid_blocks(const dim3& g) { return (long long)g.x * g.y * g.z; }

struct KernelScope {
  Ctx* ctx;
  int slot = -1;
  // blocks < 0: unknown grid, always instrumented.  prof_min_blocks > 0 restricts the event pairs to the large-grid
  // launches (the kernels a roofline is quoted for): bracketing every few-microsecond single-block helper of a
  // 0.45 ms step with two event records costs the step 7 %.
  KernelScope(Ctx* c, const char* name, long long blocks = -1) : ctx(c) {
    c->launches++;
    if (c->profiling && (blocks < 0 || blocks >= c->prof_min_blocks)) slot = c->prof_begin(name);
  }
  ~KernelScope() {
    if (slot >= 0) ctx->prof_end(slot);
  }
};

// QGPU_SYNC_LAUNCH=1 (debug aid): synchronise after every launch and name the kernel that faulted
void debug_sync_launch(Ctx* ctx, const char* name);

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                    \
  do {                                                                 \
    ::qgpu::KernelScope _ks((ctx), #kernel, ::qgpu::grid_blocks(grid)); \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
    CUDA_CHECK(cudaGetLastError());                                    \
    ::qgpu::debug_sync_launch((ctx), #kernel);                         \
  } while (0)

static inline int grid_for(Ctx* ctx, int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx->sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace qgpu
