// Run-time specialisation (NVRTC) of the fused DENSE scan-aggregate kernel: see fused_jit.cu.
#pragma once
#include <string>

#include "qgpu_internal.h"

namespace qgpu {

struct JitKernel {
  void* fn = nullptr;  // CUfunction; null: not available (the generic body runs)
  explicit operator bool() const { return fn != nullptr; }
};

// the kernel specialised for this shape signature (FSig::s) and packing on `device`; compiled once per process
JitKernel jit_specialised_dense(int device, const uint64_t sig[4], uint32_t pack);
// <<<grid, block, smem_bytes, stream>>>(FParams): `params` points at the FParams block
void jit_launch(const JitKernel& k, int grid, int block, size_t smem_bytes, cudaStream_t stream, const void* params);
// compile only (no GPU needed): the CUBIN, empty on failure; log_out receives the compiler log
std::string jit_compile_cubin(const uint64_t sig[4], uint32_t pack, std::string* log_out);

}  // namespace qgpu
