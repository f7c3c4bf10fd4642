// Run-time specialisation of the fused DENSE scan-aggregate kernel (fused_device.cuh) with NVRTC.
//
// The hand-written kernel has two tile bodies: a generic one that interprets the plan structure per tile, and
// `SpecBody<signature>`, the same pipeline with the plan's SHAPE (operand widths, accumulator kinds / chaining / factor
// counts, accumulators packed into the count word) as template parameters -- constants stay run-time parameters.  Two
// shapes (TPC-H Q1, Q6) are instantiated ahead of time.  For every other DENSE plan this file instantiates the SAME source
// for that plan's signature at run time: nvrtcCompileProgram(fused_device.cuh + one __global__ wrapper) -> CUBIN for
// sm_100a -> cuModuleLoadData -> cuLaunchKernel.  One compilation per (signature, packing) and process (~1 s), cached.
// libnvrtc.so.12 and libcuda.so.1 are dlopen'ed on first use; if either is missing, or the compilation fails, the plan
// simply runs the generic body (QGPU_JIT=0 forces that; QGPU_JIT_DEBUG=1 prints the compile log).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>

#include "fused_jit.h"

namespace qgpu {

namespace {

const char* kDeviceSource =
#include "build/fused_device_src.inc"
    ;

struct Api {
  bool tried = false, ok = false;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
  decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
  decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
  decltype(&cuLaunchKernel) LaunchKernel = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
};
Api g_api;
std::mutex g_mu;
struct Key {
  uint64_t s[4];
  uint32_t pack;
  int device;
  bool operator<(const Key& o) const {
    for (int i = 0; i < 4; ++i)
      if (s[i] != o.s[i]) return s[i] < o.s[i];
    if (pack != o.pack) return pack < o.pack;
    return device < o.device;
  }
};
std::map<Key, JitKernel> g_cache;

bool debug() {
  static const bool d = getenv("QGPU_JIT_DEBUG") != nullptr;
  return d;
}

Api& api() {
  if (g_api.tried) return g_api;
  g_api.tried = true;
  void* rtc = dlopen("libnvrtc.so.12", RTLD_NOW | RTLD_GLOBAL);
  if (!rtc) rtc = dlopen("/usr/local/cuda/lib64/libnvrtc.so.12", RTLD_NOW | RTLD_GLOBAL);
  if (!rtc) rtc = dlopen("libnvrtc.so", RTLD_NOW | RTLD_GLOBAL);
  void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (debug() && (!rtc || !drv)) fprintf(stderr, "[qgpu jit] %s not loadable: %s\n", rtc ? "libcuda.so.1" : "libnvrtc.so.12", dlerror());
  bool ok = true;
#define QSYM(lib, field, name)                                           \
  g_api.field = lib ? (decltype(g_api.field))dlsym(lib, name) : nullptr; \
  ok = ok && g_api.field != nullptr
  QSYM(rtc, CreateProgram, "nvrtcCreateProgram");
  QSYM(rtc, CompileProgram, "nvrtcCompileProgram");
  QSYM(rtc, GetCUBINSize, "nvrtcGetCUBINSize");
  QSYM(rtc, GetCUBIN, "nvrtcGetCUBIN");
  QSYM(rtc, GetProgramLogSize, "nvrtcGetProgramLogSize");
  QSYM(rtc, GetProgramLog, "nvrtcGetProgramLog");
  QSYM(rtc, DestroyProgram, "nvrtcDestroyProgram");
  QSYM(drv, ModuleLoadData, "cuModuleLoadData");
  QSYM(drv, ModuleGetFunction, "cuModuleGetFunction");
  QSYM(drv, FuncSetAttribute, "cuFuncSetAttribute");
  QSYM(drv, LaunchKernel, "cuLaunchKernel");
  QSYM(drv, GetErrorString, "cuGetErrorString");
#undef QSYM
  g_api.ok = ok;
  return g_api;
}

std::string hex(uint64_t v) {
  char b[32];
  snprintf(b, sizeof(b), "0x%llxULL", (unsigned long long)v);
  return b;
}

}  // namespace

// CUBIN of the kernel specialised for (signature, pack); empty on failure.  Needs no GPU: the build check and the CPU-side
// tests call this too.
std::string jit_compile_cubin(const uint64_t sig[4], uint32_t pack, std::string* log_out) {
  Api& a = api();
  if (!a.CreateProgram || !a.CompileProgram || !a.GetCUBIN) {
    if (log_out) *log_out = "libnvrtc.so.12 is not loadable";
    return std::string();
  }
  std::string src = kDeviceSource;
  src += "\nextern \"C\" __global__ void __launch_bounds__(qgpu::F_NT + 32, 1) k_fused_scan_agg_jit(const __grid_constant__ qgpu::FParams p) {\n";
  src += "  qgpu::fused_main<qgpu::FM_DENSE, qgpu::SpecBody<" + hex(sig[0]) + ", " + hex(sig[1]) + ", " + hex(sig[2]) + ", " + hex(sig[3]) + ", " +
         std::to_string(pack) + "u>>(p);\n}\n";
  nvrtcProgram prog = nullptr;
  if (a.CreateProgram(&prog, src.c_str(), "fused_device_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return std::string();
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-DQGPU_JIT", "-device-int128", "-default-device", "-lineinfo"};
  const nvrtcResult rc = a.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
  size_t ls = 0;
  a.GetProgramLogSize(prog, &ls);
  std::string log(ls, '\0');
  if (ls > 1) a.GetProgramLog(prog, &log[0]);
  if (log_out) *log_out = log;
  std::string cubin;
  if (rc == NVRTC_SUCCESS) {
    size_t n = 0;
    if (a.GetCUBINSize(prog, &n) == NVRTC_SUCCESS && n > 0) {
      cubin.resize(n);
      if (a.GetCUBIN(prog, &cubin[0]) != NVRTC_SUCCESS) cubin.clear();
    }
  } else if (debug()) {
    fprintf(stderr, "[qgpu jit] compilation failed:\n%s\n", log.c_str());
  }
  a.DestroyProgram(&prog);
  return cubin;
}

JitKernel jit_specialised_dense(int device, const uint64_t sig[4], uint32_t pack) {
  JitKernel none;
  const char* e = getenv("QGPU_JIT");
  if (e && e[0] == '0') return none;
  std::lock_guard<std::mutex> lk(g_mu);
  Key k{{sig[0], sig[1], sig[2], sig[3]}, pack, device};
  auto it = g_cache.find(k);
  if (it != g_cache.end()) return it->second;
  JitKernel out;
  Api& a = api();
  if (a.ok) {
    std::string log;
    const std::string cubin = jit_compile_cubin(sig, pack, &log);
    if (!cubin.empty()) {
      CUmodule mod = nullptr;
      CUfunction fn = nullptr;
      CUresult r = a.ModuleLoadData(&mod, cubin.data());
      if (r == CUDA_SUCCESS) r = a.ModuleGetFunction(&fn, mod, "k_fused_scan_agg_jit");
      if (r == CUDA_SUCCESS) {
        out.fn = (void*)fn;
      } else if (debug()) {
        const char* m = nullptr;
        a.GetErrorString(r, &m);
        fprintf(stderr, "[qgpu jit] loading the compiled kernel failed: %s\n", m ? m : "?");
      }
    }
  }
  g_cache[k] = out;  // failures are cached too: the generic body runs, no second attempt per execution
  return out;
}

void jit_launch(const JitKernel& k, int grid, int block, size_t smem_bytes, cudaStream_t stream, const void* params) {
  Api& a = api();
  CUfunction fn = (CUfunction)k.fn;
  CUresult r = a.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes);
  void* args[] = {(void*)params};
  if (r == CUDA_SUCCESS) r = a.LaunchKernel(fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, (unsigned)smem_bytes, (CUstream)stream, args, nullptr);
  if (r != CUDA_SUCCESS) {
    const char* m = nullptr;
    a.GetErrorString(r, &m);
    throw QError(QGPU_ERR_CUDA, std::string("CUDA error launching the run-time specialised kernel: ") + (m ? m : "?"));
  }
}

}  // namespace qgpu
