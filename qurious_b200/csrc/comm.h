// Multi-GPU plumbing BELOW the C ABI: one qgpu_ctx per process and GPU.
//   * an NCCL communicator owned by the library (libnccl.so.2 is dlopen'ed on qgpu_comm_init: a single-GPU host never
//     needs it), used for bootstrap exchanges and for the plain collectives of the generic paths;
//   * a SYMMETRIC peer buffer per rank (cudaMalloc, exported through CUDA IPC and mapped by every peer over
//     NVLink / NVSwitch): the fused "aggregate -> exchange -> merge" epilogue kernel (epilogue.cu) stores its state
//     block straight into every peer's buffer and synchronises through epoch flags in the same buffers -- no
//     collective call and no host round trip on the per-step path.
// The reference is single-process: nothing here mirrors a reference file.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

constexpr int COMM_MAX_WORLD = 8;
constexpr size_t COMM_FLAG_BYTES = 4096;               // flags[2][COMM_MAX_WORLD] u64 (+ padding)
constexpr size_t COMM_SLOT_BYTES = (size_t)1 << 20;    // one state block per (epoch parity, source rank)

struct Xfer {  // one point-to-point transfer of a grouped exchange
  int peer;
  void* ptr;
  size_t bytes;
};
struct LocalGroup;  // rendezvous board of an in-process group (comm.cu)

struct Comm {
  Ctx* ctx = nullptr;
  std::shared_ptr<LocalGroup> group;  // local == true: host rendezvous + device-to-device copies instead of NCCL
  int world = 1, rank = 0;
  bool local = false;          // ranks of ONE process wired together without NCCL (tests on a single GPU)
  void* nccl = nullptr;        // ncclComm_t
  void* base = nullptr;        // this rank's symmetric buffer: flags | slots[2][world]
  size_t bytes = 0;
  void* peer[COMM_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool opened[COMM_MAX_WORLD] = {false, false, false, false, false, false, false, false};
  unsigned long long epoch = 0;  // one per fused collective; every rank runs the same sequence (SPMD)
  ~Comm();
};

void comm_unique_id(void* out128);
void comm_init(Ctx* ctx, const void* id128, int rank, int world);
void comm_init_local(Ctx** ctxs, int n);
void comm_destroy(Ctx* ctx);
// NCCL collectives on ctx->stream (device buffers); QGPU_ERR_NCCL on failure or when no communicator exists
void comm_all_gather(Ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
void comm_all_to_all(Ctx* ctx, const void* send, const int64_t* send_off, const int64_t* send_bytes, void* recv, const int64_t* recv_off,
                     const int64_t* recv_bytes);
void comm_barrier(Ctx* ctx);  // all ranks have reached this point of their streams (tiny all-gather + stream sync)
// One grouped exchange on ctx->stream: every rank's k-th send to peer p pairs with p's k-th receive from that rank (self
// transfers included).  NCCL: one ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd, asynchronous.  In-process groups
// (qgpu_comm_init_local, one thread per rank): host rendezvous + cudaMemcpyAsync, returns after the copies completed.
void comm_exchange(Ctx* ctx, const std::vector<Xfer>& sends, const std::vector<Xfer>& recvs);
// fixed-size all-gather that also works on in-process groups
void comm_all_gather_any(Ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);

}  // namespace qgpu
