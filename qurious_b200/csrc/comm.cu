// NCCL (dlopen) + CUDA-IPC symmetric buffers: see comm.h.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>

#include "comm.h"

namespace qgpu {

namespace {
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;

[[noreturn]] void throw_nccl(const std::string& m) { throw QError(QGPU_ERR_NCCL, "NcclError: " + m); }

NcclApi& nccl_api() {
  if (g_nccl.lib) return g_nccl;
  // SONAME lookup: inside a process that already loaded a libnccl.so.2 (torch ships one) this binds to that copy
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw_nccl(std::string("cannot load libnccl.so.2: ") + dlerror());
  NcclApi a;
  a.lib = h;
#define QSYM(name)                                                       \
  a.name = (decltype(a.name))dlsym(h, "nccl" #name);                     \
  if (!a.name) throw_nccl("libnccl.so.2 has no symbol nccl" #name)
  QSYM(GetUniqueId);
  QSYM(CommInitRank);
  QSYM(CommDestroy);
  QSYM(AllGather);
  QSYM(Send);
  QSYM(Recv);
  QSYM(GroupStart);
  QSYM(GroupEnd);
  QSYM(GetErrorString);
#undef QSYM
  g_nccl = a;
  return g_nccl;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw_nccl(std::string(what) + ": " + nccl_api().GetErrorString(r));
}

void alloc_symmetric(Comm& c) {
  c.bytes = COMM_FLAG_BYTES + 2 * (size_t)c.world * COMM_SLOT_BYTES;
  CUDA_CHECK(cudaMalloc(&c.base, c.bytes));  // plain cudaMalloc: exportable through CUDA IPC (pool memory is not)
  CUDA_CHECK(cudaMemsetAsync(c.base, 0, c.bytes, c.ctx->stream));
  c.peer[c.rank] = c.base;
}
}  // namespace

Comm::~Comm() {
  for (int r = 0; r < COMM_MAX_WORLD; ++r)
    if (opened[r] && peer[r]) cudaIpcCloseMemHandle(peer[r]);
  if (base) cudaFree(base);
  if (nccl) nccl_api().CommDestroy((ncclComm_t)nccl);
}

void comm_unique_id(void* out128) {
  ncclUniqueId id;
  nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
}

void comm_init(Ctx* ctx, const void* id128, int rank, int world) {
  if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world) throw_internal("qgpu_comm_init: world must be in [1, 8] and 0 <= rank < world");
  if (ctx->comm) throw_internal("qgpu_comm_init: this context already has a communicator");
  auto c = std::make_shared<Comm>();
  c->ctx = ctx;
  c->world = world;
  c->rank = rank;
  NcclApi& api = nccl_api();
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  nccl_check(api.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  c->nccl = comm;
  alloc_symmetric(*c);
  // exchange the IPC handles of the symmetric buffers through the new communicator
  struct Pub {
    cudaIpcMemHandle_t h;
    uint64_t pid, ptr;
  };
  Pub mine;
  memset(&mine, 0, sizeof(mine));
  CUDA_CHECK(cudaIpcGetMemHandle(&mine.h, c->base));
  mine.pid = (uint64_t)getpid();
  mine.ptr = (uint64_t)(uintptr_t)c->base;
  DBufP send = ctx->alloc(sizeof(Pub)), recv = ctx->alloc(sizeof(Pub) * (size_t)world);
  CUDA_CHECK(cudaMemcpyAsync(send->ptr, &mine, sizeof(Pub), cudaMemcpyHostToDevice, ctx->stream));
  nccl_check(api.AllGather(send->ptr, recv->ptr, sizeof(Pub), ncclUint8, comm, ctx->stream), "ncclAllGather(ipc handles)");
  std::vector<Pub> all((size_t)world);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), recv->ptr, sizeof(Pub) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    if (all[r].pid == mine.pid) {  // another context of this process: the pointer is directly usable
      c->peer[r] = (void*)(uintptr_t)all[r].ptr;
    } else {
      void* p = nullptr;
      CUDA_CHECK(cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess));
      c->peer[r] = p;
      c->opened[r] = true;
    }
  }
  ctx->comm = c;
  comm_barrier(ctx);  // every rank has zeroed its flags before anybody's first epoch can arrive
}

// In-process groups: the ranks are threads of one process.  A grouped exchange is a host rendezvous -- every rank posts its
// send list, waits for the others, copies what is addressed to it on its own stream, and nobody leaves before all copies
// are done (the send buffers stay alive).
struct LocalGroup {
  std::mutex mu;
  std::condition_variable cv;
  int world = 1, arrived = 0;
  unsigned long long gen = 0;
  std::vector<std::vector<Xfer>> sends;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long g = gen;
    if (++arrived == world) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g; })) {
      --arrived;
      throw QError(QGPU_ERR_NCCL, "NcclError: a rank of the in-process group never arrived at the exchange");
    }
  }
};

void comm_init_local(Ctx** ctxs, int n) {
  if (n < 1 || n > COMM_MAX_WORLD) throw_internal("qgpu_comm_init_local: 1 to 8 contexts");
  auto group = std::make_shared<LocalGroup>();
  group->world = n;
  group->sends.resize((size_t)n);
  std::vector<std::shared_ptr<Comm>> cs;
  for (int r = 0; r < n; ++r) {
    if (ctxs[r]->comm) throw_internal("qgpu_comm_init_local: context already has a communicator");
    auto c = std::make_shared<Comm>();
    c->ctx = ctxs[r];
    c->world = n;
    c->rank = r;
    c->local = true;
    c->group = group;
    CUDA_CHECK(cudaSetDevice(ctxs[r]->device));
    alloc_symmetric(*c);
    CUDA_CHECK(cudaStreamSynchronize(ctxs[r]->stream));
    cs.push_back(c);
  }
  for (int r = 0; r < n; ++r) {
    for (int q = 0; q < n; ++q) {
      cs[r]->peer[q] = cs[q]->base;
      if (q != r && ctxs[q]->device != ctxs[r]->device) {
        CUDA_CHECK(cudaSetDevice(ctxs[r]->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
        cudaGetLastError();
      }
    }
    ctxs[r]->comm = cs[r];
  }
}

int comm_world_size(Ctx* ctx) { return ctx->comm ? ctx->comm->world : 0; }

void comm_destroy(Ctx* ctx) {
  if (!ctx->comm) return;
  cudaStreamSynchronize(ctx->stream);
  ctx->comm.reset();
}

static Comm& need_nccl(Ctx* ctx) {
  if (!ctx->comm) throw_nccl("no communicator: call qgpu_comm_init first");
  if (!ctx->comm->nccl) throw_nccl("this communicator is an in-process group (qgpu_comm_init_local): it has no NCCL collectives");
  return *ctx->comm;
}

void comm_all_gather(Ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  Comm& c = need_nccl(ctx);
  nccl_check(nccl_api().AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)c.nccl, ctx->stream), "ncclAllGather");
}

void comm_all_to_all(Ctx* ctx, const void* send, const int64_t* send_off, const int64_t* send_bytes, void* recv, const int64_t* recv_off,
                     const int64_t* recv_bytes) {
  Comm& c = need_nccl(ctx);
  NcclApi& api = nccl_api();
  nccl_check(api.GroupStart(), "ncclGroupStart");
  for (int r = 0; r < c.world; ++r) {
    if (send_bytes[r] > 0)
      nccl_check(api.Send((const char*)send + send_off[r], (size_t)send_bytes[r], ncclUint8, r, (ncclComm_t)c.nccl, ctx->stream), "ncclSend");
    if (recv_bytes[r] > 0)
      nccl_check(api.Recv((char*)recv + recv_off[r], (size_t)recv_bytes[r], ncclUint8, r, (ncclComm_t)c.nccl, ctx->stream), "ncclRecv");
  }
  nccl_check(api.GroupEnd(), "ncclGroupEnd");
}

void comm_exchange(Ctx* ctx, const std::vector<Xfer>& sends, const std::vector<Xfer>& recvs) {
  if (!ctx->comm) throw_nccl("no communicator: call qgpu_comm_init first");
  Comm& c = *ctx->comm;
  if (!c.local) {
    NcclApi& api = nccl_api();
    nccl_check(api.GroupStart(), "ncclGroupStart");
    for (const Xfer& x : sends)
      if (x.bytes) nccl_check(api.Send(x.ptr, x.bytes, ncclUint8, x.peer, (ncclComm_t)c.nccl, ctx->stream), "ncclSend");
    for (const Xfer& x : recvs)
      if (x.bytes) nccl_check(api.Recv(x.ptr, x.bytes, ncclUint8, x.peer, (ncclComm_t)c.nccl, ctx->stream), "ncclRecv");
    nccl_check(api.GroupEnd(), "ncclGroupEnd");
    return;
  }
  LocalGroup& g = *c.group;
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // my send buffers are complete
  {
    std::lock_guard<std::mutex> lk(g.mu);
    g.sends[(size_t)c.rank] = sends;
  }
  g.barrier();
  std::vector<size_t> next((size_t)c.world, 0);  // per peer: how many of its sends addressed to me were consumed
  std::string err;
  for (const Xfer& x : recvs) {
    const std::vector<Xfer>& ps = g.sends[(size_t)x.peer];
    size_t& k = next[(size_t)x.peer];
    while (k < ps.size() && ps[k].peer != c.rank) ++k;
    if (k >= ps.size() || ps[k].bytes != x.bytes) {
      err = "grouped exchange: a receive has no matching send of the same size";
      break;
    }
    if (x.bytes && cudaMemcpyAsync(x.ptr, ps[k].ptr, x.bytes, cudaMemcpyDefault, ctx->stream) != cudaSuccess) {
      err = "grouped exchange: device copy failed";
      break;
    }
    ++k;
  }
  cudaStreamSynchronize(ctx->stream);
  g.barrier();  // every rank has copied: the send buffers may go
  if (!err.empty()) throw_nccl(err);
}

void comm_all_gather_any(Ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (!ctx->comm) throw_nccl("no communicator: call qgpu_comm_init first");
  Comm& c = *ctx->comm;
  if (!c.local) {
    comm_all_gather(ctx, send, recv, bytes_per_rank);
    return;
  }
  std::vector<Xfer> sends, recvs;
  for (int r = 0; r < c.world; ++r) {
    sends.push_back({r, (void*)send, bytes_per_rank});
    recvs.push_back({r, (char*)recv + (size_t)r * bytes_per_rank, bytes_per_rank});
  }
  comm_exchange(ctx, sends, recvs);
}

void comm_barrier(Ctx* ctx) {
  if (ctx->comm && ctx->comm->local) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->comm->group->barrier();
    return;
  }
  Comm& c = need_nccl(ctx);
  DBufP b = ctx->alloc_zero(8 * (size_t)(c.world + 1));
  nccl_check(nccl_api().AllGather(b->ptr, (char*)b->ptr + 8, 8, ncclUint8, (ncclComm_t)c.nccl, ctx->stream), "ncclAllGather(barrier)");
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace qgpu
