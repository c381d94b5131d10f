// Multi-GPU plumbing behind the C ABI: the path shards by page with no data-path collective, so
// the library needs exactly two NCCL operations -- ONE broadcast of the converted weights into place at start-up and
// ONE all-gather of the fixed-stride decoded-sequence buffers per batch -- both enqueued on the context's stream.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (e.g. through
// torch.distributed) shares it, a single-GPU user never needs it, and a missing library is a loud error, not a
// fallback.  Only the long-stable entry points are used.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "alm_internal.h"

namespace alm {

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      err = std::string("NCCL is not loadable (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [](const char* n) {
      void* p = dlsym(api.lib, n);
      if (!p) err = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!err.empty()) throw AlmError{ALM_ERR_UNSUPPORTED, err};
  return api;
}

#define ALM_CHECK_NCCL(expr)                                                                                \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess)                                                                                  \
      throw alm::AlmError{ALM_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r)};              \
  } while (0)

}  // namespace

void comm_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ALM_CHECK_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
}

void comm_init(Ctx* c, const void* id128, int rank, int world) {
  ALM_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, ALM_ERR_INVALID, "alm_comm_init arguments");
  comm_release(c);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ALM_CHECK_NCCL(nccl().CommInitRank(&comm, world, id, rank));
  c->comm = comm;
  c->own_comm = true;
  c->comm_rank = rank;
  c->comm_world = world;
}

void comm_attach(Ctx* c, void* nccl_comm, int rank, int world) {
  ALM_REQUIRE(nccl_comm && world >= 1 && rank >= 0 && rank < world, ALM_ERR_INVALID, "alm_comm_attach arguments");
  nccl();  // the entry points must resolve in this process
  comm_release(c);
  c->comm = nccl_comm;
  c->own_comm = false;
  c->comm_rank = rank;
  c->comm_world = world;
}

void comm_release(Ctx* c) {
  if (c->comm && c->own_comm) {
    cudaStreamSynchronize(c->stream);
    nccl().CommDestroy(static_cast<ncclComm_t>(c->comm));
  }
  c->comm = nullptr;
  c->own_comm = false;
  c->comm_rank = 0;
  c->comm_world = 1;
}

// One broadcast of the converted weights (bf16 hi/lo planes, fp32 vectors: the slabs as they sit in HBM) from `root`
// into the same slabs of every other rank.  Precondition: every rank ran alm_load_weights with the same tensor names
// and shapes (placeholders on the non-root ranks), so the bump-allocated layout is identical; checked by size.
void comm_broadcast_weights(Ctx* c, int root) {
  ALM_REQUIRE(c->wstore && !c->wstore->slabs.empty(), ALM_ERR_STATE, "alm_broadcast_weights before alm_load_weights");
  if (c->comm == nullptr && c->comm_world == 1) return;  // single-GPU job without a communicator: nothing to do
  ALM_REQUIRE(c->comm != nullptr, ALM_ERR_STATE, "alm_broadcast_weights without a communicator (alm_comm_init)");
  ALM_REQUIRE(root >= 0 && root < c->comm_world, ALM_ERR_INVALID, "broadcast root");
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  // layout check: [n_slabs, used bytes of each slab] must agree with the root's
  const std::vector<size_t>& used = c->wstore->used;
  ALM_REQUIRE(used.size() == c->wstore->slabs.size() && used.size() <= 62, ALM_ERR_STATE, "weight slab bookkeeping");
  std::vector<unsigned long long> sig(64, 0), mine(64, 0);
  mine[0] = used.size();
  for (size_t i = 0; i < used.size(); ++i) mine[1 + i] = used[i];
  sig = mine;
  unsigned long long* dsig = nullptr;
  ALM_CHECK_CUDA(cudaMalloc(&dsig, 64 * sizeof(unsigned long long)));
  ALM_CHECK_CUDA(cudaMemcpyAsync(dsig, sig.data(), 64 * 8, cudaMemcpyHostToDevice, c->stream));
  ALM_CHECK_NCCL(nccl().Broadcast(dsig, dsig, 64 * 8, ncclUint8, root, comm, c->stream));
  ALM_CHECK_CUDA(cudaMemcpyAsync(sig.data(), dsig, 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dsig);
  ALM_REQUIRE(sig == mine, ALM_ERR_STATE,
              "weight layout differs from the root rank's (load the same tensor names / shapes on every rank first)");
  ALM_CHECK_NCCL(nccl().GroupStart());
  for (size_t i = 0; i < used.size(); ++i)
    ALM_CHECK_NCCL(nccl().Broadcast(c->wstore->slabs[i], c->wstore->slabs[i], used[i], ncclUint8, root, comm, c->stream));
  ALM_CHECK_NCCL(nccl().GroupEnd());
  ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
}

// All-gather of one fixed-size buffer per rank (the decoded sequences of a batch: int32 ids + fp32 probabilities packed
// by the caller) on the context's stream.  send: host or device, `bytes` per rank; recv: host buffer of
// world * bytes (rank-major) or NULL on ranks that do not need the result.
void comm_gather(Ctx* c, const void* send, size_t bytes, void* recv_host) {
  ALM_REQUIRE(send && bytes > 0, ALM_ERR_INVALID, "alm_gather_sequences arguments");
  const int world = c->comm_world;
  if (world == 1 && c->comm == nullptr) {
    if (recv_host) {
      cudaPointerAttributes a;
      const bool dev = cudaPointerGetAttributes(&a, send) == cudaSuccess && a.type == cudaMemoryTypeDevice;
      cudaGetLastError();
      if (dev) {
        ALM_CHECK_CUDA(cudaMemcpyAsync(recv_host, send, bytes, cudaMemcpyDeviceToHost, c->stream));
        ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
      } else {
        memcpy(recv_host, send, bytes);
      }
    }
    return;
  }
  ALM_REQUIRE(c->comm != nullptr, ALM_ERR_STATE, "alm_gather_sequences without a communicator (alm_comm_init)");
  if (c->gather_cap < bytes * (world + 1)) {
    if (c->gather_buf) cudaFree(c->gather_buf);
    c->gather_buf = nullptr;
    c->gather_cap = 0;
    ALM_CHECK_CUDA(cudaMalloc(&c->gather_buf, bytes * (world + 1)));
    c->gather_cap = bytes * (world + 1);
  }
  char* dsend = static_cast<char*>(c->gather_buf);
  char* drecv = dsend + bytes;
  ALM_CHECK_CUDA(cudaMemcpyAsync(dsend, send, bytes, cudaMemcpyDefault, c->stream));
  ALM_CHECK_NCCL(nccl().AllGather(dsend, drecv, bytes, ncclUint8, static_cast<ncclComm_t>(c->comm), c->stream));
  if (recv_host) ALM_CHECK_CUDA(cudaMemcpyAsync(recv_host, drecv, bytes * world, cudaMemcpyDeviceToHost, c->stream));
  ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
}

}  // namespace alm
