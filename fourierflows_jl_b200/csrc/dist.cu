// NCCL plumbing for the slab-decomposed transforms.  The reference is single-device (README.md:58,
// docs/src/gpu.md:57): there is no reference counterpart; this is new capability named by BASELINE.json north_star.
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <mutex>
#include <vector>
#include "dist.h"
#include "ffb_common.cuh"

namespace ffb {

// minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, ncclResult_t = int, ncclDataType ncclInt8 = 0
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, char[128], int) = nullptr;  // id passed BY VALUE (struct of 128 bytes)
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;

struct UniqueId { char internal[128]; };
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);

static int load_nccl() {
  if (g_nccl.h) return FFB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }   // reuse the one torch loaded
  if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return set_error(FFB_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, char[128], int))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  g_nccl.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return set_error(FFB_ENCCL, "libnccl is missing required symbols");
  g_nccl.h = h;
  return FFB_OK;
}

#define FFB_NCCL(call)                                                                                           \
  do {                                                                                                           \
    int _r = (call);                                                                                             \
    if (_r != 0) return set_error(FFB_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
  } while (0)

cudaEvent_t dist_next_event(ffb_dist* d) {
  cudaEvent_t e = d->ev[d->ev_next];
  d->ev_next = (d->ev_next + 1) % ffb_dist::kEvents;
  return e;
}

int dist_barrier(ffb_dist* d, cudaStream_t st) {
  if (d->nranks == 1) return FFB_OK;
  FFB_REQUIRE(g_nccl.AllReduce, FFB_ENCCL, "ncclAllReduce not available");
  FFB_NCCL(g_nccl.AllReduce(d->barrier_buf, d->barrier_buf, 1, /*ncclFloat32*/ 7, /*ncclSum*/ 0, d->comm, st));
  return FFB_OK;
}

int dist_alltoall_bytes(ffb_dist* d, const void* sendbuf, void* recvbuf, size_t count, size_t stride, cudaStream_t st) {
  const char* sb = reinterpret_cast<const char*>(sendbuf);
  char* rb = reinterpret_cast<char*>(recvbuf);
  FFB_CUDA(cudaMemcpyAsync(rb + (size_t)d->rank * stride, sb + (size_t)d->rank * stride, count, cudaMemcpyDeviceToDevice, st));
  if (d->nranks == 1) return FFB_OK;
  FFB_NCCL(g_nccl.GroupStart());
  for (int i = 1; i < d->nranks; ++i) {
    const int to = (d->rank + i) % d->nranks, from = (d->rank - i + d->nranks) % d->nranks;
    FFB_NCCL(g_nccl.Send(sb + (size_t)to * stride, count, /*ncclInt8*/ 0, to, d->comm, st));
    FFB_NCCL(g_nccl.Recv(rb + (size_t)from * stride, count, /*ncclInt8*/ 0, from, d->comm, st));
  }
  FFB_NCCL(g_nccl.GroupEnd());
  return FFB_OK;
}

int dist_push_blocks(ffb_dist* d, const void* sendbuf, void* const* peer_bufs, size_t dst_off, size_t count, size_t stride, cudaStream_t after) {
  const char* sb = reinterpret_cast<const char*>(sendbuf);
  if (after) {
    cudaEvent_t e = dist_next_event(d);
    FFB_CUDA(cudaEventRecord(e, after));
    for (int k = 0; k < d->ncopy; ++k) FFB_CUDA(cudaStreamWaitEvent(d->copy_streams[k], e, 0));
  }
  // staggered destinations: at step i every rank writes to a different peer
  for (int i = 0; i < d->nranks; ++i) {
    const int to = (d->rank + i) % d->nranks;
    FFB_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(peer_bufs[to]) + dst_off, sb + (size_t)to * stride, count, cudaMemcpyDeviceToDevice,
                             d->copy_streams[i % d->ncopy]));
  }
  for (int k = 0; k < d->ncopy; ++k) {
    cudaEvent_t e = dist_next_event(d);
    FFB_CUDA(cudaEventRecord(e, d->copy_streams[k]));
    FFB_CUDA(cudaStreamWaitEvent(d->comm_stream, e, 0));
  }
  return FFB_OK;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_dist_unique_id(void* id128) {
  FFB_REQUIRE(id128, FFB_EINVAL, "id buffer is NULL");
  int rc = load_nccl(); if (rc) return rc;
  FFB_NCCL(g_nccl.GetUniqueId(id128));
  return FFB_OK;
}

int ffb_dist_init(ffb_dist** out, int rank, int nranks, const void* id128) {
  FFB_REQUIRE(out && id128, FFB_EINVAL, "NULL argument");
  FFB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, FFB_EINVAL, "bad rank %d of %d", rank, nranks);
  *out = nullptr;
  int rc = load_nccl(); if (rc) return rc;
  auto* d = new ffb_dist();
  d->rank = rank; d->nranks = nranks; d->comm = nullptr; d->ev_next = 0;
  UniqueId id;
  memcpy(id.internal, id128, 128);
  int r = reinterpret_cast<CommInitRankFn>(g_nccl.CommInitRank)(&d->comm, nranks, id, rank);
  if (r != 0) { delete d; return set_error(FFB_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); }
  // the communication stream carries the arrival barriers (tiny NCCL kernels) and the NCCL exchange: highest priority so that
  // its blocks are placed as soon as a compute CTA retires instead of after the running pass has drained
  int prio_lo = 0, prio_hi = 0;
  FFB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  FFB_CUDA(cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, prio_hi));
  for (auto& e : d->ev) FFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  d->ncopy = nranks < 8 ? (nranks > 1 ? nranks : 1) : 8;   // one stream (copy engine) per destination
  if (const char* e = getenv("FFB_COPY_STREAMS")) { const int v = atoi(e); if (v >= 1 && v <= 8) d->ncopy = v; }
  for (auto& cs : d->copy_streams) FFB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  FFB_CUDA(cudaMalloc(&d->barrier_buf, 256));
  FFB_CUDA(cudaMemset(d->barrier_buf, 0, 256));
  *out = d;
  return FFB_OK;
}

int ffb_dist_destroy(ffb_dist* d) {
  if (!d) return FFB_OK;
  cudaStreamSynchronize(d->comm_stream);
  if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
  for (auto& e : d->ev) cudaEventDestroy(e);
  for (auto& cs : d->copy_streams) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
  cudaStreamDestroy(d->comm_stream);
  cudaFree(d->barrier_buf);
  delete d;
  return FFB_OK;
}

int ffb_dist_info(const ffb_dist* d, int* rank, int* nranks) {
  FFB_REQUIRE(d, FFB_EINVAL, "dist is NULL");
  if (rank) *rank = d->rank;
  if (nranks) *nranks = d->nranks;
  return FFB_OK;
}

// CUDA IPC plumbing for the exchanges through peer memory: a rank exports the handle of the ALLOCATION that holds one of its
// buffers (64 bytes) plus the buffer's offset inside it (cudaMalloc sub-allocates small requests from shared blocks), the
// launcher moves both to the other ranks, which map the allocation once (cache keyed by the handle bytes, reference counted;
// opening the same allocation twice in one process fails with "resource already mapped") and add the offset.
namespace {
struct IpcMapping { char handle[64]; char* base; int refs; };
std::vector<IpcMapping> g_ipc;
std::mutex g_ipc_mu;
typedef int (*MemGetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);
}  // namespace

int ffb_dist_ipc_export(void* dev_ptr, void* host_handle64, size_t* offset) {
  FFB_REQUIRE(dev_ptr && host_handle64 && offset, FFB_EINVAL, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  char* base = reinterpret_cast<char*>(dev_ptr);
  static MemGetAddressRangeFn get_range = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      get_range = reinterpret_cast<MemGetAddressRangeFn>(fn);
    else cudaGetLastError();
    looked_up = true;
  }
  if (get_range) {
    unsigned long long b = 0; size_t sz = 0;
    if (get_range(&b, &sz, (unsigned long long)(uintptr_t)dev_ptr) == 0 && b) base = reinterpret_cast<char*>((uintptr_t)b);
  }
  cudaIpcMemHandle_t h;
  FFB_CUDA(cudaIpcGetMemHandle(&h, base));
  memcpy(host_handle64, &h, 64);
  *offset = (size_t)(reinterpret_cast<char*>(dev_ptr) - base);
  return FFB_OK;
}

int ffb_dist_ipc_open(const void* host_handle64, size_t offset, void** dev_ptr) {
  FFB_REQUIRE(dev_ptr && host_handle64, FFB_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (auto& m : g_ipc)
    if (memcmp(m.handle, host_handle64, 64) == 0) { ++m.refs; *dev_ptr = m.base + offset; return FFB_OK; }
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle64, 64);
  void* base = nullptr;
  FFB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  IpcMapping m;
  memcpy(m.handle, host_handle64, 64);
  m.base = reinterpret_cast<char*>(base); m.refs = 1;
  g_ipc.push_back(m);
  *dev_ptr = m.base + offset;
  return FFB_OK;
}

int ffb_dist_ipc_close(void* dev_ptr) {
  if (!dev_ptr) return FFB_OK;
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  int best = -1;   // the mapping with the greatest base <= dev_ptr (mapped allocations do not overlap)
  for (int i = 0; i < (int)g_ipc.size(); ++i)
    if (g_ipc[i].base <= reinterpret_cast<char*>(dev_ptr) && (best < 0 || g_ipc[i].base > g_ipc[best].base)) best = i;
  FFB_REQUIRE(best >= 0, FFB_EINVAL, "pointer was not returned by ffb_dist_ipc_open");
  if (--g_ipc[best].refs == 0) {
    FFB_CUDA(cudaIpcCloseMemHandle(g_ipc[best].base));
    g_ipc.erase(g_ipc.begin() + best);
  }
  return FFB_OK;
}

int ffb_dist_barrier(ffb_dist* d) {
  FFB_REQUIRE(d, FFB_EINVAL, "dist is NULL");
  return dist_barrier(d, current_stream());
}

// test / bench aid: all-to-all of equal blocks (block_bytes per peer) on the library stream
int ffb_dist_alltoall(ffb_dist* d, const void* sendbuf, void* recvbuf, size_t block_bytes) {
  FFB_REQUIRE(d && sendbuf && recvbuf, FFB_EINVAL, "NULL argument");
  return dist_alltoall_bytes(d, sendbuf, recvbuf, block_bytes, block_bytes, current_stream());
}

}  // extern "C"
