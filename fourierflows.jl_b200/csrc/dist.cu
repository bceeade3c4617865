// NCCL plumbing for the slab-decomposed transforms.  The reference is single-device (README.md:58,
// docs/src/gpu.md:57): there is no reference counterpart; this is new capability named by BASELINE.json north_star.
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>
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
  d->ev_next = (d->ev_next + 1) % 64;
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
  FFB_CUDA(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
  for (auto& e : d->ev) FFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  d->ncopy = 4;
  if (const char* e = getenv("FFB_COPY_STREAMS")) { const int v = atoi(e); if (v >= 1 && v <= 4) d->ncopy = v; }
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

// CUDA IPC plumbing for the fused pass + collective path: a rank exports the handle of one of its buffers (64 bytes), the
// launcher moves it to the other ranks, which map it (peer access is enabled lazily).
int ffb_dist_ipc_export(void* dev_ptr, void* host_handle64) {
  FFB_REQUIRE(dev_ptr && host_handle64, FFB_EINVAL, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  cudaIpcMemHandle_t h;
  FFB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(host_handle64, &h, 64);
  return FFB_OK;
}

int ffb_dist_ipc_open(const void* host_handle64, void** dev_ptr) {
  FFB_REQUIRE(dev_ptr && host_handle64, FFB_EINVAL, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle64, 64);
  FFB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FFB_OK;
}

int ffb_dist_ipc_close(void* dev_ptr) {
  if (dev_ptr) FFB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
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
