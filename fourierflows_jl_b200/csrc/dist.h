// Multi-GPU plumbing shared by dist.cu and fft_plan.cu: one process per GPU, NCCL communicator handed in through
// the C ABI (ffb_dist_*).  NCCL is loaded lazily (dlopen) so that the library also loads on machines without it.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

struct ffb_dist {
  void* comm;              // ncclComm_t
  int rank, nranks;
  cudaStream_t comm_stream;
  static constexpr int kEvents = 256;
  cudaEvent_t ev[kEvents]; // ring of events for compute <-> comm ordering (one transform uses at most ~10 per chunk, <= 8 chunks)
  int ev_next;
  float* barrier_buf;      // device scratch of the barrier all-reduce
  cudaStream_t copy_streams[8];   // copy-engine exchange: peer copies are spread over these streams
  int ncopy;
};

namespace ffb {
// grouped exchange: for every peer s, send `count` bytes at sendbuf + s*stride_bytes and receive into recvbuf + s*stride_bytes
// (the own block is a device-to-device copy).  Enqueued on `st`.
int dist_alltoall_bytes(ffb_dist* d, const void* sendbuf, void* recvbuf, size_t count, size_t stride_bytes, cudaStream_t st);
cudaEvent_t dist_next_event(ffb_dist* d);
// stream-ordered barrier across all ranks (1-element NCCL all-reduce): every rank's earlier work on `st` has completed
// (including its stores into peer memory) before any rank's later work starts
int dist_barrier(ffb_dist* d, cudaStream_t st);
// copy-engine exchange (push): once the work already enqueued on `after` has finished (nullptr: no dependency), block s of
// sendbuf (count bytes at sendbuf + s*stride_bytes) is copied into peer s's buffer at peer_bufs[s] + dst_off -- cudaMemcpyAsync
// over NVLink, no SM involved.  d->comm_stream is made to wait for all of the copies (follow with dist_barrier on it).
int dist_push_blocks(ffb_dist* d, const void* sendbuf, void* const* peer_bufs, size_t dst_off, size_t count, size_t stride_bytes, cudaStream_t after);
}  // namespace ffb
