// Multi-GPU plumbing shared by dist.cu and fft_plan.cu: one process per GPU, NCCL communicator handed in through
// the C ABI (ffb_dist_*).  NCCL is loaded lazily (dlopen) so that the library also loads on machines without it.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

struct ffb_dist {
  void* comm;              // ncclComm_t
  int rank, nranks;
  cudaStream_t comm_stream;
  cudaEvent_t ev[64];      // ring of events for compute <-> comm ordering
  int ev_next;
  float* barrier_buf;      // device scratch of the barrier all-reduce
};

namespace ffb {
// grouped exchange: for every peer s, send `count` bytes at sendbuf + s*stride_bytes and receive into recvbuf + s*stride_bytes
// (the own block is a device-to-device copy).  Enqueued on `st`.
int dist_alltoall_bytes(ffb_dist* d, const void* sendbuf, void* recvbuf, size_t count, size_t stride_bytes, cudaStream_t st);
cudaEvent_t dist_next_event(ffb_dist* d);
// stream-ordered barrier across all ranks (1-element NCCL all-reduce): every rank's earlier work on `st` has completed
// (including its stores into peer memory) before any rank's later work starts
int dist_barrier(ffb_dist* d, cudaStream_t st);
}  // namespace ffb
