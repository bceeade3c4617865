// Dispatch entry points exported by the fft_l2four_inst.cu objects (one per real type): the four-step sub-pass tiles of
// fft_fs.cuh as stand-alone kernels (fs_call_*) and as the persistent fused kernel with the intermediate in L2 (l2four_call_*).
// Return FFB_OK, a negative ffb_status, or 1 when the size is not instantiated.
//   l2four op = 0: launch with `grid` CTAs; op = 1: return the number of CTAs of the kernel that fit one SM (<= 0: error).
//   hook: 0 = none, 1 = prologue on sub-pass A (inverse transforms), 2 = epilogue on sub-pass B (forward transforms).
#pragma once
#include <cstddef>

namespace ffb {
constexpr int kFsThreads = 256;
inline int fs_points_per_thread(int real_bytes) { return real_bytes == 8 ? 8 : 16; }
inline bool fs_has_n(int N) { return N == 32 || N == 64 || N == 128 || N == 256; }
inline bool l2four_has(int N1, int N2) { return fs_has_n(N1) && fs_has_n(N2) && (N2 == N1 || N2 == 2 * N1); }
// radix sequence of the sub-transform plans of fft_fs.cuh (twiddle tables must match): returns the pass count
inline int fs_radices(int N, int real_bytes, int* r) {
  if (real_bytes == 8) {
    switch (N) {
      case 32: r[0] = 8; r[1] = 4; return 2;
      case 64: r[0] = 8; r[1] = 8; return 2;
      case 128: r[0] = 8; r[1] = 8; r[2] = 2; return 3;
      case 256: r[0] = 8; r[1] = 8; r[2] = 4; return 3;
    }
    return 0;
  }
  switch (N) {
    case 32: r[0] = 16; r[1] = 2; return 2;
    case 64: r[0] = 16; r[1] = 4; return 2;
    case 128: r[0] = 16; r[1] = 8; return 2;
    case 256: r[0] = 16; r[1] = 16; return 2;
  }
  return 0;
}
}  // namespace ffb

int l2four_call_float(int op, int N1, int N2, int dir, int hook, const void* params, int grid, size_t smem, void* stream);
int l2four_call_double(int op, int N1, int N2, int dir, int hook, const void* params, int grid, size_t smem, void* stream);
// stand-alone sub-pass: is_a != 0: sub-pass A of length N (inter-pass twiddle, optional prologue), else sub-pass B (optional epilogue)
int fs_call_float(int is_a, int N, int dir, int hook, const void* launch, int gx, int gy, size_t smem, void* stream);
int fs_call_double(int is_a, int N, int dir, int hook, const void* launch, int gx, int gy, size_t smem, void* stream);
// sub-pass A of up to kFsMaxVariants inverse transforms of one input (FsMultiLaunch): each variant has its own prologue and output
int fs_multi_call_float(int N, int dir, const void* launch, int gx, int gy, size_t smem, void* stream);
int fs_multi_call_double(int N, int dir, const void* launch, int gx, int gy, size_t smem, void* stream);
