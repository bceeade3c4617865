// Dispatch entry points exported by the fft_l2four_inst.cu objects (one per real type).
// launch: FFB_OK, a negative ffb_status, or 1 when the (N1, N2) pair is not instantiated.  op = 0: launch with `grid` CTAs;
// op = 1: return the number of CTAs of this kernel that fit one SM with `smem` bytes of dynamic shared memory (<= 0: error).
#pragma once
#include <cstddef>

namespace ffb {
constexpr int kL2FourThreads = 128;
inline bool l2four_has(int N1, int N2) {
  return (N1 == 32 && (N2 == 32 || N2 == 64)) || (N1 == 64 && (N2 == 64 || N2 == 128)) || (N1 == 128 && (N2 == 128 || N2 == 256)) ||
         (N1 == 256 && N2 == 256);
}
}  // namespace ffb

int l2four_call_float(int op, int N1, int N2, int dir, const void* params, int grid, size_t smem, void* stream);
int l2four_call_double(int op, int N1, int N2, int dir, const void* params, int grid, size_t smem, void* stream);
