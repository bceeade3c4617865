// Dispatch entry points exported by the fft_pow2_inst.cu objects (one per real type and size group).
// Each returns FFB_OK, a negative ffb_status, or 1 when the length is not in its group.
#pragma once
#include <cstddef>

namespace ffb {
// Largest power-of-two line length the register-resident kernels cover for a real type of `real_bytes` bytes.
inline int pow2_max_n(int real_bytes) { return real_bytes == 8 ? 8192 : 16384; }
inline int pow2_max_threads(int real_bytes) { return real_bytes == 8 ? 512 : 1024; }
inline int pow2_points_per_thread(int N) { return N < 16 ? N : 16; }
// Radix sequence of each instantiated plan (must match the switch in fft_pow2_inst.cu); returns the pass count.
inline int pow2_radices(int N, int* r) {
  switch (N) {
    case 2: r[0] = 2; return 1;
    case 4: r[0] = 4; return 1;
    case 8: r[0] = 8; return 1;
    case 16: r[0] = 16; return 1;
    case 32: r[0] = 16; r[1] = 2; return 2;
    case 64: r[0] = 16; r[1] = 4; return 2;
    case 128: r[0] = 16; r[1] = 8; return 2;
    case 256: r[0] = 16; r[1] = 16; return 2;
    case 512: r[0] = 16; r[1] = 16; r[2] = 2; return 3;
    case 1024: r[0] = 16; r[1] = 16; r[2] = 4; return 3;
    case 2048: r[0] = 16; r[1] = 16; r[2] = 8; return 3;
    case 4096: r[0] = 16; r[1] = 16; r[2] = 16; return 3;
    case 8192: r[0] = 16; r[1] = 16; r[2] = 8; r[3] = 4; return 4;
    case 16384: r[0] = 16; r[1] = 16; r[2] = 16; r[3] = 4; return 4;
  }
  return 0;
}
// Float64 row passes with 8 points per thread (64 registers, 32 warps per SM instead of 16): radix sequence / availability
inline int pow2_radices_r8(int N, int* r) {
  switch (N) {
    case 512: r[0] = 8; r[1] = 8; r[2] = 8; return 3;
    case 1024: r[0] = 8; r[1] = 8; r[2] = 8; r[3] = 2; return 4;
    case 2048: r[0] = 8; r[1] = 8; r[2] = 8; r[3] = 4; return 4;
    case 4096: r[0] = 8; r[1] = 8; r[2] = 8; r[3] = 8; return 4;
  }
  return 0;
}
}  // namespace ffb

// rows (C2C_ROWS / R2C_ROWS / C2R_ROWS) in Float64 with 8 points per thread; 1 = not instantiated
int pow2_launch_double_r8(int N, int mode, int dir, const void* params, int gx, int gy, int threads, size_t smem, void* stream);

#define FFB_POW2_DECL(tn, g) \
  int pow2_launch_##tn##_g##g(int N, int mode, int dir, const void* params, int gx, int gy, int threads, size_t smem, void* stream);
FFB_POW2_DECL(float, 0) FFB_POW2_DECL(float, 1) FFB_POW2_DECL(float, 2) FFB_POW2_DECL(float, 3)
FFB_POW2_DECL(double, 0) FFB_POW2_DECL(double, 1) FFB_POW2_DECL(double, 2) FFB_POW2_DECL(double, 3)
#undef FFB_POW2_DECL
