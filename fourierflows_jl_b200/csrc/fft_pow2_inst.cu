// Instantiations of the power-of-two FFT kernels.  Compiled several times (see Makefile) with
//   -DFFB_REAL=float|double  -DFFB_GROUP=0..3
// so that the heavy unrolled kernels build in parallel.  Each object exports one dispatch function.
#include "fft_pow2.cuh"
#include "fft_pow2_dispatch.h"

#ifndef FFB_REAL
#define FFB_REAL double
#endif
#ifndef FFB_GROUP
#define FFB_GROUP 0
#endif

namespace ffb {

using real_t = FFB_REAL;
// Float64: 16 complex points = 64 data registers -> 512 threads x 128 registers.
// Float32: 32 data registers -> 1024 threads x 64 registers, so a CTA can own twice as many points.
constexpr int kMaxT = sizeof(real_t) == 8 ? 512 : 1024;

#if FFB_GROUP == 4
constexpr int kMinBlocks = 2;   // R = 8 row plans: 512 threads x 64 registers, two CTAs (32 warps) per SM
#else
constexpr int kMinBlocks = 1;
#endif

template <int MODE, int DIR, int R, int... Rs>
static int launch_one(const Pow2Params<real_t>& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  auto kern = fft_pow2_kernel<real_t, DIR, MODE, kMaxT, kMinBlocks, R, Rs...>;
  static size_t configured = 0;  // per instantiation
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  kern<<<grid, threads, smem, st>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "fft_pow2 launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

template <int R, int... Rs>
static int launch_n(int mode, int dir, const Pow2Params<real_t>& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  switch (mode) {
    case C2C_ROWS: return dir < 0 ? launch_one<C2C_ROWS, -1, R, Rs...>(p, grid, threads, smem, st) : launch_one<C2C_ROWS, 1, R, Rs...>(p, grid, threads, smem, st);
    case C2C_COLS: return dir < 0 ? launch_one<C2C_COLS, -1, R, Rs...>(p, grid, threads, smem, st) : launch_one<C2C_COLS, 1, R, Rs...>(p, grid, threads, smem, st);
    case C2C_COLS_LEAN: return dir < 0 ? launch_one<C2C_COLS_LEAN, -1, R, Rs...>(p, grid, threads, smem, st) : launch_one<C2C_COLS_LEAN, 1, R, Rs...>(p, grid, threads, smem, st);
    case C2C_COLS_TW:
      if constexpr (radix_product<Rs...>::value <= 256)
        return dir < 0 ? launch_one<C2C_COLS_TW, -1, R, Rs...>(p, grid, threads, smem, st) : launch_one<C2C_COLS_TW, 1, R, Rs...>(p, grid, threads, smem, st);
      else return set_error(FFB_EUNSUPPORTED, "four-step sub-transforms are instantiated for N <= 256");
    case R2C_ROWS: return launch_one<R2C_ROWS, -1, R, Rs...>(p, grid, threads, smem, st);
    case C2R_ROWS: return launch_one<C2R_ROWS, 1, R, Rs...>(p, grid, threads, smem, st);
  }
  return set_error(FFB_EINVAL, "bad fft mode %d", mode);
}

// contiguous-line modes only (the R = 8 row plans of group 4)
template <int R, int... Rs>
static int launch_rows(int mode, int dir, const Pow2Params<real_t>& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  switch (mode) {
    case C2C_ROWS: return dir < 0 ? launch_one<C2C_ROWS, -1, R, Rs...>(p, grid, threads, smem, st) : launch_one<C2C_ROWS, 1, R, Rs...>(p, grid, threads, smem, st);
    case R2C_ROWS: return launch_one<R2C_ROWS, -1, R, Rs...>(p, grid, threads, smem, st);
    case C2R_ROWS: return launch_one<C2R_ROWS, 1, R, Rs...>(p, grid, threads, smem, st);
  }
  return 1;
}

#define FFB_CAT2(a, b) a##b
#define FFB_CAT(a, b) FFB_CAT2(a, b)
#if FFB_GROUP == 4
#define FFB_FN(tn) FFB_CAT(FFB_CAT(pow2_launch_, tn), _r8)
#elif FFB_GROUP == 0
#define FFB_FN(tn) FFB_CAT(FFB_CAT(pow2_launch_, tn), _g0)
#elif FFB_GROUP == 1
#define FFB_FN(tn) FFB_CAT(FFB_CAT(pow2_launch_, tn), _g1)
#elif FFB_GROUP == 2
#define FFB_FN(tn) FFB_CAT(FFB_CAT(pow2_launch_, tn), _g2)
#else
#define FFB_FN(tn) FFB_CAT(FFB_CAT(pow2_launch_, tn), _g3)
#endif

// returns 1 if N is not handled by this group
static int dispatch(int N, int mode, int dir, const Pow2Params<real_t>& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  switch (N) {
#if FFB_GROUP == 4
    case 512: return launch_rows<8, 8, 8, 8>(mode, dir, p, grid, threads, smem, st);
    case 1024: return launch_rows<8, 8, 8, 8, 2>(mode, dir, p, grid, threads, smem, st);
    case 2048: return launch_rows<8, 8, 8, 8, 4>(mode, dir, p, grid, threads, smem, st);
    case 4096: return launch_rows<8, 8, 8, 8, 8>(mode, dir, p, grid, threads, smem, st);
#elif FFB_GROUP == 0
    case 2: return launch_n<2, 2>(mode, dir, p, grid, threads, smem, st);
    case 4: return launch_n<4, 4>(mode, dir, p, grid, threads, smem, st);
    case 8: return launch_n<8, 8>(mode, dir, p, grid, threads, smem, st);
    case 16: return launch_n<16, 16>(mode, dir, p, grid, threads, smem, st);
    case 32: return launch_n<16, 16, 2>(mode, dir, p, grid, threads, smem, st);
    case 64: return launch_n<16, 16, 4>(mode, dir, p, grid, threads, smem, st);
#elif FFB_GROUP == 1
    case 128: return launch_n<16, 16, 8>(mode, dir, p, grid, threads, smem, st);
    case 256: return launch_n<16, 16, 16>(mode, dir, p, grid, threads, smem, st);
    case 512: return launch_n<16, 16, 16, 2>(mode, dir, p, grid, threads, smem, st);
#elif FFB_GROUP == 2
    case 1024: return launch_n<16, 16, 16, 4>(mode, dir, p, grid, threads, smem, st);
    case 2048: return launch_n<16, 16, 16, 8>(mode, dir, p, grid, threads, smem, st);
#else
    case 4096: return launch_n<16, 16, 16, 16>(mode, dir, p, grid, threads, smem, st);
    case 8192: return launch_n<16, 16, 16, 8, 4>(mode, dir, p, grid, threads, smem, st);
    case 16384:
      if constexpr (sizeof(real_t) == 4) return launch_n<16, 16, 16, 16, 4>(mode, dir, p, grid, threads, smem, st);
      else return 1;
#endif
    default: return 1;
  }
}

}  // namespace ffb

int FFB_FN(FFB_REAL)(int N, int mode, int dir, const void* params, int gx, int gy, int threads, size_t smem, void* stream) {
  return ffb::dispatch(N, mode, dir, *reinterpret_cast<const ffb::Pow2Params<ffb::real_t>*>(params), dim3(gx, gy, 1), threads, smem,
                       reinterpret_cast<cudaStream_t>(stream));
}
