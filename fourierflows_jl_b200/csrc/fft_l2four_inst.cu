// Instantiations of the fused (L2-resident) four-step kernel.  Compiled once per real type: -DFFB_REAL=float|double.
#include "fft_l2four.cuh"
#include "fft_l2four_dispatch.h"

#ifndef FFB_REAL
#define FFB_REAL double
#endif

namespace ffb {

using real_t = FFB_REAL;
// Float64: 64 data registers -> 128 registers x 512 threads per SM (4 CTAs of 128); Float32: 64 registers x 1024 threads (8 CTAs)
constexpr int kMinB = sizeof(real_t) == 8 ? 4 : 8;

template <int DIR, class PA, class PB>
static int call_one(int op, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = fft_l2four_kernel<real_t, DIR, PA, PB, kL2FourThreads, kMinB>;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  if (op == 1) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kL2FourThreads, smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
    return nb;
  }
  kern<<<grid, kL2FourThreads, smem, st>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "fft_l2four launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

using P32 = RadixPlan<16, 16, 2>;
using P64 = RadixPlan<16, 16, 4>;
using P128 = RadixPlan<16, 16, 8>;
using P256 = RadixPlan<16, 16, 16>;

template <class PA, class PB>
static int call_dir(int op, int dir, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  return dir < 0 ? call_one<-1, PA, PB>(op, p, grid, smem, st) : call_one<1, PA, PB>(op, p, grid, smem, st);
}

static int dispatch(int op, int N1, int N2, int dir, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  if (N1 == 32 && N2 == 32) return call_dir<P32, P32>(op, dir, p, grid, smem, st);
  if (N1 == 32 && N2 == 64) return call_dir<P32, P64>(op, dir, p, grid, smem, st);
  if (N1 == 64 && N2 == 64) return call_dir<P64, P64>(op, dir, p, grid, smem, st);
  if (N1 == 64 && N2 == 128) return call_dir<P64, P128>(op, dir, p, grid, smem, st);
  if (N1 == 128 && N2 == 128) return call_dir<P128, P128>(op, dir, p, grid, smem, st);
  if (N1 == 128 && N2 == 256) return call_dir<P128, P256>(op, dir, p, grid, smem, st);
  if (N1 == 256 && N2 == 256) return call_dir<P256, P256>(op, dir, p, grid, smem, st);
  return 1;
}

}  // namespace ffb

#define FFB_CAT2(a, b) a##b
#define FFB_CAT(a, b) FFB_CAT2(a, b)
int FFB_CAT(l2four_call_, FFB_REAL)(int op, int N1, int N2, int dir, const void* params, int grid, size_t smem, void* stream) {
  return ffb::dispatch(op, N1, N2, dir, *reinterpret_cast<const ffb::L2FourParams<ffb::real_t>*>(params), grid, smem,
                       reinterpret_cast<cudaStream_t>(stream));
}
