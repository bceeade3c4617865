// Instantiations and launch of the streaming strided-pass kernels (fft_stream.cuh).  Compiled with -DFFB_REAL=float|double.
#include "fft_stream.cuh"
#include "fft_stream_dispatch.h"

#ifndef FFB_REAL
#define FFB_REAL double
#endif

namespace ffb {

using real_t = FFB_REAL;
constexpr int kMaxT = sizeof(real_t) == 8 ? 512 : 1024;

template <int DIR, bool SPLIT, int... Rs>
static int launch_stream(const StreamParams<real_t>& p, int threads, size_t smem, cudaStream_t st) {
  auto kern = fft_cols_stream_kernel<real_t, DIR, SPLIT, kMaxT, Rs...>;
  static size_t configured = 0;
  static int occ_threads = 0, occ = 0;
  static size_t occ_smem = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "stream kernel smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  if (occ_threads != threads || occ_smem != smem) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
    if (e != cudaSuccess || nb < 1) return set_error(FFB_ECUDA, "stream kernel does not fit (threads %d, smem %zu)", threads, smem);
    occ = nb; occ_threads = threads; occ_smem = smem;
  }
  long long grid = (long long)occ * num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  kern<<<(unsigned)grid, threads, smem, st>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "stream FFT launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

template <int... Rs>
static int launch_n(int dir, int split, const StreamParams<real_t>& p, int threads, size_t smem, cudaStream_t st) {
  if constexpr (sizeof(real_t) == 4) {
    if (!split) return dir < 0 ? launch_stream<-1, false, Rs...>(p, threads, smem, st) : launch_stream<1, false, Rs...>(p, threads, smem, st);
  }
  return dir < 0 ? launch_stream<-1, true, Rs...>(p, threads, smem, st) : launch_stream<1, true, Rs...>(p, threads, smem, st);
}

}  // namespace ffb

#define FFB_CAT2(a, b) a##b
#define FFB_CAT(a, b) FFB_CAT2(a, b)

int FFB_CAT(stream_launch_, FFB_REAL)(int N, int dir, int split, const void* params, int threads, size_t smem, void* stream) {
  using namespace ffb;
  const auto& p = *reinterpret_cast<const StreamParams<real_t>*>(params);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (N) {   // radix sequences as in fft_pow2_inst.cu (the twiddle tables are shared)
    case 256: return launch_n<16, 16>(dir, split, p, threads, smem, st);
    case 512: return launch_n<16, 16, 2>(dir, split, p, threads, smem, st);
    case 1024: return launch_n<16, 16, 4>(dir, split, p, threads, smem, st);
    case 2048: return launch_n<16, 16, 8>(dir, split, p, threads, smem, st);
  }
  return 1;
}
