// Instantiations of the four-step sub-pass kernels (fft_fs.cuh) and of the fused, L2-resident four-step kernel
// (fft_l2four.cuh).  Compiled once per real type: -DFFB_REAL=float|double.
#include "fft_l2four.cuh"
#include "fft_l2four_dispatch.h"

#ifndef FFB_REAL
#define FFB_REAL double
#endif

namespace ffb {

using real_t = FFB_REAL;
// Float64: R = 8 points per thread, <= 85 registers -> 3 CTAs of 256 threads (24 warps) per SM; Float32: R = 16, 64 registers -> 4 CTAs
constexpr int kMinB = sizeof(real_t) == 8 ? 3 : 4;

template <class K>
static int configure(K kern, size_t smem, size_t& configured) {
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  return FFB_OK;
}

template <int DIR, class PA, class PB, bool HA, bool HB>
static int l2_one(int op, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = fft_l2four_kernel<real_t, DIR, PA, PB, HA, HB, kFsThreads, kMinB>;
  static size_t configured = 0;
  int rc = configure(kern, smem, configured);
  if (rc) return rc;
  if (op == 1) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kFsThreads, smem);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
    return nb;
  }
  kern<<<grid, kFsThreads, smem, st>>>(p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "fft_l2four launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

template <int N1, int N2>
static int l2_pair(int op, int dir, int hook, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  using PA = typename fs_plan_for<real_t, N1>::type;
  using PB = typename fs_plan_for<real_t, N2>::type;
  if (hook == 0) return dir < 0 ? l2_one<-1, PA, PB, false, false>(op, p, grid, smem, st) : l2_one<1, PA, PB, false, false>(op, p, grid, smem, st);
  if (hook == 1 && dir > 0) return l2_one<1, PA, PB, true, false>(op, p, grid, smem, st);
  if (hook == 2 && dir < 0) return l2_one<-1, PA, PB, false, true>(op, p, grid, smem, st);
  return set_error(FFB_EINVAL, "fused four-step: prologues belong to inverse, epilogues to forward transforms");
}

static int l2_dispatch(int op, int N1, int N2, int dir, int hook, const L2FourParams<real_t>& p, int grid, size_t smem, cudaStream_t st) {
  if (N1 == 32 && N2 == 32) return l2_pair<32, 32>(op, dir, hook, p, grid, smem, st);
  if (N1 == 32 && N2 == 64) return l2_pair<32, 64>(op, dir, hook, p, grid, smem, st);
  if (N1 == 64 && N2 == 64) return l2_pair<64, 64>(op, dir, hook, p, grid, smem, st);
  if (N1 == 64 && N2 == 128) return l2_pair<64, 128>(op, dir, hook, p, grid, smem, st);
  if (N1 == 128 && N2 == 128) return l2_pair<128, 128>(op, dir, hook, p, grid, smem, st);
  if (N1 == 128 && N2 == 256) return l2_pair<128, 256>(op, dir, hook, p, grid, smem, st);
  if (N1 == 256 && N2 == 256) return l2_pair<256, 256>(op, dir, hook, p, grid, smem, st);
  return 1;
}

template <int DIR, bool IS_A, bool HOOK, int N>
static int fs_one(const FsLaunch<real_t>& q, dim3 grid, size_t smem, cudaStream_t st) {
  using PL = typename fs_plan_for<real_t, N>::type;
  // (measured: squeezing the hook-free Float64 sub-passes to 64 registers for 4 CTAs = 32 warps per SM changes nothing, 14.93 vs
  // 14.84 ms per C3 step -- they are no longer latency-bound)
  auto kern = fs_pass_kernel<real_t, DIR, IS_A, HOOK, PL, kFsThreads, kMinB>;
  static size_t configured = 0;
  int rc = configure(kern, smem, configured);
  if (rc) return rc;
  kern<<<grid, kFsThreads, smem, st>>>(q);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "four-step sub-pass launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

template <int N>
static int fs_n(int is_a, int dir, int hook, const FsLaunch<real_t>& q, dim3 grid, size_t smem, cudaStream_t st) {
  if (is_a) {
    if (hook) return dir > 0 ? fs_one<1, true, true, N>(q, grid, smem, st) : set_error(FFB_EINVAL, "prologue on a forward transform");
    return dir < 0 ? fs_one<-1, true, false, N>(q, grid, smem, st) : fs_one<1, true, false, N>(q, grid, smem, st);
  }
  if (hook) return dir < 0 ? fs_one<-1, false, true, N>(q, grid, smem, st) : set_error(FFB_EINVAL, "epilogue on an inverse transform");
  return dir < 0 ? fs_one<-1, false, false, N>(q, grid, smem, st) : fs_one<1, false, false, N>(q, grid, smem, st);
}

static int fs_dispatch(int is_a, int N, int dir, int hook, const FsLaunch<real_t>& q, dim3 grid, size_t smem, cudaStream_t st) {
  switch (N) {
    case 32: return fs_n<32>(is_a, dir, hook, q, grid, smem, st);
    case 64: return fs_n<64>(is_a, dir, hook, q, grid, smem, st);
    case 128: return fs_n<128>(is_a, dir, hook, q, grid, smem, st);
    case 256: return fs_n<256>(is_a, dir, hook, q, grid, smem, st);
  }
  return 1;
}

template <int N>
static int fs_multi_one(int dir, const FsMultiLaunch<real_t>& mq, dim3 grid, size_t smem, cudaStream_t st) {
  using PL = typename fs_plan_for<real_t, N>::type;
  if (dir < 0) return set_error(FFB_EINVAL, "multi-variant sub-pass A: inverse transforms only");
  auto kern = fs_pass_multi_kernel<real_t, 1, PL, kFsThreads, kMinB>;
  static size_t configured = 0;
  int rc = configure(kern, smem, configured);
  if (rc) return rc;
  kern<<<grid, kFsThreads, smem, st>>>(mq);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "multi-variant four-step sub-pass launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

static int fs_multi_dispatch(int N, int dir, const FsMultiLaunch<real_t>& mq, dim3 grid, size_t smem, cudaStream_t st) {
  switch (N) {
    case 32: return fs_multi_one<32>(dir, mq, grid, smem, st);
    case 64: return fs_multi_one<64>(dir, mq, grid, smem, st);
    case 128: return fs_multi_one<128>(dir, mq, grid, smem, st);
    case 256: return fs_multi_one<256>(dir, mq, grid, smem, st);
  }
  return 1;
}

}  // namespace ffb

#define FFB_CAT2(a, b) a##b
#define FFB_CAT(a, b) FFB_CAT2(a, b)
int FFB_CAT(l2four_call_, FFB_REAL)(int op, int N1, int N2, int dir, int hook, const void* params, int grid, size_t smem, void* stream) {
  return ffb::l2_dispatch(op, N1, N2, dir, hook, *reinterpret_cast<const ffb::L2FourParams<ffb::real_t>*>(params), grid, smem,
                          reinterpret_cast<cudaStream_t>(stream));
}
int FFB_CAT(fs_call_, FFB_REAL)(int is_a, int N, int dir, int hook, const void* launch, int gx, int gy, size_t smem, void* stream) {
  return ffb::fs_dispatch(is_a, N, dir, hook, *reinterpret_cast<const ffb::FsLaunch<ffb::real_t>*>(launch), dim3(gx, gy, 1), smem,
                          reinterpret_cast<cudaStream_t>(stream));
}
int FFB_CAT(fs_multi_call_, FFB_REAL)(int N, int dir, const void* launch, int gx, int gy, size_t smem, void* stream) {
  return ffb::fs_multi_dispatch(N, dir, *reinterpret_cast<const ffb::FsMultiLaunch<ffb::real_t>*>(launch), dim3(gx, gy, 1), smem,
                                reinterpret_cast<cudaStream_t>(stream));
}
