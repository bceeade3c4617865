// Instantiations and launch of the cluster (DSMEM) four-step kernels.  Compiled with -DFFB_REAL=float|double.
#include "fft_cluster.cuh"
#include "fft_cluster_dispatch.h"

#ifndef FFB_REAL
#define FFB_REAL double
#endif

namespace ffb {

using real_t = FFB_REAL;
constexpr int kC = 8;                                   // portable cluster size
constexpr int kW = sizeof(real_t) == 8 ? 4 : 8;        // 64-byte wide rows

template <int DIR, typename RA, typename RB>
static int launch_cluster(const ClusterParams<real_t>& p, long long ntiles, int nslices, cudaStream_t st) {
  using G = ClusterGeom<kC, kW, RA::N, RB::N, real_t>;
  constexpr int kRegs = sizeof(real_t) == 8 ? 128 : 64;
  constexpr int kMinB = (65536 / (G::NT * kRegs)) > 8 ? 8 : ((65536 / (G::NT * kRegs)) < 1 ? 1 : (65536 / (G::NT * kRegs)));
  auto kern = fft_cluster_kernel<real_t, DIR, kC, kW, RA, RB, kMinB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes);
    if (e != cudaSuccess) return set_error(FFB_ECUDA, "cluster kernel smem attribute (%zu bytes): %s", (size_t)G::smem_bytes, cudaGetErrorString(e));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ntiles * kC), (unsigned)nslices, 1);
  cfg.blockDim = dim3(G::NT, 1, 1);
  cfg.dynamicSmemBytes = G::smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  count_launch();
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "cluster FFT launch failed: %s", cudaGetErrorString(e));
  return FFB_OK;
}

template <typename RA, typename RB>
static int launch_dir(int dir, const ClusterParams<real_t>& p, long long ntiles, int nslices, cudaStream_t st) {
  return dir < 0 ? launch_cluster<-1, RA, RB>(p, ntiles, nslices, st) : launch_cluster<1, RA, RB>(p, ntiles, nslices, st);
}

}  // namespace ffb

#define FFB_CAT2(a, b) a##b
#define FFB_CAT(a, b) FFB_CAT2(a, b)

int FFB_CAT(cluster_launch_, FFB_REAL)(int N, int dir, const void* params, long long ntiles, int nslices, void* stream) {
  using namespace ffb;
  const auto& p = *reinterpret_cast<const ClusterParams<real_t>*>(params);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (N) {
    case 1024: return launch_dir<Radix<16, 2>, Radix<16, 2>>(dir, p, ntiles, nslices, st);
    case 2048: return launch_dir<Radix<16, 2>, Radix<16, 4>>(dir, p, ntiles, nslices, st);
    case 4096: return launch_dir<Radix<16, 4>, Radix<16, 4>>(dir, p, ntiles, nslices, st);
    case 8192: return launch_dir<Radix<16, 4>, Radix<16, 8>>(dir, p, ntiles, nslices, st);
  }
  return 1;
}
