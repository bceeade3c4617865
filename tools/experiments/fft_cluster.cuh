// Cluster (distributed-shared-memory) four-step FFT for long strided lines -- sm_100a thread-block clusters.
//
// A strided line of N = N1*N2 points cannot be owned by one CTA with wide rows (8192 Float64 points = 128 KB already
// fill the register file: one 16-byte column per CTA runs at 25 % of HBM peak).  Here a CLUSTER of C CTAs
// owns a tile of W adjacent columns x N rows and performs the four-step algorithm
//     n = N2*n1 + n2,  k = k1 + N1*k2:
//     A) CTA c: for its n2 in [c*N2/C, (c+1)*N2/C): length-N1 transform over n1, times exp(-/+2*pi*i*n2*k1/N)
//     T) transpose through DSMEM: element (k1, n2) is stored into the receive buffer of CTA k1/(N1/C)
//     B) CTA d: for its k1: length-N2 transform over n2, written to row k1 + N1*k2
// so the tile makes ONE HBM round trip with W*sizeof(complex)-wide rows, in place (every row of the tile is read before the
// cluster barrier, written after it).  Index maps validated in tools/cluster_model.py.
#pragma once
#include <cooperative_groups.h>
#include "fft_pow2.cuh"

namespace ffb {

namespace cg = cooperative_groups;

template <int... Rs> struct Radix {
  static constexpr int N = radix_product<Rs...>::value;
};

template <typename T>
struct ClusterParams {
  const cx<T>* in;
  cx<T>* out;
  long long es;         // element stride along the transform dimension (= number of columns of the array)
  long long os;         // outer stride (next slice)
  long long ncols;      // columns (lines) per slice
  T scale;
  const cx<T>* tw1;     // base twiddles of the N1 plan
  const cx<T>* tw2;     // base twiddles of the N2 plan
  const cx<T>* twN;     // exp(-2*pi*i*q/N), q < N
  typename Pow2Params<T>::Fuse pro, epi;   // fused load prologue / store epilogue (see fft_pow2.cuh)
};

template <typename T, int DIR, int N, int... Rs>
FFB_D void run_line(Radix<Rs...>, cx<T> (&v)[16], int t, int lw, int LW, typename xword<T>::type* xb, const cx<T>* tw) {
  run_passes<T, DIR, true, 16, N, 1, 0, Rs...>(v, t, lw, LW, xb, tw);
}

template <int C, int W, int N1, int N2, typename T> struct ClusterGeom {
  static constexpr int Tn1 = N1 / 16, Tn2 = N2 / 16;
  static constexpr int NLA = (N2 / C) * W, NLB = (N1 / C) * W;   // sub-lines per CTA in steps A and B
  static constexpr int NT = NLA * Tn1;                           // threads per CTA
  static constexpr int PAD = (W * (int)sizeof(cx<T>) >= 128) ? 0 : 64 / (int)sizeof(cx<T>);
  static constexpr int RS = (N1 / C) * W + PAD;                  // receive-buffer row stride (elements)
  static constexpr size_t xb_words = (size_t)((xpad_len(N1) * NLA > xpad_len(N2) * NLB) ? xpad_len(N1) * NLA : xpad_len(N2) * NLB);
  static constexpr size_t smem_bytes = xb_words * 8 + (size_t)N2 * RS * sizeof(cx<T>);
  static_assert(NT == NLB * Tn2, "steps A and B must use the same number of threads");
  static_assert(N1 % (16 * 1) == 0 && N2 % 16 == 0 && N1 % C == 0 && N2 % C == 0, "bad cluster geometry");
};

template <typename T, int DIR, int C, int W, typename RA, typename RB, int MINB>
__global__ void __launch_bounds__((ClusterGeom<C, W, RA::N, RB::N, T>::NT), MINB) fft_cluster_kernel(const ClusterParams<T> p) {
  constexpr int N1 = RA::N, N2 = RB::N, N = N1 * N2;
  using G = ClusterGeom<C, W, N1, N2, T>;
  constexpr int Tn1 = G::Tn1, Tn2 = G::Tn2, NLA = G::NLA, NLB = G::NLB, RS = G::RS;
  using XW = typename xword<T>::type;
  extern __shared__ __align__(16) unsigned char ffb_smem[];
  XW* xb = reinterpret_cast<XW*>(ffb_smem);
  cx<T>* rbuf = reinterpret_cast<cx<T>*>(ffb_smem + G::xb_words * 8);

  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const long long tile = blockIdx.x / C;
  const long long col0 = tile * W;
  const long long slice = blockIdx.y;
  const cx<T>* in = p.in + slice * p.os;
  cx<T>* out = p.out + slice * p.os;

  // ---------------- step A: length-N1 transforms over n1 for this CTA's n2 ----------------
  cx<T> v[16];
  {
    const int lw = tid % NLA, t = tid / NLA;
    const int w = lw % W, n2 = c * (N2 / C) + lw / W;
    const bool active = col0 + w < p.ncols;
#pragma unroll
    for (int m = 0; m < 16; ++m)
      v[m] = active ? ldc(in + (long long)(N2 * (t + m * Tn1) + n2) * p.es + col0 + w) : mk<T>(0, 0);
    if (p.pro.on && active) {
      const int i0 = (int)((col0 + w) % p.pro.n0);
      const long long io = p.pro.other_from_col == 1 ? (col0 + w) / p.pro.n0 : slice;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int it = N2 * (t + m * Tn1) + n2;
        const long long off = slice * p.os + (long long)it * p.es + col0 + w;
        v[m] = fuse_factor<T>(p.pro.cr, p.pro.ci, p.pro.k0, p.pro.kt, p.pro.ko, p.pro.w, i0, it, io, off) * v[m];
      }
    }
    cluster.sync();   // every CTA of the cluster is resident before any remote shared-memory access
    run_line<T, DIR, N1>(RA{}, v, t, lw, NLA, xb, p.tw1);
    // inter-step twiddle exp(-/+2*pi*i*n2*k1/N) and scatter into the owner of k1
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int k1 = t + m * Tn1;
      if (k1 != 0) v[m] = v[m] * load_tw<T, DIR>(p.twN + ((n2 * k1) & (N - 1)));
      const int d = k1 / (N1 / C), k1l = k1 % (N1 / C);
      cx<T>* remote = cluster.map_shared_rank(rbuf, d);
      using V = typename vec2<T>::type;
      V q; q.x = v[m].x; q.y = v[m].y;
      *reinterpret_cast<V*>(remote + ((lw / W + c * (N2 / C)) * RS + k1l * W + w)) = q;
    }
  }
  cluster.sync();     // all contributions have landed in every receive buffer
  // ---------------- step B: length-N2 transforms over n2 for this CTA's k1 ----------------
  {
    const int lw = tid % NLB, t2 = tid / NLB;
    const int w = lw % W, k1l = lw / W;
    const bool active = col0 + w < p.ncols;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      using V = typename vec2<T>::type;
      const V q = *reinterpret_cast<const V*>(rbuf + ((t2 + m * Tn2) * RS + k1l * W + w));
      v[m] = mk<T>(q.x, q.y);
    }
    run_line<T, DIR, N2>(RB{}, v, t2, lw, NLB, xb, p.tw2);
    if (active) {
      const int k1 = c * (N1 / C) + k1l;
      const T sc = p.scale;
      if (p.epi.on) {
        const int i0 = (int)((col0 + w) % p.epi.n0);
        const long long io = p.epi.other_from_col == 1 ? (col0 + w) / p.epi.n0 : slice;
        const bool dead0 = p.epi.dealias && ((p.epi.lo0 > 0 && i0 >= p.epi.lo0 - 1 && i0 < p.epi.hi0) || (p.epi.loo > 0 && io >= p.epi.loo - 1 && io < p.epi.hio));
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int it = k1 + N1 * (t2 + m * Tn2);
          const long long o = (long long)it * p.es + col0 + w;
          cx<T> r;
          if (dead0 || (p.epi.dealias && p.epi.lot > 0 && it >= p.epi.lot - 1 && it < p.epi.hit)) {
            r = mk<T>(0, 0);
          } else {
            r = fuse_factor<T>(p.epi.cr, p.epi.ci, p.epi.k0, p.epi.kt, p.epi.ko, p.epi.w, i0, it, io, slice * p.os + o) * (sc * v[m]);
            if (p.epi.acc) r = r + fuse_factor<T>(p.epi.ar, p.epi.ai, p.epi.a0, p.epi.at, p.epi.ao, (const T*)nullptr, i0, it, io, 0) * ldc(p.epi.acc + slice * p.os + o);
          }
          stc(out + o, r);
        }
      } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const long long o = (long long)(k1 + N1 * (t2 + m * Tn2)) * p.es + col0 + w;
          stc(out + o, sc != T(1) ? sc * v[m] : v[m]);
        }
      }
    }
  }
}

}  // namespace ffb
