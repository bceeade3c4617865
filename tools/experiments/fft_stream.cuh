// Streaming (software-pipelined, persistent) variant of the strided register-resident FFT pass -- sm_100a.
//
// Long strided lines leave room for ONE CTA per SM (a tile of N x W points fills the register file), so in the plain kernel
// (fft_pow2.cuh) the load, butterfly / exchange and store phases of a tile run one after the other and HBM idles while the
// SM computes (ncu: 35 % of HBM peak for Float32 N = 2048, issue slots half idle).  Here a CTA walks over many tiles and
// every thread copies ITS OWN 16 points of the next tile into a private shared-memory slot with cp.async (LDGSTS) while the
// current tile is transformed; no thread ever reads a slot it did not fill, so cp.async.wait_group is the only
// synchronisation the staging needs.  The address arithmetic of the plain kernel (segment masks, fusion hooks) is replaced
// by per-register offsets precomputed on the host (kernel parameters = constant-bank operands).
//
// Scope: plain C2C strided passes (in_ls = out_ls = 1), optional segmented strides whose segment length is a multiple of
// N/16, optional output scale.  Fused passes keep using fft_pow2_kernel.
#pragma once
#include "fft_pow2.cuh"

namespace ffb {

template <typename T>
struct StreamParams {
  const cx<T>* in;
  cx<T>* out;
  long long in_es, in_os, out_es, out_os;   // element / outer strides (complex elements); adjacent lines are contiguous
  long long in_off[16], out_off[16];        // offset of register m's point relative to the thread's first point
  long long nlines;                         // lines (columns) per outer index
  int W;                                    // columns per tile
  int gx;                                   // tiles per outer index = ceil(nlines / W)
  long long ntiles;                         // gx * nouter
  T scale;
  const cx<T>* tw;
  int keep_out;
};

template <int BYTES> FFB_D void cp_async(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
FFB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
FFB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// exchange word: SPLIT moves re and im in two phases of scalar words (half the shared memory), otherwise one complex word
template <typename T, bool SPLIT> struct sword { using type = T; static constexpr int phases = 2; };
template <> struct sword<float, false> { using type = float2; static constexpr int phases = 1; };

template <typename T, bool SPLIT, int R, int N, int Ns, int r>
FFB_D void exchange_s(cx<T> (&v)[R], int t, int w, int W, typename sword<T, SPLIT>::type* xb) {
  constexpr int Tn = N / R, nb = R / r;
  constexpr int lNs = ce_log2(Ns), lr = ce_log2(r);
  constexpr int PH = sword<T, SPLIT>::phases;
  auto addr = [&](int idx) { return xpad(idx) * W + w; };
  static_for<0, PH>([&](auto P) {
    [[maybe_unused]] constexpr int ph = decltype(P)::value;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int j = t + b * Tn;
      const int base = ((j >> lNs) << (lNs + lr)) + (j & (Ns - 1));
#pragma unroll
      for (int k = 0; k < r; ++k) {
        if constexpr (PH == 2) xb[addr(base + k * Ns)] = ph == 0 ? v[b + k * nb].x : v[b + k * nb].y;
        else xb[addr(base + k * Ns)] = make_float2(v[b + k * nb].x, v[b + k * nb].y);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if constexpr (PH == 2) { if (ph == 0) v[m].x = xb[addr(t + m * Tn)]; else v[m].y = xb[addr(t + m * Tn)]; }
      else { const float2 q = xb[addr(t + m * Tn)]; v[m].x = q.x; v[m].y = q.y; }
    }
  });
}

template <typename T, int DIR, bool SPLIT, int R, int N, int Ns, int TWOFF, int r, int... Rest>
FFB_D void run_passes_s(cx<T> (&v)[R], int t, int w, int W, typename sword<T, SPLIT>::type* xb, const cx<T>* tw) {
  constexpr int Tn = N / R, nb = R / r;
  cx<T> wb[nb];
  if constexpr (Ns > 1) {
#pragma unroll
    for (int b = 0; b < nb; ++b) wb[b] = load_tw<T, DIR>(tw + TWOFF + ((t + b * Tn) & (Ns - 1)));
  }
  static_for<0, nb>([&](auto B) {
    constexpr int b = decltype(B)::value;
    if constexpr (Ns > 1) apply_twiddle_powers<T, R, r, b>(v, wb[b]);
    bfly_at<DIR, R, r, b>(v);
  });
  if constexpr (sizeof...(Rest) > 0) {
    exchange_s<T, SPLIT, R, N, Ns, r>(v, t, w, W, xb);
    run_passes_s<T, DIR, SPLIT, R, N, Ns * r, TWOFF + (Ns > 1 ? Ns : 0), Rest...>(v, t, w, W, xb, tw);
  }
}

// shared memory: [exchange buffer | staging slots]
template <typename T, bool SPLIT> __host__ __device__ constexpr size_t stream_xb_bytes(int N, int W) {
  return (((size_t)xpad_len(N) * W * sizeof(typename sword<T, SPLIT>::type)) + 15) / 16 * 16;
}
template <typename T, bool SPLIT> constexpr size_t stream_smem_bytes(int N, int W) {
  return stream_xb_bytes<T, SPLIT>(N, W) + (size_t)N * W * sizeof(cx<T>);
}

template <typename T, int DIR, bool SPLIT, int MAXT, int... Rs>
__global__ void __launch_bounds__(MAXT, 1) fft_cols_stream_kernel(const StreamParams<T> p) {
  constexpr int N = radix_product<Rs...>::value;
  constexpr int R = 16, Tn = N / R;
  static_assert(N % R == 0 && N >= 256, "streaming kernel is for long lines");
  using XW = typename sword<T, SPLIT>::type;
  using V = typename vec2<T>::type;
  extern __shared__ __align__(16) unsigned char ffb_smem[];
  XW* xb = reinterpret_cast<XW*>(ffb_smem);
  const int W = p.W;
  const int NT = Tn * W;
  V* stage = reinterpret_cast<V*>(ffb_smem + stream_xb_bytes<T, SPLIT>(N, W)) + threadIdx.x;   // this thread's slots: stage[m*NT]
  const int tid = threadIdx.x;
  const int w = tid % W, t = tid / W;
  const long long tin = (long long)t * p.in_es + w, tout = (long long)t * p.out_es + w;

  auto tile_active = [&](long long tile) { return (tile % p.gx) * W + w < p.nlines; };
  auto prefetch = [&](long long tile) {
    if (tile < p.ntiles && tile_active(tile)) {
      const cx<T>* src = p.in + (tile / p.gx) * p.in_os + (tile % p.gx) * W + tin;
#pragma unroll
      for (int m = 0; m < R; ++m) cp_async<(int)sizeof(cx<T>)>(stage + m * NT, src + p.in_off[m]);
    }
    cp_async_commit();
  };

  long long tile = blockIdx.x;
  prefetch(tile);
  for (; tile < p.ntiles; tile += gridDim.x) {
    const bool active = tile_active(tile);
    cx<T> v[R];
    cp_async_wait_all();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (active) { const V q = stage[m * NT]; v[m] = mk<T>(q.x, q.y); }
      else v[m] = mk<T>(0, 0);
    }
    prefetch(tile + gridDim.x);   // lands in the slots just read while this tile is transformed
    run_passes_s<T, DIR, SPLIT, R, N, 1, 0, Rs...>(v, t, w, W, xb, p.tw);
    if (active) {
      cx<T>* dst = p.out + (tile / p.gx) * p.out_os + (tile % p.gx) * W + tout;
      const T sc = p.scale;
      if (sc != T(1)) {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(dst + p.out_off[m], sc * v[m], p.keep_out);
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(dst + p.out_off[m], v[m], p.keep_out);
      }
    }
  }
}

}  // namespace ffb
