// Lean tiles for the two sub-passes of a four-step transform of long strided lines (N = N1*N2, fft_plan.cu:fourstep_min):
//   A: N1-point transforms over n1 (element stride N2*inner) for fixed n2, output k1 in place of n1, times exp(-/+2 pi i n2 k1/N)
//   B: N2-point transforms over n2 for fixed k1, output index k1 + N1*k2
// Used by the stand-alone sub-pass kernels below (one CTA per tile) and by the persistent fused kernel of fft_l2four.cuh
// (intermediate kept in L2).  Compared with the general strided pass of fft_pow2.cuh (segmented strides, runtime fusion hooks:
// ~4100 SASS instructions, 128 registers in Float64) these tiles
//   * get every tile-dependent address from the caller and add compile-time multiples of two strides: ~750 instructions;
//   * hold R = 8 points per thread in Float64 (32 data registers, <= 80 registers): 24 warps per SM instead of 16 -- the
//     Float64 passes are latency-bound (ncu: issue slots 37 % busy, barrier / long-scoreboard stalls), not bandwidth-bound;
//   * compile the calcN! fusion hooks (spectral factor on load, accumulate + dealias on store) in or out (PRO / EPI).
#pragma once
#include "fft_pow2.cuh"

namespace ffb {

template <typename T>
struct FsParams {
  long long in_es, out_es;     // element (transform index) strides of input / output, in complex elements
  long long in_ms, out_ms;     // Tn * in_es, Tn * out_es: stride between a thread's consecutive registers
  int W, lgW;                  // columns per tile (power of two)
  T scale;                     // applied to the output when != 1
  const cx<T>* tw;             // base twiddles of this sub-transform (fft_pow2.cuh run_passes)
  const cx<T>* twN;            // A: exp(-2 pi i q / N), q < N
  int twN_mask;                // N - 1
  typename Pow2Params<T>::Fuse hook;   // PRO (A) or EPI (B) operands; .w / .acc are indexed like the true array
  DeadCols dead;               // forward transforms followed by dealias!: A skips tiles of aliased columns, B writes their zeros unread
};

struct FsTile {
  long long in_off, out_off;   // offsets of the tile origin (column line0, transform index 0) in the input / output array
  long long hook_off;          // the same origin in the TRUE array the hook operands are laid out like (A: input, B: output)
  long long line0;             // global column index of the tile's first column
  long long o_hi;              // outer slice index
  int o_lo;                    // A: n2, B: k1
  int ncols;                   // active columns of this tile (ragged last tile)
};

// `sched.before_store()` runs when the transform is done and nothing of this thread is outstanding any more: the persistent fused
// kernel publishes the previous tile there (a fence at that point returns at once) and looks at the next tile's dependency.
struct FsNoSched { FFB_D void before_store() const {} };

template <typename T, int DIR, bool IS_A, bool HOOK, bool IN_CG, bool OUT_KEEP, int R, int... Rs, class Sched = FsNoSched>
FFB_D void fs_tile(const FsParams<T>& p, const cx<T>* pin, cx<T>* pout, const FsTile& tl, Sched sched = Sched()) {
  constexpr int N = radix_product<Rs...>::value;
  constexpr int Tn = N / R;
  static_assert(N % R == 0, "R must divide N");
  using XW = typename xword<T, true>::type;
  extern __shared__ __align__(16) unsigned char ffb_smem[];
  XW* xb = reinterpret_cast<XW*>(ffb_smem);
  const int W = p.W;
  const int tid = threadIdx.x;
  const int w = tid & (W - 1);
  const int t = tid >> p.lgW;
  const bool active = w < tl.ncols;
  if (cols_all_dead(p.dead, tl.line0, tl.ncols)) {   // uniform over the CTA
    sched.before_store();
    if constexpr (!IS_A) {
      if (p.dead.on == 2 && active) {
        cx<T>* out = pout + tl.out_off + w + (long long)t * p.out_es;
#pragma unroll
        for (int m = 0; m < R; ++m) stc(out + (long long)m * p.out_ms, mk<T>(0, 0));
      }
    }
    return;
  }
  if constexpr (HOOK && !IS_A) {
    // the accumulated array is consumed after the transform: pull the live part of its tile towards L2 now, or its loads would be a
    // second exposed DRAM round trip (one request per 128-byte line; a warp's lanes cover adjacent columns)
    const typename Pow2Params<T>::Fuse& h = p.hook;
    if (active && h.acc) {
      const unsigned line = (unsigned)(tl.line0 + w);
      unsigned ci0, cio;
      col_coords(line, (unsigned)h.n0, ci0, cio);
      const int i0 = (int)ci0;
      const long long io = h.other_from_col == 1 ? (long long)cio : (h.other_from_col == 2 ? (long long)tl.o_lo : tl.o_hi);
      const bool dead0 = h.dealias && ((h.lo0 > 0 && i0 >= h.lo0 - 1 && i0 < h.hi0) || (h.loo > 0 && io >= h.loo - 1 && io < h.hio));
      if (!dead0) {
        const cx<T>* ap = h.acc + tl.hook_off + w + (long long)t * p.out_es;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int it = h.idm * (t + m * Tn) + h.ido * tl.o_lo;
          const cx<T>* a = ap + (long long)m * p.out_ms;
          const bool live = !(h.dealias && h.lot > 0 && it >= h.lot - 1 && it < h.hit);
          if (live && (w == 0 || (reinterpret_cast<uintptr_t>(a) & 127) < sizeof(cx<T>))) prefetch_l2(a);
        }
      }
    }
  }
  cx<T> v[R];
  // ---------------- load ----------------
  const cx<T>* in = pin + tl.in_off + w + (long long)t * p.in_es;
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = active ? ldin<T, IN_CG>(in + (long long)m * p.in_ms) : mk<T>(0, 0);
  if constexpr (HOOK && IS_A) {
    // prologue: (cr + i ci) * k0[i0] * kt[it] * ko[io] * w[.] * x, evaluated left to right like `im * l * invKrsq * sol`.
    // Every operand load is issued before anything is consumed (the pass is latency-bound: ncu long_scoreboard), and the column
    // coordinates use 32-bit arithmetic (a 64-bit modulo is a subroutine call that the loads cannot be scheduled across).
    if (active) {
      const typename Pow2Params<T>::Fuse& h = p.hook;
      const unsigned line = (unsigned)(tl.line0 + w);
      const T* wp = h.w ? h.w + tl.hook_off + w + (long long)t * p.in_es : nullptr;
      // Float32 holds 16 points per thread under a 64-register cap: its transform-index factors are loaded where they are used
      constexpr bool EARLY_KT = sizeof(T) == 8;
      T wv[R], kq[EARLY_KT ? R : 1];
#pragma unroll
      for (int m = 0; m < R; ++m) wv[m] = wp ? __ldcs(wp + (long long)m * p.in_ms) : T(1);
      if constexpr (EARLY_KT) {
#pragma unroll
        for (int m = 0; m < R; ++m) kq[m] = h.kt ? __ldg(h.kt + (h.idm * (t + m * Tn) + h.ido * tl.o_lo)) : T(1);
      }
      unsigned ci0, cio;
      col_coords(line, (unsigned)h.n0, ci0, cio);
      const int i0 = (int)ci0;
      const long long io = h.other_from_col == 1 ? (long long)cio : (h.other_from_col == 2 ? (long long)tl.o_lo : tl.o_hi);
      T fr0 = h.cr, fi0 = h.ci;
      if (h.k0) { const T q = __ldg(h.k0 + i0); fr0 *= q; fi0 *= q; }
      const T qo = h.ko ? __ldg(h.ko + io) : T(1);
#pragma unroll
      for (int m = 0; m < R; ++m) {
        T fr = fr0, fi = fi0;
        if (h.kt) {
          const T q = EARLY_KT ? kq[EARLY_KT ? m : 0] : __ldg(h.kt + (h.idm * (t + m * Tn) + h.ido * tl.o_lo));
          fr *= q; fi *= q;
        }
        if (h.ko) { fr *= qo; fi *= qo; }
        if (wp) { fr *= wv[m]; fi *= wv[m]; }
        v[m] = mk<T>(fr, fi) * v[m];
      }
    }
  }
  // ---------------- transform ----------------
  run_passes<T, DIR, true, R, N, 1, 0, Rs...>(v, t, w, W, xb, p.tw);
  // ---------------- store ----------------
  if constexpr (IS_A) {
    // inter-pass twiddle exp(-/+2 pi i n2 k1 / N), k1 = t + m*Tn, n2 = o_lo (the same for every column of the tile)
    cx<T> wv[R];
#pragma unroll
    for (int m = 0; m < R; ++m) wv[m] = load_tw<T, DIR>(p.twN + ((tl.o_lo * (t + m * Tn)) & p.twN_mask));
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = v[m] * wv[m];
  }
  sched.before_store();
  if (!active) return;
  cx<T>* out = pout + tl.out_off + w + (long long)t * p.out_es;
  const T sc = p.scale;
  if constexpr (HOOK && !IS_A) {
    // epilogue: out = dealias( F * (sc * y) + G * acc ), F = (cr + i ci) k0 kt ko w, G = (ar + i ai) a0 at ao
    const typename Pow2Params<T>::Fuse& h = p.hook;
    const unsigned line = (unsigned)(tl.line0 + w);
    unsigned ci0, cio;
    col_coords(line, (unsigned)h.n0, ci0, cio);
    const int i0 = (int)ci0;
    const long long io = h.other_from_col == 1 ? (long long)cio : (h.other_from_col == 2 ? (long long)tl.o_lo : tl.o_hi);
    const bool dead0 = h.dealias && ((h.lo0 > 0 && i0 >= h.lo0 - 1 && i0 < h.hi0) || (h.loo > 0 && io >= h.loo - 1 && io < h.hio));
    const long long hb = tl.hook_off + w + (long long)t * p.out_es;
    // loop-invariant operands first (factor order as in fuse_factor: scalar, k0[i0], kt[it], ko[io], w)
    T fr0 = h.cr, fi0 = h.ci, gr0 = h.ar, gi0 = h.ai;
    if (h.k0) { const T q = __ldg(h.k0 + i0); fr0 *= q; fi0 *= q; }
    if (h.acc && h.a0) { const T q = __ldg(h.a0 + i0); gr0 *= q; gi0 *= q; }
    const T qo = h.ko ? __ldg(h.ko + io) : T(1);
    const T go = (h.acc && h.ao) ? __ldg(h.ao + io) : T(1);
    // Two halves: R accumulated values beside the R data points do not fit the register budget (the one-batch form spilled 40-72
    // bytes per thread); the accumulated tile is already on its way to L2 (prefetch at tile start).
    constexpr int H = R / 2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      cx<T> av[H];
      bool dead[H];
#pragma unroll
      for (int j = 0; j < H; ++j) {   // independent loads of the accumulated array first
        const int m = half * H + j;
        const int it = h.idm * (t + m * Tn) + h.ido * tl.o_lo;
        dead[j] = dead0 || (h.dealias && h.lot > 0 && it >= h.lot - 1 && it < h.hit);
        av[j] = (h.acc && !dead[j]) ? ldc(h.acc + hb + (long long)m * p.out_ms) : mk<T>(0, 0);
      }
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const int m = half * H + j;
        const int it = h.idm * (t + m * Tn) + h.ido * tl.o_lo;
        cx<T> r = mk<T>(0, 0);
        if (!dead[j]) {
          T fr = fr0, fi = fi0;
          if (h.kt) { const T q = __ldg(h.kt + it); fr *= q; fi *= q; }
          if (h.ko) { fr *= qo; fi *= qo; }
          if (h.w) { const T q = __ldcs(h.w + hb + (long long)m * p.out_ms); fr *= q; fi *= q; }
          r = mk<T>(fr, fi) * (sc * v[m]);
          if (h.acc) {
            T gr = gr0, gi = gi0;
            if (h.at) { const T q = __ldg(h.at + it); gr *= q; gi *= q; }
            if (h.ao) { gr *= go; gi *= go; }
            r = r + mk<T>(gr, gi) * av[j];
          }
        }
        // dealias = 2: the alias box is don't-care, nothing is stored there
        if (!(dead[j] && h.dealias == 2)) stc(out + (long long)m * p.out_ms, r);
      }
    }
  } else {
    if (sc != T(1)) {
#pragma unroll
      for (int m = 0; m < R; ++m) stk(out + (long long)m * p.out_ms, sc * v[m], OUT_KEEP ? 1 : 0);
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) stk(out + (long long)m * p.out_ms, v[m], OUT_KEEP ? 1 : 0);
    }
  }
}

// L2 prefetch of an A tile's input (and dense prologue factor): one request per 128-byte line
template <typename T, int R, int N>
FFB_D void fs_prefetch(const FsParams<T>& p, const cx<T>* pin, const FsTile& tl, bool with_w) {
  constexpr int Tn = N / R;
  const int w = threadIdx.x & (p.W - 1), t = threadIdx.x >> p.lgW;
  if (w >= tl.ncols) return;
  const cx<T>* in = pin + tl.in_off + w + (long long)t * p.in_es;
  if (!(((reinterpret_cast<uintptr_t>(in) & 127) < sizeof(cx<T>)) || w == 0)) return;
#pragma unroll
  for (int m = 0; m < R; ++m) prefetch_l2(in + (long long)m * p.in_ms);
  if (with_w && p.hook.w) {
    const T* wp = p.hook.w + tl.hook_off + w + (long long)t * p.in_es;
#pragma unroll
    for (int m = 0; m < R; ++m) prefetch_l2(wp + (long long)m * p.in_ms);
  }
}

// radix plan of a sub-transform as a type: the fused kernel is a template over two of them
template <int R_, int... Rs> struct FsPlan {
  static constexpr int R = R_;
  static constexpr int N = radix_product<Rs...>::value;
  template <typename T, int DIR, bool IS_A, bool HOOK, bool IN_CG, bool OUT_KEEP, class Hook>
  static FFB_D void tile(const FsParams<T>& p, const cx<T>* pin, cx<T>* pout, const FsTile& tl, Hook hook) {
    fs_tile<T, DIR, IS_A, HOOK, IN_CG, OUT_KEEP, R_, Rs...>(p, pin, pout, tl, hook);
  }
  template <typename T>
  static FFB_D void prefetch(const FsParams<T>& p, const cx<T>* pin, const FsTile& tl, bool with_w) { fs_prefetch<T, R_, N>(p, pin, tl, with_w); }
};

// Float64: 8 points per thread; Float32: 16 (32 data registers either way)
template <typename T, int N> struct fs_plan_for;
template <> struct fs_plan_for<double, 32> { using type = FsPlan<8, 8, 4>; };
template <> struct fs_plan_for<double, 64> { using type = FsPlan<8, 8, 8>; };
template <> struct fs_plan_for<double, 128> { using type = FsPlan<8, 8, 8, 2>; };
template <> struct fs_plan_for<double, 256> { using type = FsPlan<8, 8, 8, 4>; };
template <> struct fs_plan_for<float, 32> { using type = FsPlan<16, 16, 2>; };
template <> struct fs_plan_for<float, 64> { using type = FsPlan<16, 16, 4>; };
template <> struct fs_plan_for<float, 128> { using type = FsPlan<16, 16, 8>; };
template <> struct fs_plan_for<float, 256> { using type = FsPlan<16, 16, 16>; };

// ---------------------------------------------------------------- stand-alone sub-pass kernels (two-kernel four-step)
// grid: x = column tile, y = o_lo + mod * o_hi
template <typename T>
struct FsLaunch {
  FsParams<T> p;
  const cx<T>* in;
  cx<T>* out;
  long long in_os, in_os2, out_os, out_os2;   // strides of o_lo / o_hi
  long long nlines;
  int mod;                                    // N2 (A) or N1 (B)
};

template <typename T, int DIR, bool IS_A, bool HOOK, class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fs_pass_kernel(const __grid_constant__ FsLaunch<T> q) {
  FsTile tl;
  tl.o_lo = (int)(blockIdx.y % (unsigned)q.mod);
  tl.o_hi = blockIdx.y / (unsigned)q.mod;
  tl.line0 = (long long)blockIdx.x * q.p.W;
  tl.ncols = (int)min((long long)q.p.W, q.nlines - tl.line0);
  tl.in_off = tl.o_lo * q.in_os + tl.o_hi * q.in_os2 + tl.line0;
  tl.out_off = tl.o_lo * q.out_os + tl.o_hi * q.out_os2 + tl.line0;
  tl.hook_off = IS_A ? tl.in_off : tl.out_off;
  PL::template tile<T, DIR, IS_A, HOOK, false, false>(q.p, q.in, q.out, tl, FsNoSched());
}

// Sub-pass A of several inverse transforms of the SAME spectral array that differ only in their prologue factor (calcN! of the 2-D
// vorticity equation: zeta, u and v all come from `sol`).  One CTA runs the variants of its tile back to back: the first one pulls
// the tile (and a dense factor shared between variants) from DRAM, the others find it in L1 / L2, so the input is read once
// instead of `nv` times.  Tile geometry comes from q[0]; q[v] carries variant v's output and prologue.
constexpr int kFsMaxVariants = 4;
template <typename T>
struct FsMultiLaunch {
  FsLaunch<T> q[kFsMaxVariants];
  int nv;
};

template <typename T, int DIR, class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fs_pass_multi_kernel(const __grid_constant__ FsMultiLaunch<T> mq) {
  const FsLaunch<T>& q0 = mq.q[0];
  FsTile tl;
  tl.o_lo = (int)(blockIdx.y % (unsigned)q0.mod);
  tl.o_hi = blockIdx.y / (unsigned)q0.mod;
  tl.line0 = (long long)blockIdx.x * q0.p.W;
  tl.ncols = (int)min((long long)q0.p.W, q0.nlines - tl.line0);
  tl.in_off = tl.o_lo * q0.in_os + tl.o_hi * q0.in_os2 + tl.line0;
  tl.out_off = tl.o_lo * q0.out_os + tl.o_hi * q0.out_os2 + tl.line0;
  tl.hook_off = tl.in_off;
  // dense factors of the later variants: on their way from DRAM to L2 while variant 0 runs (one request per 128-byte line)
  {
    const int w = threadIdx.x & (q0.p.W - 1), t = threadIdx.x >> q0.p.lgW;
    const T* prev = nullptr;
#pragma unroll 1
    for (int v = 1; v < mq.nv; ++v) {
      const T* wq = mq.q[v].p.hook.w;
      if (wq && wq != prev && w < tl.ncols) {
        const T* a = wq + tl.hook_off + w + (long long)t * q0.p.in_es;
#pragma unroll
        for (int m = 0; m < PL::R; ++m)
          if (w == 0 || (reinterpret_cast<uintptr_t>(a + (long long)m * q0.p.in_ms) & 127) < sizeof(T)) prefetch_l2(a + (long long)m * q0.p.in_ms);
      }
      prev = wq;
    }
  }
  // all but the last variant load with .cg (normal L2 priority: the tile is about to be read again), the last one streams (.cs)
#pragma unroll 1
  for (int v = 0; v + 1 < mq.nv; ++v) {
    const FsLaunch<T>& q = mq.q[v];
    PL::template tile<T, DIR, true, true, true, false>(q.p, q.in, q.out, tl, FsNoSched());
    __syncthreads();   // the exchange buffer is reused: every gather of this variant is done
  }
  const FsLaunch<T>& q = mq.q[mq.nv - 1];
  PL::template tile<T, DIR, true, true, false, false>(q.p, q.in, q.out, tl, FsNoSched());
}

}  // namespace ffb
