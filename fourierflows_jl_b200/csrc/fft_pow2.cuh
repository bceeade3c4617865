// Register-resident Stockham FFT for power-of-two line lengths (sm_100a).
//
// One thread owns R points of one line for the whole transform; data stays in registers, shared memory is used
// only for the inter-pass exchange (a pure permutation), so a line makes exactly one HBM round trip.
//   pass p (radix r, Ns = product of earlier radices), butterfly j = t + b*N/R, b < R/r:
//     in : x[j + k*N/r]                      = register  b + k*R/r      (the same registers in every pass)
//     out: y[(j/Ns)*Ns*r + (j%Ns) + k*Ns]    -> scattered store to shared memory, then re-read at t + m*N/R
//   first pass loads global memory at t + m*N/R, last pass lands on the same positions: both coalesced.
// Exchange words are 8 bytes (Float64 rows: re and im in two phases; Float32: one float2 phase; Float64 strided passes move
// whole 16-byte complex values in one phase) padded by one word
// every 16, which the design model (tools/fft_model.py) shows to be bank-conflict free for every plan below.
//
// Modes (replace the cuFFT/FFTW plans of src/domains.jl:2-5 executed by mul!/ldiv!, src/diffusion.jl:137-139):
//   C2C_ROWS : contiguous lines (x-dimension of `fftplan`, 1-D grids)
//   C2C_COLS : strided lines; a CTA owns W adjacent columns so every HBM access is W*sizeof(complex) wide
//   R2C      : real line of 2N reals -> N+1 complex (half spectrum), fused post-process
//   C2R      : N+1 complex -> 2N reals, fused pre-process and 1/(nx*ny*nz) scaling
#pragma once
#include <type_traits>
#include <utility>
#include "fft_radix.cuh"

namespace ffb {

// C2C_COLS_TW: first half of a four-step transform of a long strided line (N = N1*N2): length-N1 sub-transform over n1
// for fixed n2 (= blockIdx.y % outer_mod), output k1 multiplied by the inter-pass twiddle exp(-/+2*pi*i*n2*k1/N).
// C2C_COLS_LEAN: C2C_COLS without fusion hooks / peer stores, element offsets precomputed on the host (in_off / out_off):
// the plain strided passes are instruction-issue bound, this variant executes ~30 % fewer instructions.
enum FftMode { C2C_ROWS = 0, C2C_COLS = 1, R2C_ROWS = 2, C2R_ROWS = 3, C2C_COLS_TW = 4, C2C_COLS_LEAN = 5 };

// Columns of a forward transform that `dealias!` zeroes afterwards (the aliased kx range, and for a pass whose columns are
// (kx, ky) pairs the aliased ky range) need neither be stored, transformed nor exchanged: intermediate passes skip tiles made of
// such columns only, the last pass writes their zeros without loading anything.  Column c: i0 = c % n0, other = c / n0;
// dead iff i0 in [dlo, dhi) or other in [olo, ohi) (0-based, half open; dhi = 0 / ohi = 0: no range).
struct DeadCols {
  int on;        // 0 = off, 1 = skip all-dead tiles, 2 = final pass: all-dead tiles are zero-filled
  int n0, dlo, dhi, olo, ohi;
  // blocked passes of the slab decomposition (C2C_COLS_LEAN): outer slices `by` in [bylo, byhi) are dead; output elements whose
  // transform index is in [tlo, thi) are not stored (the pass that feeds the exchange does not send aliased modes)
  int bylo, byhi, tlo, thi;
};
// Column coordinates (i0, other) of line `c`: lines are < 2^31 (host check), so 32-bit arithmetic -- a 64-bit division is a
// subroutine call in front of every tile's loads -- and no division at all when the pass has a single row of columns (2-D).
FFB_D void col_coords(unsigned c, unsigned n0, unsigned& i0, unsigned& other) {
  if (c < n0) { i0 = c; other = 0; }
  else { other = c / n0; i0 = c - other * n0; }
}
FFB_D bool cols_all_dead(const DeadCols& d, long long first, int ncols) {
  if (!d.on || ncols <= 0) return false;
  unsigned a, b, r0, r1;
  col_coords((unsigned)first, (unsigned)d.n0, a, r0);
  col_coords((unsigned)first + (unsigned)ncols - 1u, (unsigned)d.n0, b, r1);
  if (d.ohi > d.olo && (int)r0 >= d.olo && (int)r1 < d.ohi) return true;
  if (d.dhi > d.dlo && r0 == r1) return (int)a >= d.dlo && (int)b < d.dhi;
  return false;
}

template <typename T>
struct Pow2Params {
  const void* in;
  void* out;
  long long in_es, in_ls, in_os;     // element / line / outer strides of the input, in complex elements
  long long out_es, out_ls, out_os;  // same for the output
  // Segmented element stride (slab-decomposed transforms): element i lives at (i & seg_mask)*es + (i >> seg_shift)*seg_stride,
  // so a pass can read / write destination-rank-major blocks without a pack kernel.  Unsegmented: mask = ~0, shift = 31.
  int in_seg_mask, in_seg_shift, out_seg_mask, out_seg_shift;
  long long in_seg_stride, out_seg_stride;
  // two-level outer index: blockIdx.y = o -> (o % outer_mod)*os + (o / outer_mod)*os2   (outer_mod = 1<<30 when unused)
  int outer_mod;
  long long in_os2, out_os2;
  // Fused elementwise work (north star: spectral multiplies, products and the dealias mask folded into the passes).
  // Coordinates of an element of a strided pass: i0 = col % n0 (kx index), it = transform index
  // (= idm*(t + m*N/16) + ido*o_lo), io = "other" index (col / n0 for a z-pass, o_hi for a y-pass).
  struct Fuse {
    int on;                      // 0 = none
    T cr, ci;                    // complex scalar
    const T* k0; const T* kt; const T* ko;   // optional wavenumber factors by i0 / it / io
    const T* w;                  // optional dense real factor, same layout as the array (pre-offset like in / out)
    const cx<T>* acc;            // epilogue only: out = f*own + g*acc[idx]
    T ar, ai; const T* a0; const T* at; const T* ao;
    int dealias; int lo0, hi0, lot, hit, loo, hio;   // 1-based inclusive alias ranges on i0 / it / io (lo = 0: none)
    int n0, idm, ido, other_from_col;   // other_from_col: 0 = o_hi, 1 = col / n0, 2 = o_lo
  };
  Fuse pro, epi;
  const T* rmul;                     // C2R_ROWS: real result multiplied by this real field (same layout as out)
  int rsq;                           // R2C_ROWS: the real input is squared on load (`@. c = c * c` folded into the transform)
  // R2C_ROWS store / C2R_ROWS load of a half spectrum cut into blocks of 2^row_seg_shift wavenumbers (2-D slab decomposition: block q
  // belongs to rank q; the Nyquist wavenumber N rides with the last block): element k lives at (k & mask) + (k >> shift)*stride,
  // k = N at row_nyq.  row_seg = 0: plain contiguous half spectrum.
  int row_seg, row_seg_mask, row_seg_shift;
  long long row_seg_stride, row_nyq;
  int row_dead_lo, row_dead_hi;      // R2C_ROWS: half-spectrum elements k in [lo, hi) are not stored (dealiased later; see DeadCols)
  DeadCols dead;                     // strided modes: see DeadCols
  int pf_ahead;                      // ROWS modes: prefetch the line this many tiles ahead into L2 (0 = off)
  int reverse;                       // walk tiles / slices backwards (snake ordering between consecutive passes: the tail
                                     // of what the previous kernel wrote is still in the 126 MB L2)
  int keep_out;                      // 1: store with the default policy (data re-read by the next pass), 0: evict-first
  const cx<T>* twN;                  // C2C_COLS_TW: exp(-2*pi*i*q/Nfull), q < Nfull
  int twN_mask;                      // Nfull - 1
  long long nlines;                  // lines per outer index
  int W;                             // lines per CTA
  T scale;                           // applied to the output when != 1 (inverse normalisation)
  const cx<T>* tw;                   // base twiddles, forward sign: for each pass with Ns > 1, exp(-2 pi i a/(Ns r)), a < Ns
  const cx<T>* twr;                  // exp(-i*pi*k/N), k < N, for the r2c / c2r split step
  // C2C_COLS_LEAN: tile bx starts at bx*ts, register m's element lives at in + in_off[m] / out_m[m] (segmented strides and,
  // for the fused pass + collective of the slab decomposition, the destination rank's receive buffer -- peer memory mapped
  // through CUDA IPC, NVLink stores -- are folded into these by the host)
  long long in_ts, out_ts;
  long long in_off[16];
  cx<T>* out_m[16];
};

template <int... Rs> struct radix_product;
template <> struct radix_product<> { static constexpr int value = 1; };
template <int r, int... Rs> struct radix_product<r, Rs...> { static constexpr int value = r * radix_product<Rs...>::value; };

__host__ __device__ constexpr int ce_log2(int v) { return v <= 1 ? 0 : 1 + ce_log2(v >> 1); }

FFB_D int xpad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int xpad_len(int n) { return n + (n >> 4) + 1; }

// ---- exchange word access: F64 moves re / im separately, F32 moves the whole float2 ----
// Float64 moves whole complex values (16-byte words) in ONE phase, Float32 one float2: half the barriers and half the shared-memory
// instructions of the re / im two-phase form round 1 used (ncu: the Float64 passes stall on barriers and the MIO queue, not on
// DRAM).  The index pattern is conflict free for any word size: strided passes put adjacent columns (contiguous words) on a warp's
// lanes; contiguous-line passes scatter with a stride of 17 words (the pad) in the first pass and to consecutive words afterwards.
// -DFFB_ROWS_TWO_PHASE builds the round-1 form of the contiguous-line Float64 passes (re and im as 8-byte words in two phases) for
// A/B timing of the two library builds on one box (FFB_LIB_PATH).
template <typename T, bool COLS> struct xword;
#ifdef FFB_ROWS_TWO_PHASE
template <> struct xword<double, false> { using type = double; static constexpr int phases = 2; };
template <> struct xword<double, true> { using type = double2; static constexpr int phases = 1; };
#else
template <bool COLS> struct xword<double, COLS> { using type = double2; static constexpr int phases = 1; };
#endif
template <bool COLS> struct xword<float, COLS> { using type = float2; static constexpr int phases = 1; };

template <typename T, bool COLS, int PH> FFB_D typename xword<T, COLS>::type xget(const cx<T>& c) {
  if constexpr (sizeof(T) == 8 && sizeof(typename xword<T, COLS>::type) == 8) return PH == 0 ? c.x : c.y;
  else if constexpr (sizeof(T) == 8) return make_double2(c.x, c.y);
  else return make_float2(c.x, c.y);
}
template <typename T, bool COLS, int PH> FFB_D void xput(cx<T>& c, typename xword<T, COLS>::type w) {
  if constexpr (sizeof(T) == 8 && sizeof(typename xword<T, COLS>::type) == 8) { if (PH == 0) c.x = w; else c.y = w; }
  else { c.x = w.x; c.y = w.y; }
}

template <int I, int N, typename F> FFB_D void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Scatter the outputs of a pass into the exchange buffer and gather the inputs of the next pass.
// COLS: word address = xpad(idx)*W + w (columns interleaved);  ROWS: w*xpad_len(N) + xpad(idx).
// The long strided Float32 passes are issue-bound (ncu: ~110 instructions per point against a budget of 88 at the HBM
// roofline), so the padded index is split into a per-thread base, computed once, plus compile-time constants: adding a
// multiple of 16 commutes with xpad, and xpad(16*j + k) = 17*j + k for k < 16.
template <typename T, bool COLS, int R, int N, int Ns, int r>
FFB_D void exchange(cx<T> (&v)[R], int t, int w, int W, typename xword<T, COLS>::type* xb) {
  constexpr int Tn = N / R, nb = R / r;
  constexpr int lNs = ce_log2(Ns), lr = ce_log2(r);
  constexpr bool kfast = (Ns % 16 == 0) || (Ns == 1 && r == 16);   // k*Ns moves the padded index by a constant
  constexpr bool mfast = (Tn % 16 == 0);                           // so does m*Tn
  auto lin = [&](int xp) { return COLS ? xp * W + w : w * xpad_len(N) + xp; };   // padded index -> word address
  auto step = [&](int c) { return COLS ? c * W : c; };                           // constant padded-index offset -> words
  int sb[nb];
#pragma unroll
  for (int b = 0; b < nb; ++b) {
    const int j = t + b * Tn;
    const int base = ((j >> lNs) << (lNs + lr)) + (j & (Ns - 1));
    sb[b] = (Ns == 1 && r == 16) ? lin(17 * j) : lin(xpad(base));
  }
  const int gb = lin(xpad(t));
  static_for<0, xword<T, COLS>::phases>([&](auto PH) {
    constexpr int ph = decltype(PH)::value;
    // previous gather finished before the buffer is overwritten.  Not needed before the first scatter of a tile (Ns == 1, first
    // phase): nothing of this tile has touched the buffer yet, and a persistent kernel ends every tile with a barrier.
    if constexpr (!(Ns == 1 && ph == 0)) __syncthreads();
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const int j = t + b * Tn;
      const int base = ((j >> lNs) << (lNs + lr)) + (j & (Ns - 1));
#pragma unroll
      for (int k = 0; k < r; ++k) {
        const int a = kfast ? sb[b] + step(Ns == 1 ? k : (k * Ns / 16) * 17) : lin(xpad(base + k * Ns));
        xb[a] = xget<T, COLS, ph>(v[b + k * nb]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int a = mfast ? gb + step((m * Tn / 16) * 17) : lin(xpad(t + m * Tn));
      xput<T, COLS, ph>(v[m], xb[a]);
    }
  });
}

template <typename T, int DIR> FFB_D cx<T> load_tw(const cx<T>* p) {
  using V = typename vec2<T>::type;
  V q = __ldg(reinterpret_cast<const V*>(p));
  return mk<T>(q.x, DIR < 0 ? q.y : -q.y);
}

// Multiply v[b + k*nb] by w1^k, k = 1..r-1, using the binary expansion of k: only w1, w1^2, w1^4, w1^8 are formed
// (squarings), so a butterfly needs ONE table load instead of r-1 and at most log2(r) roundings per twiddle.
template <typename T, int R, int r, int b> FFB_D void apply_twiddle_powers(cx<T> (&v)[R], cx<T> w) {
  constexpr int nb = R / r;
#pragma unroll
  for (int bit = 1; bit < r; bit <<= 1) {
#pragma unroll
    for (int k = 1; k < r; ++k)
      if (k & bit) v[b + k * nb] = v[b + k * nb] * w;
    if (bit * 2 < r) w = mk<T>(w.x * w.x - w.y * w.y, (w.x + w.x) * w.y);
  }
}

// All passes of one line, recursively over the radix pack.  TWOFF = offset of this pass's base twiddles
// (tw[TWOFF + a] = exp(-2*pi*i*a/(Ns*r)), a < Ns) in the table.
template <typename T, int DIR, bool COLS, int R, int N, int Ns, int TWOFF, int r, int... Rest>
FFB_D void run_passes(cx<T> (&v)[R], int t, int w, int W, typename xword<T, COLS>::type* xb, const cx<T>* tw) {
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
    exchange<T, COLS, R, N, Ns, r>(v, t, w, W, xb);
    run_passes<T, DIR, COLS, R, N, Ns * r, TWOFF + (Ns > 1 ? Ns : 0), Rest...>(v, t, w, W, xb, tw);
  }
}

// exp(-i*pi*m/16), m = 0..15: the split-step twiddle of point k = t + m*N/16 is twr[t] times this constant
template <typename T> struct split_consts {
  static constexpr T c[16] = {T(1.0L), T(0.98078528040323044912618223613424L), T(0.92387953251128675612818318939679L),
                              T(0.83146961230254523707878837761791L), T(0.70710678118654752440084436210485L),
                              T(0.55557023301960222474283081394853L), T(0.38268343236508977172845998403040L),
                              T(0.19509032201612826784828486847702L), T(0.0L), T(-0.19509032201612826784828486847702L),
                              T(-0.38268343236508977172845998403040L), T(-0.55557023301960222474283081394853L),
                              T(-0.70710678118654752440084436210485L), T(-0.83146961230254523707878837761791L),
                              T(-0.92387953251128675612818318939679L), T(-0.98078528040323044912618223613424L)};
};
// twr[t + m*Tn] for the thread's 16 (or R) points; R == 16 uses one load and constants, smaller R loads the table
template <typename T, int R, int N, int M> FFB_D cx<T> split_twiddle(const cx<T>* twr, cx<T> base, int t) {
  constexpr int Tn = N / R;
  if constexpr (R == 16) {
    if constexpr (M == 0) return base;
    else {
      constexpr T c = split_consts<T>::c[M], s = -split_consts<T>::c[(M + 8) & 15] * ((M + 8) >= 16 ? T(-1) : T(1));
      // exp(-i*pi*M/16) = cos(pi M/16) - i sin(pi M/16);  sin(pi M/16) = cos(pi (M-8)/16) = -cos(pi (M+8)/16)
      return mk<T>(base.x * c - base.y * (-s), base.x * (-s) + base.y * c);
    }
  } else {
    return load_tw<T, -1>(twr + t + M * Tn);
  }
}

// Field data is touched once per pass: stream it (evict-first) so L1/L2 keep the twiddle tables.
template <typename T> FFB_D cx<T> ldc(const cx<T>* p) {
  using V = typename vec2<T>::type;
  V q = __ldcs(reinterpret_cast<const V*>(p));
  return mk<T>(q.x, q.y);
}
template <typename T> FFB_D void stc(cx<T>* p, cx<T> c) {
  using V = typename vec2<T>::type;
  V q; q.x = c.x; q.y = c.y;
  __stcs(reinterpret_cast<V*>(p), q);
}
template <typename T> FFB_D void stk(cx<T>* p, cx<T> c, int keep) {
  using V = typename vec2<T>::type;
  V q; q.x = c.x; q.y = c.y;
  if (keep) *reinterpret_cast<V*>(p) = q; else __stcs(reinterpret_cast<V*>(p), q);
}

template <typename T> FFB_D cx<T> ldg_cg(const cx<T>* p) {
  using V = typename vec2<T>::type;
  V q = __ldcg(reinterpret_cast<const V*>(p));
  return mk<T>(q.x, q.y);
}
template <typename T, bool CG> FFB_D cx<T> ldin(const cx<T>* p) {
  if constexpr (CG) return ldg_cg(p); else return ldc(p);
}

FFB_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// factor (cr + i ci) * k0[i0] * kt[it] * ko[io] * w[off], evaluated left to right like `im * kr * invKrsq * sol`
template <typename T>
FFB_D cx<T> fuse_factor(T cr, T ci, const T* k0, const T* kt, const T* ko, const T* w, int i0, int it, long long io, long long off) {
  T fr = cr, fi = ci;
  if (k0) { const T v = __ldg(k0 + i0); fr *= v; fi *= v; }
  if (kt) { const T v = __ldg(kt + it); fr *= v; fi *= v; }
  if (ko) { const T v = __ldg(ko + io); fr *= v; fi *= v; }
  if (w) { const T v = __ldcs(w + off); fr *= v; fi *= v; }
  return mk<T>(fr, fi);
}

// One tile (W lines) of a pass.  (bx, by) = tile coordinates (blockIdx of the plain launch), gdx / gdy = extents of the tile
// grid, pin / pout = array bases (p.in / p.out for the plain launch; the persistent fused four-step kernel of fft_l2four.cuh
// substitutes its L2-resident scratch ring), nlines = first inactive line.  IN_CG: load the input with ld.global.cg (L2 only):
// data written by other CTAs of the SAME kernel must not be served from a stale L1 line.
struct NoHook { FFB_D void operator()() const {} };

// after_load() runs once the tile's global loads have been issued (before anything waits for them): the persistent kernel of
// fft_l2four.cuh publishes the previous tile there, so that the fence it needs overlaps this tile's load latency.
template <typename T, int DIR, int MODE, bool IN_CG, int R, int... Rs, class AfterLoad = NoHook>
FFB_D void fft_pow2_tile(const Pow2Params<T>& p, const unsigned bx, const unsigned by, const unsigned gdx, const unsigned gdy, const void* pin,
                         void* pout, const long long nlines, AfterLoad after_load = AfterLoad()) {
  constexpr int N = radix_product<Rs...>::value;
  constexpr int Tn = N / R;
  constexpr bool COLS = (MODE == C2C_COLS || MODE == C2C_COLS_TW || MODE == C2C_COLS_LEAN);
  constexpr bool LEAN = (MODE == C2C_COLS_LEAN);
  static_assert(N % R == 0, "R must divide N");
  using XW = typename xword<T, COLS>::type;
  extern __shared__ __align__(16) unsigned char ffb_smem[];
  XW* xb = reinterpret_cast<XW*>(ffb_smem);

  const int W = p.W;
  const int tid = threadIdx.x;
  const int w = COLS ? tid % W : tid / Tn;
  const int t = COLS ? tid / W : tid % Tn;
  const long long line = (long long)bx * W + w;
  const bool active = line < nlines;
  const int o_lo = (int)(by % (unsigned)p.outer_mod);
  const long long o_hi = by / (unsigned)p.outer_mod;

  // ---------------- fused operands: pull them towards L2 now, they are consumed only after the data has arrived ----------------
  // (ncu: without this the dense factor / accumulated array / multiplier field cost a second exposed DRAM round trip per CTA)
  if (active) {
    if constexpr (MODE == C2C_COLS || MODE == C2C_COLS_TW) {
      if (p.pro.on && p.pro.w) {
        const T* wp = p.pro.w + o_lo * p.in_os + o_hi * p.in_os2 + line * p.in_ls;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int i = t + m * Tn;
          prefetch_l2(wp + (long long)(i & p.in_seg_mask) * p.in_es + (long long)(i >> p.in_seg_shift) * p.in_seg_stride);
        }
      }
      if (p.epi.on && p.epi.acc) {
        const cx<T>* ap = p.epi.acc + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int i = t + m * Tn;
          prefetch_l2(ap + (long long)(i & p.out_seg_mask) * p.out_es + (long long)(i >> p.out_seg_shift) * p.out_seg_stride);
        }
      }
    }
    if constexpr (MODE == C2R_ROWS) {
      if (p.rmul) {
        const char* mp = reinterpret_cast<const char*>(reinterpret_cast<const cx<T>*>(p.rmul) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls);
#pragma unroll 1
        for (int off = t * 128; off < N * (int)sizeof(cx<T>); off += Tn * 128) prefetch_l2(mp + off);
      }
    }
  }
  if constexpr (COLS) {
    // dealias-aware forward transforms: a tile made of aliased columns only is skipped (intermediate pass) or zero-filled (last pass)
    const long long first = (long long)bx * W;
    if (cols_all_dead(p.dead, first, (int)min((long long)W, nlines - first)) ||
        (LEAN && p.dead.on && (int)by >= p.dead.bylo && (int)by < p.dead.byhi)) {   // uniform over the CTA
      if (p.dead.on == 2 && active) {
        if constexpr (LEAN) {
          const long long oo = (long long)by * p.out_os + (long long)bx * p.out_ts + w + (long long)t * p.out_es;
#pragma unroll
          for (int m = 0; m < R; ++m) stk(p.out_m[m] + oo, mk<T>(0, 0), p.keep_out);
        } else {
          cx<T>* out = reinterpret_cast<cx<T>*>(pout) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
#pragma unroll
          for (int m = 0; m < R; ++m) {
            const int i = t + m * Tn;
            stc(out + (long long)(i & p.out_seg_mask) * p.out_es + (long long)(i >> p.out_seg_shift) * p.out_seg_stride, mk<T>(0, 0));
          }
        }
      }
      return;
    }
  }
  cx<T> v[R];
  // position of half-spectrum element k within its line (R2C_ROWS store, C2R_ROWS load): plain, or cut into per-rank blocks
  auto rowpos = [&](int k) -> long long {
    if (!p.row_seg) return (long long)k;
    return k == N ? p.row_nyq : (long long)(k & p.row_seg_mask) + (long long)(k >> p.row_seg_shift) * p.row_seg_stride;
  };
  // ---------------- L2 prefetch of the tile a later CTA on this SM will load (contiguous lines only) ----------------
  // The row kernels are latency-bound at 2-4 CTAs/SM (ncu: long_scoreboard); pulling the line `pf_ahead` tiles ahead into
  // L2 now turns its DRAM miss into an L2 hit when that CTA starts.
  if constexpr (!COLS) {
    if (p.pf_ahead > 0) {
      const long long pl = line + (long long)p.pf_ahead * W;
      if (pl < nlines) {
        const char* base = reinterpret_cast<const char*>(reinterpret_cast<const cx<T>*>(pin) + pl * p.in_ls);
        constexpr int line_bytes = (MODE == C2R_ROWS ? (N + 1) : N) * (int)sizeof(cx<T>);
#pragma unroll 1
        for (int off = t * 128; off < line_bytes; off += Tn * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
      }
    }
  }
  // ---------------- load ----------------
  if constexpr (MODE == C2R_ROWS) {
    // Z[k] = (X[k] + conj(X[N-k])) + i*exp(+i*pi*k/N)*(X[k] - conj(X[N-k]))
    const cx<T>* in = reinterpret_cast<const cx<T>*>(pin) + o_lo * p.in_os + o_hi * p.in_os2 + line * p.in_ls;
    const cx<T> wbase = load_tw<T, -1>(p.twr + t);
    // all loads of X[k] first, then the partners X[N-k] in batches: independent loads are in flight together
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = active ? ldc(in + rowpos(t + m * Tn)) : mk<T>(0, 0);
    constexpr int CB = R < 8 ? R : 8;
    static_for<0, R / CB>([&](auto MB) {
      constexpr int m0 = decltype(MB)::value * CB;
      cx<T> bv[CB];
#pragma unroll
      for (int q = 0; q < CB; ++q) bv[q] = active ? ldc(in + rowpos(N - (t + (m0 + q) * Tn))) : mk<T>(0, 0);
      static_for<0, CB>([&](auto Q) {
        constexpr int m = m0 + decltype(Q)::value;
        const int k = t + m * Tn;
        cx<T> a = v[m], b = bv[m - m0];
        if (k == 0) { a.y = T(0); b.y = T(0); }  // c2r ignores Im X[0] and Im X[N] (FFTW / cuFFT / pocketfft convention)
        const cx<T> wk = conj(split_twiddle<T, R, N, m>(p.twr, wbase, t));
        const cx<T> s = a + conj(b), d = a - conj(b);
        v[m] = s + mul_i(wk * d);
      });
    });
  } else if constexpr (LEAN) {
    const cx<T>* in = reinterpret_cast<const cx<T>*>(pin) + (long long)by * p.in_os + (long long)bx * p.in_ts + w + (long long)t * p.in_es;
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = active ? ldin<T, IN_CG>(in + p.in_off[m]) : mk<T>(0, 0);
    // one CTA per SM runs its load / butterfly / store phases back to back: pull the tile that starts `pf_ahead` CTAs later
    // into L2 while this one computes, so that its load phase is an L2 hit
    if (p.pf_ahead > 0) {
      const long long lt = (long long)by * gdx + bx + p.pf_ahead;
      const long long pby = lt / gdx, pbx = lt % gdx;
      if (pby < gdy && pbx * W + w < nlines) {
        const cx<T>* pfp = reinterpret_cast<const cx<T>*>(pin) + pby * p.in_os + pbx * p.in_ts + w + (long long)t * p.in_es;
#pragma unroll
        for (int m = 0; m < R; ++m) prefetch_l2(pfp + p.in_off[m]);
      }
    }
  } else {
    const cx<T>* in = reinterpret_cast<const cx<T>*>(pin) + o_lo * p.in_os + o_hi * p.in_os2 + line * p.in_ls;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int i = t + m * Tn;
      v[m] = active ? ldin<T, IN_CG>(in + (long long)(i & p.in_seg_mask) * p.in_es + (long long)(i >> p.in_seg_shift) * p.in_seg_stride) : mk<T>(0, 0);
    }
    if constexpr (MODE == R2C_ROWS) {
      if (p.rsq) {   // a packed pair of reals: square each
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = mk<T>(v[m].x * v[m].x, v[m].y * v[m].y);
      }
    }
    after_load();
    if (p.pro.on && active) {
      unsigned ci0, cio;
      col_coords((unsigned)line, (unsigned)p.pro.n0, ci0, cio);
      const int i0 = (int)ci0;
      const long long io = p.pro.other_from_col == 1 ? (long long)cio : (p.pro.other_from_col == 2 ? (long long)o_lo : o_hi);
      const long long base = (in - reinterpret_cast<const cx<T>*>(pin));
      // the dense factor is loaded in batches (independent loads first, uses after) so its latency is paid once per batch
      constexpr int PB = R < 8 ? R : 8;
#pragma unroll
      for (int m0 = 0; m0 < R; m0 += PB) {
        T wv[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const int i = t + (m0 + q) * Tn;
          const long long off = base + (long long)(i & p.in_seg_mask) * p.in_es + (long long)(i >> p.in_seg_shift) * p.in_seg_stride;
          wv[q] = p.pro.w ? __ldcs(p.pro.w + off) : T(1);
        }
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const int i = t + (m0 + q) * Tn;
          cx<T> f = fuse_factor<T>(p.pro.cr, p.pro.ci, p.pro.k0, p.pro.kt, p.pro.ko, (const T*)nullptr, i0, p.pro.idm * i + p.pro.ido * o_lo, io, 0);
          if (p.pro.w) f = mk<T>(f.x * wv[q], f.y * wv[q]);
          v[m0 + q] = f * v[m0 + q];
        }
      }
    }
  }
  // ---------------- transform ----------------
  run_passes<T, DIR, COLS, R, N, 1, 0, Rs...>(v, t, w, W, xb, p.tw);
  // ---------------- store ----------------
  if constexpr (MODE == R2C_ROWS) {
    // X[k] = 1/2 [ (Z[k] + conj(Z[N-k])) - i*exp(-i*pi*k/N)*(Z[k] - conj(Z[N-k])) ],  k = 0..N
    // Z is staged once in shared memory (Float64: separate re / im planes) so the partner Z[N-k] can be read back.
    constexpr int PL = xword<T, false>::phases;
    const int plane = xpad_len(N) * W;
    auto addr = [&](int idx) { return w * xpad_len(N) + xpad(idx); };
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      xb[addr(t + m * Tn)] = xget<T, false, 0>(v[m]);
      if constexpr (PL == 2) xb[plane + addr(t + m * Tn)] = xget<T, false, 1>(v[m]);
    }
    __syncthreads();
    cx<T>* out = reinterpret_cast<cx<T>*>(pout) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
    const T half = T(0.5);
    const cx<T> wbase = load_tw<T, -1>(p.twr + t);
    static_for<0, R>([&](auto M) {
      constexpr int m = decltype(M)::value;
      const int k = t + m * Tn;
      const int kp = (N - k) & (N - 1);
      cx<T> zp;
      xput<T, false, 0>(zp, xb[addr(kp)]);
      if constexpr (PL == 2) xput<T, false, 1>(zp, xb[plane + addr(kp)]);
      const cx<T> wk = split_twiddle<T, R, N, m>(p.twr, wbase, t);
      const cx<T> s = v[m] + conj(zp), d = v[m] - conj(zp);
      const cx<T> x = half * (s + mul_mi(wk * d));
      if (active) {
        if (!(k >= p.row_dead_lo && k < p.row_dead_hi)) stk(out + rowpos(k), x, p.keep_out);
        if (k == 0 && !(N >= p.row_dead_lo && N < p.row_dead_hi)) stk(out + rowpos(N), mk<T>(v[m].x - v[m].y, T(0)), p.keep_out);
      }
    });
  } else if constexpr (MODE == C2R_ROWS) {
    if (active) {
      cx<T>* out = reinterpret_cast<cx<T>*>(pout) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
      const T sc = p.scale;
      if (p.rmul) {
        const cx<T>* mulp = reinterpret_cast<const cx<T>*>(p.rmul) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
        constexpr int MB = R < 4 ? R : 4;
#pragma unroll
        for (int m0 = 0; m0 < R; m0 += MB) {
          cx<T> z[MB];
#pragma unroll
          for (int q = 0; q < MB; ++q) z[q] = ldc(mulp + (t + (m0 + q) * Tn));   // two consecutive reals of the multiplier field
#pragma unroll
          for (int q = 0; q < MB; ++q) stc(out + (t + (m0 + q) * Tn), mk<T>(sc * v[m0 + q].x * z[q].x, sc * v[m0 + q].y * z[q].y));
        }
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) stc(out + (t + m * Tn), sc * v[m]);
      }
    }
  } else if constexpr (LEAN) {
    if (active) {
      const long long oo = (long long)by * p.out_os + (long long)bx * p.out_ts + w + (long long)t * p.out_es;
      const T sc = p.scale;
      if (p.epi.on) {
        // epilogue of the blocked strided passes (slab decomposition): out = dealias( (cr + i ci) * k0[line] * ko[by] * kt[it] * y ),
        // coordinates: line = tile column (kx), by = outer slice, it = transform index
        const typename Pow2Params<T>::Fuse& h = p.epi;
        const int i0 = (int)line, io = (int)by;
        const bool dead0 = h.dealias && ((h.lo0 > 0 && i0 >= h.lo0 - 1 && i0 < h.hi0) || (h.loo > 0 && io >= h.loo - 1 && io < h.hio));
        T fr0 = h.cr, fi0 = h.ci;
        if (h.k0) { const T q = __ldg(h.k0 + i0); fr0 *= q; fi0 *= q; }
        if (h.ko) { const T q = __ldg(h.ko + io); fr0 *= q; fi0 *= q; }
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int it = t + m * Tn;
          const bool dead = dead0 || (h.dealias && h.lot > 0 && it >= h.lot - 1 && it < h.hit);
          T fr = fr0, fi = fi0;
          if (h.kt) { const T q = __ldg(h.kt + it); fr *= q; fi *= q; }
          stk(p.out_m[m] + oo, dead ? mk<T>(0, 0) : mk<T>(fr, fi) * (sc * v[m]), p.keep_out);
        }
      } else if (p.dead.on && p.dead.thi > p.dead.tlo) {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int it = t + m * Tn;
          if (!(it >= p.dead.tlo && it < p.dead.thi)) stk(p.out_m[m] + oo, sc * v[m], p.keep_out);
        }
      } else if (sc != T(1)) {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(p.out_m[m] + oo, sc * v[m], p.keep_out);
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(p.out_m[m] + oo, v[m], p.keep_out);
      }
    }
  } else {
    if constexpr (MODE == C2C_COLS_TW) {
      // inter-pass twiddle of the four-step split: k1 = t + m*Tn, n2 = o_lo (same value for every column of the tile)
      constexpr int TB = R < 4 ? R : 4;
#pragma unroll
      for (int m0 = 0; m0 < R; m0 += TB) {
        cx<T> wv[TB];
#pragma unroll
        for (int q = 0; q < TB; ++q) wv[q] = load_tw<T, DIR>(p.twN + ((o_lo * (t + (m0 + q) * Tn)) & p.twN_mask));
#pragma unroll
        for (int q = 0; q < TB; ++q) v[m0 + q] = v[m0 + q] * wv[q];
      }
    }
    if (active) {
      cx<T>* out = reinterpret_cast<cx<T>*>(pout) + o_lo * p.out_os + o_hi * p.out_os2 + line * p.out_ls;
      const T sc = p.scale;
      auto off = [&](int i) { return (long long)(i & p.out_seg_mask) * p.out_es + (long long)(i >> p.out_seg_shift) * p.out_seg_stride; };
      if (p.epi.on) {
        unsigned ci0, cio;
        col_coords((unsigned)line, (unsigned)p.epi.n0, ci0, cio);
        const int i0 = (int)ci0;
        const long long io = p.epi.other_from_col == 1 ? (long long)cio : (p.epi.other_from_col == 2 ? (long long)o_lo : o_hi);
        const long long base = (out - reinterpret_cast<cx<T>*>(pout));
        const bool dead0 = p.epi.dealias && ((p.epi.lo0 > 0 && i0 >= p.epi.lo0 - 1 && i0 < p.epi.hi0) || (p.epi.loo > 0 && io >= p.epi.loo - 1 && io < p.epi.hio));
        constexpr int EB = R < 4 ? R : 4;
#pragma unroll
        for (int m0 = 0; m0 < R; m0 += EB) {
          cx<T> av[EB];
          bool dead[EB];
#pragma unroll
          for (int q = 0; q < EB; ++q) {   // independent loads of the accumulated array first
            const int i = t + (m0 + q) * Tn;
            const int it = p.epi.idm * i + p.epi.ido * o_lo;
            dead[q] = dead0 || (p.epi.dealias && p.epi.lot > 0 && it >= p.epi.lot - 1 && it < p.epi.hit);
            av[q] = (p.epi.acc && !dead[q]) ? ldc(p.epi.acc + base + off(i)) : mk<T>(0, 0);
          }
#pragma unroll
          for (int q = 0; q < EB; ++q) {
            const int i = t + (m0 + q) * Tn;
            const int it = p.epi.idm * i + p.epi.ido * o_lo;
            const long long o = off(i);
            cx<T> r = mk<T>(0, 0);
            if (!dead[q]) {
              r = fuse_factor<T>(p.epi.cr, p.epi.ci, p.epi.k0, p.epi.kt, p.epi.ko, p.epi.w, i0, it, io, base + o) * (sc * v[m0 + q]);
              if (p.epi.acc) r = r + fuse_factor<T>(p.epi.ar, p.epi.ai, p.epi.a0, p.epi.at, p.epi.ao, (const T*)nullptr, i0, it, io, 0) * av[q];
            }
            stc(out + o, r);
          }
        }
      } else if (sc != T(1)) {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(out + off(t + m * Tn), sc * v[m], p.keep_out);
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) stk(out + off(t + m * Tn), v[m], p.keep_out);
      }
    }
  }
}

// L2 prefetch of the input (and the dense prologue factor) of a strided tile that this CTA will transform next
template <typename T, int R, int N>
FFB_D void fft_pow2_prefetch_cols(const Pow2Params<T>& p, const unsigned bx, const unsigned by, const void* pin, const long long nlines) {
  constexpr int Tn = N / R;
  const int W = p.W;
  const int w = threadIdx.x % W, t = threadIdx.x / W;
  const long long line = (long long)bx * W + w;
  if (line >= nlines) return;
  const int o_lo = (int)(by % (unsigned)p.outer_mod);
  const long long o_hi = by / (unsigned)p.outer_mod;
  const long long base = o_lo * p.in_os + o_hi * p.in_os2 + line * p.in_ls;
  const cx<T>* in = reinterpret_cast<const cx<T>*>(pin) + base;
  // a warp's lanes cover adjacent columns of one row: one prefetch per 128-byte line is enough
  const bool lead = ((reinterpret_cast<uintptr_t>(in) & 127) < sizeof(cx<T>)) || w == 0;
  if (!lead) return;
#pragma unroll
  for (int m = 0; m < R; ++m) prefetch_l2(in + (long long)(t + m * Tn) * p.in_es);
  if (p.pro.on && p.pro.w) {
    const T* wp = p.pro.w + base;
#pragma unroll
    for (int m = 0; m < R; ++m) prefetch_l2(wp + (long long)(t + m * Tn) * p.in_es);
  }
}

template <typename T, int DIR, int MODE, int MAXT, int MINB, int R, int... Rs>
__global__ void __launch_bounds__(MAXT, MINB) fft_pow2_kernel(const __grid_constant__ Pow2Params<T> p) {
  const unsigned bx = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned by = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  fft_pow2_tile<T, DIR, MODE, false, R, Rs...>(p, bx, by, gridDim.x, gridDim.y, p.in, p.out, p.nlines);
}

// shared-memory bytes needed by a launch with W lines per CTA (the r2c split step stages full complex values)
template <typename T> constexpr size_t pow2_smem_bytes(int N, int W, int mode) {
#ifdef FFB_ROWS_TWO_PHASE
  const bool cols = (mode == C2C_COLS || mode == C2C_COLS_TW || mode == C2C_COLS_LEAN);
  return (size_t)xpad_len(N) * W * 8 * (((mode == R2C_ROWS || cols) && sizeof(T) == 8) ? 2 : 1);
#else
  (void)mode;
  return (size_t)xpad_len(N) * W * sizeof(cx<T>);   // one complex word per point (+ padding)
#endif
}

}  // namespace ffb
