// Fused four-step transform of long strided lines with the intermediate kept in L2 (sm_100a, 126 MB L2).
//
// A strided line of N = N1*N2 points is too long for one CTA to own a wide tile of (fft_plan.cu:fourstep_min), so it is
// transformed as two short sub-passes: A = N1-point transforms over n1 for fixed n2, times exp(-/+2 pi i n2 k1 / N);
// B = N2-point transforms over n2 for fixed k1, landing on k = k1 + N1 k2.  Run as two kernels the intermediate makes a full
// HBM round trip (a 2-D transform costs P + 5S instead of the P + 3S of SURVEY 8d).  Here ONE persistent kernel runs both
// sub-passes chunk by chunk: the array is cut into chunks of Wc adjacent columns (x one outer slice); A(c) writes its output
// into a small ring of scratch slots, B(c) reads it back while it is still resident in L2.  DRAM sees the input once and the
// output once; the ring (a few tens of MB) is overwritten in place, so its dirty lines are written back at most once per
// transform instead of once per chunk.  Because A(c) has completely finished before B(c) starts and chunks own disjoint
// columns, the transform may run in place.
//
// Scheduling: CTAs draw tile tickets from a global counter.  Tickets are ordered A(0) .. A(D-1), [A(s) B(s-D)] for s = D .. C-1,
// B(C-D) .. B(C-1), so A runs D chunks ahead of B.  B(c) waits for doneA[c] == tA, A(c) waits for doneB[c-S] == tB (slot reuse,
// S = D + 2 slots).  Every wait is on tiles with SMALLER tickets, all of which are held by resident CTAs that never wait on a
// larger ticket, so the schedule cannot deadlock whatever the hardware's CTA placement; with D large enough the waits are
// already satisfied when they are reached.
#pragma once
#include "fft_pow2.cuh"

namespace ffb {

template <int R_, int... Rs> struct RadixPlan {
  static constexpr int R = R_;
  static constexpr int N = radix_product<Rs...>::value;
  template <typename T, int DIR, int MODE, bool IN_CG, class Hook>
  static FFB_D void tile(const Pow2Params<T>& p, unsigned bx, unsigned by, const void* pin, void* pout, long long nlines, Hook hook) {
    fft_pow2_tile<T, DIR, MODE, IN_CG, R_, Rs...>(p, bx, by, 1u, 1u, pin, pout, nlines, hook);
  }
  template <typename T>
  static FFB_D void prefetch(const Pow2Params<T>& p, unsigned bx, unsigned by, const void* pin, long long nlines) {
    fft_pow2_prefetch_cols<T, R_, N>(p, bx, by, pin, nlines);
  }
};

template <typename T>
struct L2FourParams {
  Pow2Params<T> a, b;      // sub-pass A (C2C_COLS_TW; a.out_* = scratch strides) and B (C2C_COLS; b.in_* = scratch strides)
  cx<T>* ring;             // nslots * slot_elems complex elements
  long long slot_elems;
  int nslots;
  int D;                   // chunks of lookahead of A over B (<= C)
  int C;                   // chunks = ncc * nouter
  int ncc;                 // column chunks per outer slice
  int Wc;                  // columns per chunk (a power of two)
  // tiles per chunk are the same for both sub-passes and a power of two: tA = (Wc/WA)*N2 = (Wc/WB)*N1 = Wc*N/(16*threads)
  int lgT;                 // log2(tiles per chunk and sub-pass)
  int lg_tca, lg_tcb;      // log2(column tiles per chunk) of A / B
  int lgWA, lgWB;          // log2(columns per tile) of A / B
  long long inner;         // columns per outer slice
  unsigned* ctr;           // [0] ticket, [1] exit count, doneA at ctr + 4, doneB at ctr + 4 + Cpad (zero between launches)
  int Cpad;                // C rounded up to a multiple of 4 (completion counters are polled four at a time)
  int pf;                  // != 0: prefetch the next A tile's input into L2 while this tile is transformed
  int acq;                 // != 0: close every wait with an acquire fence (see l2four_wait)
  unsigned long long* dbg; // measurement aid (FFB_L2_DEBUG): [0] cycles waiting for slots, [1] waiting for A, [2] publishing, [3] in the loop, [4] tiles, [5] waits
};

struct L2Tile { int isB, c, idx; };

// ticket -> (sub-pass, chunk, tile in chunk); shifts only: the scheduler runs once per tile on every thread
template <typename T> FFB_D L2Tile l2four_decode(const L2FourParams<T>& p, unsigned tk) {
  L2Tile t;
  const unsigned headA = (unsigned)p.D << p.lgT, m = (1u << p.lgT) - 1u;
  if (tk < headA) { t.isB = 0; t.c = (int)(tk >> p.lgT); t.idx = (int)(tk & m); return t; }
  tk -= headA;
  const unsigned mid = (unsigned)(p.C - p.D) << (p.lgT + 1);
  if (tk < mid) {
    const int s = p.D + (int)(tk >> (p.lgT + 1));
    t.isB = (int)((tk >> p.lgT) & 1u);
    t.c = t.isB ? s - p.D : s;
    t.idx = (int)(tk & m);
    return t;
  }
  tk -= mid;
  t.isB = 1; t.c = p.C - p.D + (int)(tk >> p.lgT); t.idx = (int)(tk & m);
  return t;
}

// Publishes a finished tile: release at gpu scope + count (red.release.gpu = MEMBAR.ALL.GPU + REDG; unlike __threadfence() it
// does not invalidate the SM's L1, which keeps the twiddle tables).  Called by thread 0 from the NEXT tile's after-load hook, so
// the fence (which waits for this thread's outstanding accesses) overlaps the load latency the warp has to sit out anyway.
struct L2Publish {
  unsigned long long* pending;   // shared-memory slot holding the completion counter's address (0: nothing to publish)
  unsigned long long* acc;       // shared-memory cycle accumulator (measurement aid) or nullptr
  FFB_D void operator()() const {
    if (threadIdx.x == 0) {
      unsigned* done = reinterpret_cast<unsigned*>(*pending);
      if (done != nullptr) {
        const long long t0 = acc ? clock64() : 0;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(done), "r"(1u) : "memory");
        if (acc) *acc += (unsigned long long)(clock64() - t0);
      }
      *pending = 0ull;
    }
  }
};

// Waits (thread 0) until done[c] >= want and returns the last chunk c' >= c of c's aligned group of four whose counters are
// all complete as well (one 16-byte load shows four counters; chunks finish roughly in order, so later tiles of this CTA can
// usually skip their poll).  Polls with relaxed loads (no L1 invalidation per poll); `acq` != 0 closes with an acquire fence as
// the PTX memory model asks for.  The data the counters guard is only ever read with ld.global.cg (L2), and L2 is the point of
// coherence, so acq = 0 is safe on this hardware (in-order issue: the loads are issued after the barrier that follows the
// poll); it is kept as a measurement switch, the default is the formally correct form.
FFB_D int l2four_wait(const unsigned* done, int c, unsigned want, int acq) {
  const unsigned* g = done + (c & ~3);
  const int k = c & 3;
  unsigned v[4];
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(g) : "memory");
    const unsigned mine = k == 0 ? v[0] : k == 1 ? v[1] : k == 2 ? v[2] : v[3];
    if (mine >= want) break;
    __nanosleep(64);
  }
  if (acq) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  int last = c;
  if (k < 1 && v[1] >= want) last = (c & ~3) + 1; else if (k < 1) return last;
  if (k < 2 && v[2] >= want) last = (c & ~3) + 2; else if (k < 2) return last;
  if (k < 3 && v[3] >= want) last = (c & ~3) + 3;
  return last;
}

template <typename T, int DIR, class PA, class PB, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fft_l2four_kernel(const __grid_constant__ L2FourParams<T> p) {
  constexpr int N1 = PA::N, N2 = PB::N;
  // loop state lives in shared memory: the transform needs every register (Float64: 64 data registers of 128)
  __shared__ unsigned s_tk[4];   // ring of drawn tickets: tile i uses s_tk[i & 3]; tickets are drawn two tiles ahead
  __shared__ int s_known[2];     // [0]: A chunks <= this are complete, [1]: B chunks (a CTA meets its waits in increasing chunk order)
  __shared__ unsigned long long s_pending;   // completion counter of the tile whose stores are issued but not yet published
  __shared__ unsigned long long s_dbg[6];
  const int tid = threadIdx.x;
  const bool dbg = p.dbg != nullptr;
  if (dbg && tid < 6) s_dbg[tid] = 0ull;
  const long long t_begin = dbg ? clock64() : 0;
  const unsigned total = (unsigned)p.C << (p.lgT + 1);
  const unsigned want = 1u << p.lgT;
  unsigned* const doneA = p.ctr + 4;
  unsigned* const doneB = p.ctr + 4 + p.Cpad;
  if (tid == 0) {
    s_tk[0] = atomicAdd(p.ctr, 1u); s_tk[1] = atomicAdd(p.ctr, 1u);
    s_known[0] = -1; s_known[1] = -1; s_pending = 0ull;
  }
  __syncthreads();
  for (unsigned it = 0;; ++it) {
    const unsigned tk = s_tk[it & 3];
    if (tk >= total) break;
    unsigned drawn = 0;
    if (tid == 0) drawn = atomicAdd(p.ctr, 1u);   // ticket of tile it + 2: consumed (stored) only after this tile's work
    const L2Tile t = l2four_decode(p, tk);
    const unsigned oc = p.ncc == p.C ? 0u : (unsigned)t.c / (unsigned)p.ncc;
    const unsigned cc = (unsigned)t.c - oc * (unsigned)p.ncc;
    const long long col0 = (long long)cc * p.Wc;
    const long long nl = min(col0 + (long long)p.Wc, p.inner);
    // scratch addressing uses the tile function's global line index: fold the chunk origin into the base pointer
    cx<T>* sbase = p.ring + (long long)((unsigned)t.c % (unsigned)p.nslots) * p.slot_elems - col0;
    // the tile after this one: pull its input towards L2 now (A tiles read HBM; B tiles read the ring, already in L2)
    if (p.pf) {
      const unsigned tk1 = s_tk[(it + 1) & 3];
      if (tk1 < total) {
        const L2Tile u = l2four_decode(p, tk1);
        if (!u.isB) {
          const unsigned uo = p.ncc == p.C ? 0u : (unsigned)u.c / (unsigned)p.ncc;
          const long long c0 = (long long)((unsigned)u.c - uo * (unsigned)p.ncc) * p.Wc;
          PA::template prefetch<T>(p.a, (unsigned)(c0 >> p.lgWA) + ((unsigned)u.idx & ((1u << p.lg_tca) - 1u)), ((unsigned)u.idx >> p.lg_tca) + (unsigned)N2 * uo,
                                   p.a.in, min(c0 + (long long)p.Wc, p.inner));
        }
      }
    }
    const L2Publish publish{&s_pending, dbg ? &s_dbg[2] : nullptr};
    if (!t.isB) {
      const int need = t.c - p.nslots;   // slot reuse: B(need) must have read the slot
      if (need > s_known[1]) {
        publish();   // a blocked CTA must have published everything it finished (its own tile may be what others wait for)
        int upto = need;
        if (tid == 0) {
          const long long t0 = dbg ? clock64() : 0;
          upto = l2four_wait(doneB, need, want, p.acq);
          if (dbg) { s_dbg[0] += (unsigned long long)(clock64() - t0); s_dbg[5] += 1; }
        }
        __syncthreads();
        if (tid == 0) s_known[1] = upto;
      }
      const unsigned j = (unsigned)t.idx & ((1u << p.lg_tca) - 1u), n2 = (unsigned)t.idx >> p.lg_tca;
      PA::template tile<T, DIR, C2C_COLS_TW, false>(p.a, (unsigned)(col0 >> p.lgWA) + j, n2 + (unsigned)N2 * oc, p.a.in, sbase, nl, publish);
      if (tid == 0) s_pending = (unsigned long long)(doneA + t.c);
    } else {
      if (t.c > s_known[0]) {
        publish();
        int upto = t.c;
        if (tid == 0) {
          const long long t0 = dbg ? clock64() : 0;
          upto = l2four_wait(doneA, t.c, want, p.acq);
          if (dbg) { s_dbg[1] += (unsigned long long)(clock64() - t0); s_dbg[5] += 1; }
        }
        __syncthreads();
        if (tid == 0) s_known[0] = upto;
      }
      const unsigned j = (unsigned)t.idx & ((1u << p.lg_tcb) - 1u), k1 = (unsigned)t.idx >> p.lg_tcb;
      PB::template tile<T, DIR, C2C_COLS, true>(p.b, (unsigned)(col0 >> p.lgWB) + j, k1 + (unsigned)N1 * oc, sbase, p.b.out, nl, publish);
      if (tid == 0) s_pending = (unsigned long long)(doneB + t.c);
    }
    if (tid == 0) { s_tk[(it + 2) & 3] = drawn; if (dbg) s_dbg[4] += 1; }
    __syncthreads();   // A: every thread's scratch stores are issued; B: every thread's scratch loads have been consumed; s_tk / s_known visible
  }
  L2Publish{&s_pending, nullptr}();
  if (dbg && tid == 0) {
    s_dbg[3] = (unsigned long long)(clock64() - t_begin);
    for (int i = 0; i < 6; ++i) atomicAdd(p.dbg + i, s_dbg[i]);
  }
  // the last CTA to leave zeroes the counters for the next launch (every other CTA has stopped using them)
  __shared__ unsigned s_last;
  if (tid == 0) { __threadfence(); s_last = atomicAdd(p.ctr + 1, 1u); }
  __syncthreads();
  if (s_last == gridDim.x - 1) {
    for (int i = tid; i < 4 + 2 * p.Cpad; i += THREADS) p.ctr[i] = 0u;
  }
}

}  // namespace ffb
