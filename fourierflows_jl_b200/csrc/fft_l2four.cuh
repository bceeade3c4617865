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
  template <typename T, int DIR, int MODE, bool IN_CG>
  static FFB_D void tile(const Pow2Params<T>& p, unsigned bx, unsigned by, const void* pin, void* pout, long long nlines) {
    fft_pow2_tile<T, DIR, MODE, IN_CG, R_, Rs...>(p, bx, by, 1u, 1u, pin, pout, nlines);
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
  int Wc;                  // columns per chunk
  int tca, tcb;            // column tiles per chunk of A / B
  int tA, tB;              // tiles per chunk: tca * N2, tcb * N1
  long long inner;         // columns per outer slice
  unsigned* ctr;           // [0] ticket, [1] exit count, [2 .. 2+C) doneA, [2+C .. 2+2C) doneB   (all zero between launches)
};

FFB_D unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct L2Tile { int isB, c, idx; };

template <typename T> FFB_D L2Tile l2four_decode(const L2FourParams<T>& p, long long tk) {
  L2Tile t;
  const long long headA = (long long)p.D * p.tA;
  const int tAB = p.tA + p.tB;
  if (tk < headA) { t.isB = 0; t.c = (int)(tk / p.tA); t.idx = (int)(tk % p.tA); return t; }
  tk -= headA;
  const long long mid = (long long)(p.C - p.D) * tAB;
  if (tk < mid) {
    const int s = p.D + (int)(tk / tAB), r = (int)(tk % tAB);
    if (r < p.tA) { t.isB = 0; t.c = s; t.idx = r; } else { t.isB = 1; t.c = s - p.D; t.idx = r - p.tA; }
    return t;
  }
  tk -= mid;
  t.isB = 1; t.c = p.C - p.D + (int)(tk / p.tB); t.idx = (int)(tk % p.tB);
  return t;
}

template <typename T, int DIR, class PA, class PB, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fft_l2four_kernel(const __grid_constant__ L2FourParams<T> p) {
  constexpr int N1 = PA::N, N2 = PB::N;
  // loop state lives in shared memory: the transform needs every register (Float64: 64 data registers of 128)
  __shared__ unsigned s_ticket;
  __shared__ int s_known[2];   // thread 0: chunks <= known are complete (a CTA meets its waits in increasing chunk order)
  const int tid = threadIdx.x;
  if (tid == 0) { s_ticket = atomicAdd(p.ctr, 1u); s_known[0] = -1; s_known[1] = -1; }
  __syncthreads();
  unsigned tk = s_ticket;
  while (tk < (unsigned)p.C * (unsigned)(p.tA + p.tB)) {
    const L2Tile t = l2four_decode(p, (long long)tk);
    const int cc = t.c % p.ncc;
    const unsigned oc = (unsigned)(t.c / p.ncc);
    const long long col0 = (long long)cc * p.Wc;
    const long long nl = min(col0 + (long long)p.Wc, p.inner);
    // scratch addressing uses the tile function's global line index: fold the chunk origin into the base pointer
    cx<T>* sbase = p.ring + (long long)(t.c % p.nslots) * p.slot_elems - col0;
    unsigned* done;
    if (!t.isB) {
      const int need = t.c - p.nslots;   // slot reuse: B(need) must have read the slot
      if (tid == 0 && need > s_known[1]) {
        const unsigned* f = p.ctr + 2 + p.C + need;
        while (ld_acquire_gpu(f) < (unsigned)p.tB) __nanosleep(100);
        s_known[1] = need;
      }
      __syncthreads();   // also: the previous tile's exchange reads have finished
      const unsigned j = (unsigned)(t.idx % p.tca), n2 = (unsigned)(t.idx / p.tca);
      PA::template tile<T, DIR, C2C_COLS_TW, false>(p.a, (unsigned)(col0 / p.a.W) + j, n2 + (unsigned)N2 * oc, p.a.in, sbase, nl);
      done = p.ctr + 2 + t.c;
    } else {
      if (tid == 0 && t.c > s_known[0]) {
        const unsigned* f = p.ctr + 2 + t.c;
        while (ld_acquire_gpu(f) < (unsigned)p.tA) __nanosleep(100);
        s_known[0] = t.c;
      }
      __syncthreads();
      const unsigned j = (unsigned)(t.idx % p.tcb), k1 = (unsigned)(t.idx / p.tcb);
      PB::template tile<T, DIR, C2C_COLS, true>(p.b, (unsigned)(col0 / p.b.W) + j, k1 + (unsigned)N1 * oc, sbase, p.b.out, nl);
      done = p.ctr + 2 + p.C + t.c;
    }
    __syncthreads();   // A: every thread's scratch stores are issued; B: every thread's scratch loads have been consumed
    if (tid == 0) {
      const unsigned next = atomicAdd(p.ctr, 1u);   // its round trip overlaps the fence
      __threadfence();                               // publish (release at gpu scope) ...
      atomicAdd(done, 1u);                           // ... then count this tile as complete
      s_ticket = next;
    }
    __syncthreads();
    tk = s_ticket;
  }
  // the last CTA to leave zeroes the counters for the next launch (every other CTA has stopped using them)
  __shared__ unsigned s_last;
  if (tid == 0) { __threadfence(); s_last = atomicAdd(p.ctr + 1, 1u); }
  __syncthreads();
  if (s_last == gridDim.x - 1) {
    for (int i = tid; i < 2 + 2 * p.C; i += THREADS) p.ctr[i] = 0u;
  }
}

}  // namespace ffb
