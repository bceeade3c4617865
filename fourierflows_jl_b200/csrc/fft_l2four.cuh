// Fused four-step transform of long strided lines with the intermediate kept in L2 (sm_100a, 126 MB L2).
//
// A strided line of N = N1*N2 points is too long for one CTA to own a wide tile of (fft_plan.cu:fourstep_min), so it is
// transformed as two short sub-passes (tiles: fft_fs.cuh): A = N1-point transforms over n1 for fixed n2, times exp(-/+2 pi i n2 k1 / N);
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
#include "fft_fs.cuh"

namespace ffb {

template <typename T>
struct L2FourParams {
  FsParams<T> a, b;        // sub-pass A (a.out_es = scratch stride N2*Wc) and B (b.in_es = scratch stride Wc)
  const cx<T>* in;         // true input array (read by A)
  cx<T>* out;              // true output array (written by B); may equal `in`
  int N;                   // N1 * N2
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

// Scheduler work of thread 0 inside a tile, run right before the tile's stores (fs_tile: sched.before_store()), when none of the
// thread's memory accesses is outstanding any more:
//   * publish the PREVIOUS tile: release at gpu scope + count (red.release.gpu = MEMBAR.ALL.GPU + REDG; unlike __threadfence()
//     it does not invalidate the SM's L1).  The previous tile's stores were issued a whole tile ago, so the fence returns at once;
//   * consume the early look at the NEXT tile's dependency (a relaxed load issued at the start of this tile): when the chunk is
//     already complete the next tile starts without a poll and without the barrier that would follow it.
struct L2Sched {
  unsigned long long* pending;   // shared-memory slot holding the completion counter's address (0: nothing to publish)
  unsigned long long* acc;       // shared-memory cycle accumulator (measurement aid) or nullptr
  int* known;                    // &s_known[k] of the next tile's dependency kind, or nullptr
  int next_c;                    // chunk the next tile depends on
  unsigned early, want;          // its completion counter as seen at the start of this tile / the complete value
  int acq;
  FFB_D void publish() const {
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
  FFB_D void before_store() const {
    publish();
    if (threadIdx.x == 0 && known != nullptr && early >= want && next_c > *known) {
      if (acq) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      *known = next_c;
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

template <typename T, int DIR, class PA, class PB, bool HOOK_A, bool HOOK_B, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fft_l2four_kernel(const __grid_constant__ L2FourParams<T> p) {
  constexpr int N1 = PA::N, N2 = PB::N;
  // loop state lives in shared memory: the transform needs the registers
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
  // tile geometry of a ticket: chunk -> (column chunk cc, outer slice oc); tile index -> (column tile j, o_lo)
  auto geometry = [&](const L2Tile& t, FsTile& tl, long long& col0) {
    const unsigned oc = p.ncc == p.C ? 0u : (unsigned)t.c / (unsigned)p.ncc;
    const unsigned cc = (unsigned)t.c - oc * (unsigned)p.ncc;
    col0 = (long long)cc * p.Wc;
    const int lg = t.isB ? p.lg_tcb : p.lg_tca, lgW = t.isB ? p.lgWB : p.lgWA;
    const unsigned j = (unsigned)t.idx & ((1u << lg) - 1u);
    tl.o_lo = (int)((unsigned)t.idx >> lg);
    tl.o_hi = oc;
    const long long loc = (long long)j << lgW;            // first column of the tile inside the chunk
    tl.line0 = col0 + loc;
    tl.ncols = (int)max(0ll, min((long long)(1 << lgW), p.inner - tl.line0));
    const long long slot = (long long)((unsigned)t.c % (unsigned)p.nslots) * p.slot_elems;
    const long long arr = (long long)tl.o_lo * p.inner + (long long)oc * p.inner * p.N + tl.line0;   // o_lo = n2 (A, input) / k1 (B, output)
    if (!t.isB) { tl.in_off = arr; tl.out_off = slot + (long long)tl.o_lo * p.Wc + loc; }
    else { tl.in_off = slot + (long long)tl.o_lo * N2 * p.Wc + loc; tl.out_off = arr; }
    tl.hook_off = arr;
  };
  for (unsigned it = 0;; ++it) {
    const unsigned tk = s_tk[it & 3];
    if (tk >= total) break;
    unsigned drawn = 0;
    if (tid == 0) drawn = atomicAdd(p.ctr, 1u);   // ticket of tile it + 2: consumed (stored) only after this tile's work
    const L2Tile t = l2four_decode(p, tk);
    FsTile tl;
    long long col0;
    geometry(t, tl, col0);
    // the tile after this one: pull its input towards L2 now (A tiles read HBM; B tiles read the ring, already in L2)
    if (p.pf) {
      const unsigned tk1 = s_tk[(it + 1) & 3];
      if (tk1 < total) {
        const L2Tile u = l2four_decode(p, tk1);
        if (!u.isB) {
          FsTile ul;
          long long c0;
          geometry(u, ul, c0);
          PA::template prefetch<T>(p.a, p.in, ul, HOOK_A);
        }
      }
    }
    // early look at the next tile's dependency: one relaxed load now, consumed before this tile's stores
    L2Sched sched{&s_pending, dbg ? &s_dbg[2] : nullptr, nullptr, 0, 0u, want, p.acq};
    {
      const unsigned tk1 = s_tk[(it + 1) & 3];
      if (tk1 < total) {
        const L2Tile u = l2four_decode(p, tk1);
        const int dep = u.isB ? u.c : u.c - p.nslots;
        if (dep >= 0) {
          sched.known = &s_known[u.isB ? 0 : 1];
          sched.next_c = dep;
          if (tid == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sched.early) : "l"((u.isB ? doneA : doneB) + dep) : "memory");
        }
      }
    }
    if (!t.isB) {
      const int need = t.c - p.nslots;   // slot reuse: B(need) must have read the slot
      if (need > s_known[1]) {
        sched.publish();   // a blocked CTA must have published everything it finished (its own tile may be what others wait for)
        int upto = need;
        if (tid == 0) {
          const long long t0 = dbg ? clock64() : 0;
          upto = l2four_wait(doneB, need, want, p.acq);
          if (dbg) { s_dbg[0] += (unsigned long long)(clock64() - t0); s_dbg[5] += 1; }
        }
        __syncthreads();
        if (tid == 0) s_known[1] = upto;
      }
      PA::template tile<T, DIR, true, HOOK_A, false, true>(p.a, p.in, p.ring, tl, sched);
      if (tid == 0) s_pending = (unsigned long long)(doneA + t.c);
    } else {
      if (t.c > s_known[0]) {
        sched.publish();
        int upto = t.c;
        if (tid == 0) {
          const long long t0 = dbg ? clock64() : 0;
          upto = l2four_wait(doneA, t.c, want, p.acq);
          if (dbg) { s_dbg[1] += (unsigned long long)(clock64() - t0); s_dbg[5] += 1; }
        }
        __syncthreads();
        if (tid == 0) s_known[0] = upto;
      }
      PB::template tile<T, DIR, false, HOOK_B, true, false>(p.b, p.ring, p.out, tl, sched);
      if (tid == 0) s_pending = (unsigned long long)(doneB + t.c);
    }
    if (tid == 0) { s_tk[(it + 2) & 3] = drawn; if (dbg) s_dbg[4] += 1; }
    __syncthreads();   // A: every thread's scratch stores are issued; B: every thread's scratch loads have been consumed; s_tk / s_known visible
  }
  L2Sched{&s_pending, nullptr, nullptr, 0, 0u, 0u, 0}.publish();
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
