// FFT plans: the B2 seam of SURVEY 8b.  Replaces `plan_flows_fft` / `plan_flows_rfft` (src/domains.jl:2-5) and the
// `mul!` / `ldiv!` executions on `grid.rfftplan` / `grid.fftplan` (src/diffusion.jl:137,139,154,155,171).
// A d-dimensional transform is one 1-D pass per dimension (P + (2d-1) S bytes of HBM traffic for r2c / c2r).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <mutex>
#include <vector>
#include "fft_generic.cuh"
#include "fft_pow2.cuh"
#include "fft_pow2_dispatch.h"
#include "dist.h"
#include "fft_l2four.cuh"
#include "fft_l2four_dispatch.h"

namespace ffb {

template <typename T>
static int upload(std::vector<cx<T>>& host, cx<T>** dev) {
  void* p = nullptr;
  int rc = ffb_malloc(&p, host.size() * sizeof(cx<T>));
  if (rc) return rc;
  cudaError_t e = cudaMemcpy(p, host.data(), host.size() * sizeof(cx<T>), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(p); return set_error(FFB_ECUDA, "twiddle upload: %s", cudaGetErrorString(e)); }
  *dev = reinterpret_cast<cx<T>*>(p);
  return FFB_OK;
}

// exp(-2*pi*i*q/M) evaluated in long double with octant reduction, rounded once to T
template <typename T>
static cx<T> unit_root(long long q, long long M) {
  q %= M;
  const long double pi = 3.14159265358979323846264338327950288L;
  // reduce to first octant for accuracy
  long long q8 = q * 8;
  int oct = (int)(q8 / M);
  long double c, s;
  long long rem = q8 - (long long)oct * M;  // angle = 2*pi*(oct*M + rem)/(8M)
  long double th = 2.0L * pi * (long double)rem / (8.0L * (long double)M);  // in [0, pi/4)
  long double c0 = cosl(th), s0 = sinl(th);
  const long double h = 0.70710678118654752440084436210485L;
  switch (oct & 7) {
    case 0: c = c0; s = s0; break;
    case 1: c = h * (c0 - s0); s = h * (c0 + s0); break;
    case 2: c = -s0; s = c0; break;
    case 3: c = -h * (c0 + s0); s = h * (c0 - s0); break;
    case 4: c = -c0; s = -s0; break;
    case 5: c = -h * (c0 - s0); s = -h * (c0 + s0); break;
    case 6: c = s0; s = -c0; break;
    default: c = h * (c0 + s0); s = -h * (c0 - s0); break;
  }
  return mk<T>((T)c, (T)(-s));
}

template <typename T>
struct DimTables {
  int N = 0;             // complex transform length
  bool pow2 = false;     // register-resident kernel available
  cx<T>* tw = nullptr;   // pow2 per-pass twiddles
  cx<T>* tw8 = nullptr;  // Float64 row passes with 8 points per thread: twiddles of the radix-8 plan (dim 0 only)
  cx<T>* wN = nullptr;   // generic: exp(-2 pi i q / N), q < N
  cx<T>* twr = nullptr;  // split step: exp(-i pi k / N), k <= N (dim 0 of R2C plans only)
  // four-step split of a long strided line: N = N1*N2, two short sub-passes with wide rows
  bool four = false;
  int N1 = 0, N2 = 0;
  cx<T>* tw1 = nullptr;  // base twiddles of the length-N1 sub-transform
  cx<T>* tw2 = nullptr;  // base twiddles of the length-N2 sub-transform
  cx<T>* twN = nullptr;  // exp(-2 pi i q / N), q < N: inter-pass twiddles
};

}  // namespace ffb

using namespace ffb;

struct ffb_plan {
  int ndim, dtype, kind, nbatch, flags;
  long long n[3];     // physical sizes
  long long nc[3];    // complex array extents (nc[0] = n0/2+1 for R2C)
  void* tables[3];    // DimTables<T>*
  void* ws[3];        // scratch, allocated on first use
  void* wsm[4];       // sub-pass A outputs of a multi-variant inverse transform (ffb_fft_inverse_multi), allocated on first use
  size_t ws_bytes;    // size of each scratch array
  std::string desc;
  ffb_dist* dist;     // non-NULL: slab-decomposed 3-D r2c plan (physical z-slabs <-> spectral y-slabs)
  long long nyl, nzl; // local extents: spectral (nkr, nyl, nz), physical (nx, ny, nzl); 2-D: physical (nx, nyl)
  long long kb;       // 2-D slab decomposition: wavenumbers per rank block (local spectral slab: (kb + 1, ny))
  int ws0_zeroed;
  int nchunks;        // exchange chunks along the local z range (overlap of all-to-all and local passes)
  // fused pass + collective: double-buffered receive buffers and their peer mappings (CUDA IPC)
  void* recv[2];
  void* peers[2][8];
  int p2p;                 // exchange mode (FFB_EXCHANGE_*)
  size_t recv_bytes;       // size of each receive buffer
  int p2p_cur;
  // fused four-step passes (fft_l2four.cuh): L2-resident scratch ring and the ticket / completion counters
  void* ring; size_t ring_bytes;
  unsigned* ctr; size_t ctr_count; int ctr_C;
};

namespace ffb {

// base twiddles of a Stockham plan with the given radix sequence: for each pass with Ns > 1, exp(-2 pi i a/(Ns r)), a < Ns
template <typename T>
static int build_tw_radices(const int* rad, int np, cx<T>** dev) {
  std::vector<cx<T>> h;
  int Ns = 1;
  for (int i = 0; i < np; ++i) {
    const int r = rad[i];
    if (Ns > 1)
      for (int a = 0; a < Ns; ++a) h.push_back(unit_root<T>((long long)a, (long long)Ns * r));
    Ns *= r;
  }
  if (h.empty()) h.push_back(mk<T>(1, 0));
  return upload(h, dev);
}
// register-resident plan for length N (radix sequence of fft_pow2_dispatch.h)
template <typename T>
static int build_pow2_tw(int N, cx<T>** dev) {
  int rad[8];
  const int np = pow2_radices(N, rad);
  return build_tw_radices<T>(rad, np, dev);
}
// four-step sub-transform of length N (radix sequence of fft_fs.cuh: 8 points per thread in Float64, 16 in Float32)
template <typename T>
static int build_fs_tw(int N, cx<T>** dev) {
  int rad[8];
  const int np = fs_radices(N, (int)sizeof(T), rad);
  if (np == 0) return set_error(FFB_EUNSUPPORTED, "no four-step sub-transform of length %d", N);
  return build_tw_radices<T>(rad, np, dev);
}

// Float64 row passes (dim 0) of 512 .. 4096 points have a second plan with 8 points per thread (64 registers, two 512-thread
// CTAs = 32 warps per SM instead of 16: the row kernels are latency-bound).  pow2_pass finds it through the table of the
// 16-point plan it is handed.  MEASURED SLOWER (8192^2 r2c 0.655 vs 0.593 ms: four passes / three two-phase exchanges with 16 warps per
// barrier cost more than the extra warps hide): opt-in with FFB_ROWS_R8=1.
static std::mutex g_tw8_mu;
static std::map<const void*, const void*> g_tw8;
static const void* rows_tw8_for(const void* tw) {
  std::lock_guard<std::mutex> lk(g_tw8_mu);
  auto it = g_tw8.find(tw);
  return it == g_tw8.end() ? nullptr : it->second;
}

// Strided lines at least this long are transformed as a four-step pair of short sub-passes (measured: one register-resident
// pass over 8192 F64 points reaches 1.6 TB/s because a CTA can own only one 16-byte wide column; sub-passes of 64 / 128
// points own 32-64 columns and run near the HBM roofline).  FFB_FOURSTEP_MIN overrides (0 disables).
template <typename T> static int fourstep_min() {
  if (const char* e = getenv("FFB_FOURSTEP_MIN")) { const int v = atoi(e); return v > 0 ? v : (1 << 30); }
  return 4096;   // Float32: 4096^2 r2c 0.084 ms split against 0.094 ms single pass (profiles/r02_fft_sweep.log)
}

template <typename T>
static int build_tables(ffb_plan* pl) {
  for (int d = 0; d < pl->ndim; ++d) {
    auto* tb = new DimTables<T>();
    pl->tables[d] = tb;
    const bool split = (pl->kind == FFB_R2C && d == 0);
    const int N = (int)(split ? pl->n[0] / 2 : pl->n[d]);
    tb->N = N;
    int rad[8];
    const int np = pow2_radices(N, rad);
    const bool allow = !(pl->flags & FFB_PLAN_FORCE_GENERIC) && is_pow2((uint64_t)N) && N >= 2;
    tb->pow2 = allow && N <= pow2_max_n(sizeof(T)) && np > 0;
    if (allow && d > 0 && N >= fourstep_min<T>() && N <= 65536) {
      const int l2 = ilog2((uint64_t)N);
      tb->four = true;
      tb->N1 = 1 << (l2 / 2);
      tb->N2 = N / tb->N1;
      tb->pow2 = true;
      int rc = build_fs_tw<T>(tb->N1, &tb->tw1);
      if (rc) return rc;
      if ((rc = build_fs_tw<T>(tb->N2, &tb->tw2))) return rc;
      std::vector<cx<T>> h((size_t)N);
      for (int q = 0; q < N; ++q) h[q] = unit_root<T>(q, N);
      if ((rc = upload(h, &tb->twN))) return rc;
      if (N <= pow2_max_n(sizeof(T)) && np > 0 && (rc = build_pow2_tw<T>(N, &tb->tw))) return rc;
    } else if (tb->pow2) {
      int rc = build_pow2_tw<T>(N, &tb->tw);
      if (rc) return rc;
    } else {
      std::vector<cx<T>> h((size_t)std::max(N, 1));
      for (int q = 0; q < N; ++q) h[q] = unit_root<T>(q, N);
      if (N == 0) h[0] = mk<T>(1, 0);
      int rc = upload(h, &tb->wN);
      if (rc) return rc;
    }
    if (d == 0 && sizeof(T) == 8 && tb->pow2 && tb->tw) {
      int r8[8];
      const int n8 = pow2_radices_r8(N, r8);
      if (n8 > 0) {
        int rc = build_tw_radices<T>(r8, n8, &tb->tw8);
        if (rc) return rc;
        std::lock_guard<std::mutex> lk(g_tw8_mu);
        g_tw8[tb->tw] = tb->tw8;
      }
    }
    if (split) {
      std::vector<cx<T>> h((size_t)N + 1);
      for (int k = 0; k <= N; ++k) h[k] = unit_root<T>(k, 2ll * N);
      int rc = upload(h, &tb->twr);
      if (rc) return rc;
    }
    char buf[96];
    if (tb->four) snprintf(buf, sizeof(buf), "dim%d:N=%d:pow2-four-step(%dx%d) ", d, N, tb->N1, tb->N2);
    else snprintf(buf, sizeof(buf), "dim%d:N=%d:%s ", d, N, tb->pow2 ? "pow2-register-stockham" : "generic-mixed-radix");
    pl->desc += buf;
  }
  // a plan that mixes an arbitrary-size dimension (generic path: needs wN of EVERY dimension it touches) with a four-step-only
  // power of two (no single-pass table) cannot execute: say so now, not at the first transform in the middle of a step
  bool any_generic = false, any_four_only = false;
  for (int d = 0; d < pl->ndim; ++d) {
    auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[d]);
    any_generic = any_generic || !tb->pow2;
    any_four_only = any_four_only || (tb->four && !tb->tw);
  }
  if (any_generic && any_four_only)
    return set_error(FFB_EUNSUPPORTED, "plan mixes an arbitrary-size dimension with a power-of-two dimension longer than %d (four-step only)", pow2_max_n(sizeof(T)));
  return FFB_OK;
}

template <typename T>
static void free_tables(ffb_plan* pl) {
  for (int d = 0; d < pl->ndim; ++d) {
    auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[d]);
    if (!tb) continue;
    if (tb->tw8) { std::lock_guard<std::mutex> lk(g_tw8_mu); g_tw8.erase(tb->tw); }
    cudaFree(tb->tw8);
    cudaFree(tb->tw); cudaFree(tb->wN); cudaFree(tb->twr); cudaFree(tb->tw1); cudaFree(tb->tw2); cudaFree(tb->twN);
    delete tb;
  }
}

static int ensure_ws(ffb_plan* pl, int count) {
  for (int i = 0; i < count; ++i)
    if (!pl->ws[i]) {
      int rc = ffb_malloc(&pl->ws[i], pl->ws_bytes);
      if (rc) return rc;
    }
  return FFB_OK;
}

template <typename T>
static int call_pow2(int N, int mode, int dir, const Pow2Params<T>& p, int gx, int gy, int threads, size_t smem, cudaStream_t st, bool r8 = false) {
  int rc;
  if constexpr (sizeof(T) == 8) {
    if (r8) {
      if ((rc = pow2_launch_double_r8(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
      return set_error(FFB_EUNSUPPORTED, "no 8-point-per-thread row kernel for N=%d", N);
    }
    if ((rc = pow2_launch_double_g0(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_double_g1(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_double_g2(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_double_g3(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
  } else {
    if ((rc = pow2_launch_float_g0(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_float_g1(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_float_g2(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
    if ((rc = pow2_launch_float_g3(N, mode, dir, &p, gx, gy, threads, smem, st)) != 1) return rc;
  }
  return set_error(FFB_EUNSUPPORTED, "no register-resident FFT kernel for N=%d", N);
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Lines-per-CTA choice (measured: tools/sweep_w.py, profiles/r01_fft_sweep_v*.log).  Throughput is set by how
// many independent CTAs an SM can overlap, so CTAs are kept small.  ROWS: one line per CTA (>= 32 threads).
// COLS: about 256 threads with W adjacent columns clamped to [64 B, 256 B] wide rows (Float64 4..16, Float32 16..32
// columns, 8 only when 16 do not fit -- N = 2048; the twiddled four-step sub-pass likes 32).
template <typename T>
static int choose_w(int N, int mode, long long nlines) {
  const int R = pow2_points_per_thread(N);
  const int Tn = N / R;
  const int maxT = pow2_max_threads(sizeof(T));
  const size_t smem_cap = (size_t)max_smem_optin() - 1024;
  const bool cols = (mode == C2C_COLS || mode == C2C_COLS_TW);
  int W = 1;
  if (cols) {
    const int lo = sizeof(T) == 8 ? 4 : 16, hi = (sizeof(T) == 8 && mode == C2C_COLS) ? 16 : 32;   // Float32: 128-byte rows when the tile fits
    int want = std::max(lo, std::min(hi, 256 / std::max(Tn, 1)));
    while (W < want && Tn * W * 2 <= maxT && pow2_smem_bytes<T>(N, W * 2, mode) <= smem_cap) W *= 2;
  } else {
    while (Tn * W < 32 && Tn * W * 2 <= maxT && pow2_smem_bytes<T>(N, W * 2, mode) <= smem_cap) W *= 2;
  }
  // tuning overrides (measurement only): FFB_W_COLS / FFB_W_ROWS force the lines-per-CTA when legal
  if (const char* e = getenv((mode == C2C_COLS || mode == C2C_COLS_TW) ? "FFB_W_COLS" : "FFB_W_ROWS")) {
    const int w = atoi(e);
    if (w >= 1 && Tn * w <= maxT && pow2_smem_bytes<T>(N, w, mode) <= smem_cap && ((mode != C2C_COLS && mode != C2C_COLS_TW) || is_pow2((uint64_t)w))) W = w;
  }
  while (W > 1 && W / 2 >= nlines) W /= 2;
  return W;
}

// One register-resident pass over `nouter` outer blocks (gridDim.y chunks of <= 65535).
// snake ordering state set by the pass scheduler (exec_pow2) for the next launch
static thread_local int g_pass_reverse = 0, g_pass_keep = 0;
static int snake_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFB_SNAKE"); v = e ? atoi(e) : 0; }
  return v;
}

struct SegStride { int seg = 0; long long stride = 0; };  // seg = 0: unsegmented
// half spectrum of the next row pass cut into per-rank blocks (2-D slab decomposition; set around the call by exec_dist2d)
struct RowSeg { int seg = 0; long long stride = 0, nyq = 0; };
static thread_local RowSeg g_row_seg;
// dealias-aware forward transforms: dead half-spectrum range of the next r2c pass / dead columns of the next strided pass (see DeadCols)
static thread_local int g_row_dead_lo = 0, g_row_dead_hi = 0;
static thread_local DeadCols g_dead = {0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
// two-level outer index (four-step sub-passes): blockIdx.y = o_lo + mod*o_hi -> o_lo*os + o_hi*os2
struct Outer2 { int mod = 0; long long nhi = 1, in_os2 = 0, out_os2 = 0; };

template <typename T>
static int pow2_pass(int N, int mode, int dir, const void* in, void* out, long long in_es, long long in_ls, long long in_os,
                     long long out_es, long long out_ls, long long out_os, long long nlines, long long nouter, T scale,
                     const cx<T>* tw, const cx<T>* twr, cudaStream_t st, SegStride in_seg = SegStride(), SegStride out_seg = SegStride(),
                     Outer2 o2 = Outer2(), const cx<T>* twN = nullptr, int twN_mask = 0,
                     const typename Pow2Params<T>::Fuse* pro = nullptr, const typename Pow2Params<T>::Fuse* epi = nullptr, const T* rmul = nullptr,
                     int rsq = 0) {
  Pow2Params<T> p;
  p.rsq = rsq;
  p.row_seg = g_row_seg.seg; p.row_seg_mask = g_row_seg.seg ? g_row_seg.seg - 1 : 0; p.row_seg_shift = g_row_seg.seg ? ilog2((uint64_t)g_row_seg.seg) : 0;
  p.row_seg_stride = g_row_seg.stride; p.row_nyq = g_row_seg.nyq;
  p.row_dead_lo = g_row_dead_lo; p.row_dead_hi = g_row_dead_hi; p.dead = g_dead;
  // plain strided passes: lean kernel variant with host-computed element offsets (needs segment lengths that are multiples of N/R)
  long long lean_out_off[16] = {0};
  static int lean_env = -1;
  if (lean_env < 0) { const char* e = getenv("FFB_LEAN"); lean_env = e ? atoi(e) : 1; }
  if (lean_env && mode == C2C_COLS && !pro && !epi && !rmul && !o2.mod && in_ls == 1 && out_ls == 1 && !g_pass_reverse) {
    const int R0 = pow2_points_per_thread(N), Tn = N / R0;
    if ((!in_seg.seg || in_seg.seg % Tn == 0) && (!out_seg.seg || out_seg.seg % Tn == 0)) {
      mode = C2C_COLS_LEAN;
      for (int m = 0; m < 16; ++m) {
        const long long i = (long long)m * Tn;
        p.in_off[m] = in_seg.seg ? (i % in_seg.seg) * in_es + (i / in_seg.seg) * in_seg.stride : i * in_es;
        lean_out_off[m] = out_seg.seg ? (i % out_seg.seg) * out_es + (i / out_seg.seg) * out_seg.stride : i * out_es;
      }
    }
  }
  if (pro) p.pro = *pro; else p.pro.on = 0;
  if (epi) p.epi = *epi; else p.epi.on = 0;
  p.rmul = rmul;
  p.reverse = g_pass_reverse;
  p.keep_out = g_pass_keep;
  p.pf_ahead = 0;
  if (mode == C2C_ROWS || mode == R2C_ROWS || mode == C2R_ROWS) {
    // distance = one full wave of resident CTAs (tuning override: FFB_PF_AHEAD, 0 disables)
    static int pf_env = -2;
    if (pf_env == -2) { const char* e = getenv("FFB_PF_AHEAD"); pf_env = e ? atoi(e) : -1; }
    // measured (round-1 sweep): +10 % for the Float64 r2c pass of 4096-point lines at one wave ahead, neutral or negative elsewhere
    p.pf_ahead = pf_env >= 0 ? pf_env : ((mode == R2C_ROWS && N * sizeof(cx<T>) >= 32768) ? num_sms() : 0);
  }
  if (mode == C2C_COLS_LEAN) {
    // measured (tools/experiments/stream_sweep.py): +5 % when the tile fills the SM (one CTA per SM), -20 % when several CTAs share an SM
    static int pfc = -2;
    if (pfc == -2) { const char* e = getenv("FFB_PF_COLS"); pfc = e ? atoi(e) : -1; }
    p.pf_ahead = pfc >= 0 ? pfc : -1;   // resolved below once the tile shape is known
  }
  p.in_seg_mask = in_seg.seg ? in_seg.seg - 1 : 0x7fffffff; p.in_seg_shift = in_seg.seg ? ilog2((uint64_t)in_seg.seg) : 31;
  p.in_seg_stride = in_seg.stride;
  p.out_seg_mask = out_seg.seg ? out_seg.seg - 1 : 0x7fffffff; p.out_seg_shift = out_seg.seg ? ilog2((uint64_t)out_seg.seg) : 31;
  p.out_seg_stride = out_seg.stride;
  p.in_es = in_es; p.in_ls = in_ls; p.in_os = in_os;
  p.out_es = out_es; p.out_ls = out_ls; p.out_os = out_os;
  p.in_os2 = o2.in_os2; p.out_os2 = o2.out_os2;
  p.twN = twN; p.twN_mask = twN_mask;
  p.nlines = nlines;
  p.W = choose_w<T>(N, mode == C2C_COLS_LEAN ? C2C_COLS : mode, nlines);
  p.scale = scale;
  p.tw = tw;
  p.twr = twr;
  int R = pow2_points_per_thread(N);
  bool r8 = false;
  if (sizeof(T) == 8 && (mode == C2C_ROWS || mode == R2C_ROWS || mode == C2R_ROWS) && p.W == 1 && env_int("FFB_ROWS_R8", 0)) {
    if (const void* t8 = rows_tw8_for(tw)) { p.tw = reinterpret_cast<const cx<T>*>(t8); R = 8; r8 = true; }
  }
  const int threads = (N / R) * p.W;
  if (mode == C2C_COLS_LEAN && p.pf_ahead < 0) p.pf_ahead = threads >= pow2_max_threads(sizeof(T)) ? num_sms() : 0;
  const size_t smem = pow2_smem_bytes<T>(N, p.W, mode);
  const long long gx = (nlines + p.W - 1) / p.W;
  FFB_REQUIRE(gx < (1ll << 31), FFB_EUNSUPPORTED, "too many lines for one launch");
  const bool cols_mode = mode == C2C_COLS || mode == C2C_COLS_TW || mode == C2C_COLS_LEAN;
  FFB_REQUIRE(!cols_mode || nlines < (1ll << 31), FFB_EUNSUPPORTED, "too many columns for a strided pass");   // col_coords: 32-bit column index
  static const char* mode_names[6] = {"c2c_rows", "c2c_cols", "r2c_rows", "c2r_rows", "c2c_cols_tw", "c2c_cols"};
  char pname[64];
  snprintf(pname, sizeof(pname), "fft_%s_%s_N%d", mode_names[mode], sizeof(T) == 8 ? "f64" : "f32", N);
  // algorithmic bytes: every element of the line set is read once and written once
  const double lines = (double)nlines * (double)nouter * (double)(o2.mod ? o2.nhi : 1);
  const double in_elems = (mode == C2R_ROWS) ? N + 1 : N, out_elems = (mode == R2C_ROWS) ? N + 1 : N;
  // fused operands are algorithmic traffic too: dense factor (one real per element), accumulated array, physical multiplier
  const double fused_bytes = lines * N * ((pro && pro->w ? sizeof(T) : 0) + (epi && epi->w ? sizeof(T) : 0) + (epi && epi->acc ? sizeof(cx<T>) : 0) +
                                          (rmul ? sizeof(cx<T>) : 0));
  ProfScope ps(pname, lines * (in_elems + out_elems) * sizeof(cx<T>) + fused_bytes);
  if (o2.mod) {
    // nouter = o2.mod low indices per high index; chunk over the high index so that gridDim.y <= 65535
    FFB_REQUIRE(o2.mod <= 65535, FFB_EUNSUPPORTED, "four-step factor too large");
    p.outer_mod = o2.mod;
    const long long hchunk = std::max<long long>(1, 65535 / o2.mod);
    for (long long h0 = 0; h0 < o2.nhi; h0 += hchunk) {
      const long long cnt = std::min<long long>(hchunk, o2.nhi - h0);
      p.in = reinterpret_cast<const cx<T>*>(in) + h0 * o2.in_os2;
      p.out = reinterpret_cast<cx<T>*>(out) + h0 * o2.out_os2;
      if (pro && pro->w) p.pro.w = pro->w + h0 * o2.in_os2;
      if (epi && epi->w) p.epi.w = epi->w + h0 * o2.out_os2;
      if (epi && epi->acc) p.epi.acc = epi->acc + h0 * o2.out_os2;
      int rc = call_pow2<T>(N, mode, dir, p, (int)gx, (int)(cnt * o2.mod), threads, smem, st, r8);
      if (rc) return rc;
    }
    return FFB_OK;
  }
  p.outer_mod = 1 << 30;
  for (long long o0 = 0; o0 < nouter; o0 += 65535) {
    const long long cnt = std::min<long long>(65535, nouter - o0);
    p.in = reinterpret_cast<const cx<T>*>(in) + o0 * in_os;
    p.out = reinterpret_cast<cx<T>*>(out) + o0 * out_os;
    if (mode == C2C_COLS_LEAN) {
      p.in_ts = p.W; p.out_ts = p.W;
      for (int m = 0; m < 16; ++m) p.out_m[m] = reinterpret_cast<cx<T>*>(p.out) + lean_out_off[m];
    }
    if (pro && pro->w) p.pro.w = pro->w + o0 * in_os;
    if (epi && epi->w) p.epi.w = epi->w + o0 * out_os;
    if (epi && epi->acc) p.epi.acc = epi->acc + o0 * out_os;
    if (pro || epi) FFB_REQUIRE(nouter <= 65535 || (!(pro && pro->ko) && !(epi && (epi->ko || epi->ao || epi->loo > 0))), FFB_EUNSUPPORTED,
                                "fused pass with more than 65535 outer slices");
    int rc = call_pow2<T>(N, mode, dir, p, (int)gx, (int)cnt, threads, smem, st, r8);
    if (rc) return rc;
  }
  return FFB_OK;
}

// Lean strided pass over tiles of W adjacent lines with explicit tile / outer / element strides, per-register input offsets
// and per-register output base pointers (used by the slab decomposition: blocked receive layouts, stores into peer memory).
template <typename T>
static int lean_tile_pass(int N, int dir, int W, const cx<T>* in, long long in_ts, long long in_os, long long in_es, const long long* in_off,
                          cx<T>* const* out_m, long long out_ts, long long out_os, long long out_es, long long nlines, long long nouter, T scale,
                          const cx<T>* tw, cudaStream_t st, const typename Pow2Params<T>::Fuse* epi = nullptr) {
  Pow2Params<T> p;
  p.rsq = 0; p.row_seg = 0; p.row_seg_mask = 0; p.row_seg_shift = 0; p.row_seg_stride = 0; p.row_nyq = 0;
  p.row_dead_lo = p.row_dead_hi = 0; p.dead = g_dead;
  const int R = pow2_points_per_thread(N), Tn = N / R;
  FFB_REQUIRE(R == 16 && Tn * W <= pow2_max_threads(sizeof(T)) && nouter <= 65535, FFB_EUNSUPPORTED,
              "blocked strided pass: line length %d with %d-wide tiles is outside the kernel range", N, W);
  p.pro.on = 0; p.epi.on = 0; p.rmul = nullptr; p.reverse = 0; p.keep_out = 0;
  if (epi) p.epi = *epi;
  p.in = in; p.out = nullptr;
  p.in_es = in_es; p.in_ls = 1; p.in_os = in_os; p.out_es = out_es; p.out_ls = 1; p.out_os = out_os;
  p.in_os2 = p.out_os2 = 0; p.outer_mod = 1 << 30;
  p.in_seg_mask = p.out_seg_mask = 0x7fffffff; p.in_seg_shift = p.out_seg_shift = 31; p.in_seg_stride = p.out_seg_stride = 0;
  p.twN = nullptr; p.twN_mask = 0; p.twr = nullptr;
  p.nlines = nlines; p.W = W; p.scale = scale; p.tw = tw;
  p.in_ts = in_ts; p.out_ts = out_ts;
  for (int m = 0; m < 16; ++m) { p.in_off[m] = in_off[m]; p.out_m[m] = out_m[m]; }
  const int threads = Tn * W;
  p.pf_ahead = threads >= pow2_max_threads(sizeof(T)) ? num_sms() : 0;
  const size_t smem = pow2_smem_bytes<T>(N, W, C2C_COLS_LEAN);
  const long long gx = (nlines + W - 1) / W;
  char pname[64];
  snprintf(pname, sizeof(pname), "fft_c2c_cols_%s_N%d", sizeof(T) == 8 ? "f64" : "f32", N);
  ProfScope ps(pname, (double)nlines * (double)nouter * 2.0 * N * sizeof(cx<T>));
  return call_pow2<T>(N, C2C_COLS_LEAN, dir, p, (int)gx, (int)nouter, threads, smem, st);
}


// Fused four-step pass (fft_l2four.cuh): both sub-passes in ONE persistent kernel, intermediate in an L2-resident scratch ring.
// Opt-in (FFB_L2FOUR=1): measured slower than the two-kernel form on B200 although it halves the DRAM traffic; FFB_L2_CHUNK = chunk width in units of the wider tile (default 1);
// FFB_L2_AHEAD = tiles of lookahead of A over B in units of 0.1 x resident CTAs (default 20).
static bool l2four_enabled(int N1, int N2) { return l2four_has(N1, N2) && env_int("FFB_L2FOUR", 0) != 0; }

template <typename T>
static void fs_params(FsParams<T>& p, int Nsub, long long in_es, long long out_es, T scale, const cx<T>* tw) {
  const int R = fs_points_per_thread((int)sizeof(T)), Tn = Nsub / R;
  memset(&p, 0, sizeof(p));
  p.in_es = in_es; p.out_es = out_es; p.in_ms = (long long)Tn * in_es; p.out_ms = (long long)Tn * out_es;
  p.W = kFsThreads / Tn; p.lgW = ilog2((uint64_t)p.W);
  p.scale = scale; p.tw = tw; p.twN = nullptr; p.twN_mask = 0;
  p.hook.on = 0;
  p.dead = g_dead;
}

// Stand-alone four-step sub-pass (two-kernel form; the fused form is l2four_pass).  part 1 = A, part 2 = B.
template <typename T>
static int fs_pass(const DimTables<T>* tb, int part, long long inner, long long outer, const cx<T>* src, cx<T>* dst, int dir, T scale,
                   cudaStream_t st, typename Pow2Params<T>::Fuse* pro, typename Pow2Params<T>::Fuse* epi) {
  const int N = tb->N, N1 = tb->N1, N2 = tb->N2;
  const bool is_a = part == 1;
  FsLaunch<T> q;
  memset(&q, 0, sizeof(q));
  if (is_a) {
    fs_params<T>(q.p, N1, (long long)N2 * inner, (long long)N2 * inner, T(1), tb->tw1);
    q.p.twN = tb->twN; q.p.twN_mask = N - 1;
    if (q.p.dead.on) q.p.dead.on = 1;   // sub-pass A only ever skips
    q.in_os = inner; q.out_os = inner; q.mod = N2;
    FFB_REQUIRE(!epi, FFB_EINVAL, "internal: epilogue on the first four-step sub-pass");
    if (pro) { pro->idm = N2; pro->ido = 1; q.p.hook = *pro; }
  } else {
    fs_params<T>(q.p, N2, inner, (long long)N1 * inner, scale, tb->tw2);
    q.in_os = (long long)N2 * inner; q.out_os = inner; q.mod = N1;
    FFB_REQUIRE(!pro, FFB_EINVAL, "internal: prologue on the second four-step sub-pass");
    if (epi) { epi->idm = N1; epi->ido = 1; q.p.hook = *epi; }
  }
  q.in_os2 = inner * N; q.out_os2 = inner * N; q.nlines = inner;
  const int hook = is_a ? (pro ? 1 : 0) : (epi ? 2 : 0);
  const int Nsub = is_a ? N1 : N2;
  const size_t smem = pow2_smem_bytes<T>(Nsub, q.p.W, C2C_COLS);
  const long long gx = (inner + q.p.W - 1) / q.p.W;
  FFB_REQUIRE(inner < (1ll << 31) && q.mod <= 65535, FFB_EUNSUPPORTED, "too many lines for one launch");   // fs_tile: 32-bit column index
  char pname[64];
  snprintf(pname, sizeof(pname), "fft_fs_%s_%s_N%d", is_a ? "a" : "b", sizeof(T) == 8 ? "f64" : "f32", Nsub);
  const double lines = (double)inner * (double)outer;
  const typename Pow2Params<T>::Fuse* h = is_a ? pro : epi;
  const double fused_bytes = lines * N * ((h && h->w ? sizeof(T) : 0) + (h && h->acc ? sizeof(cx<T>) : 0));
  ProfScope ps(pname, lines * 2.0 * N * sizeof(cx<T>) + fused_bytes);
  const long long hchunk = std::max<long long>(1, 65535 / q.mod);
  FFB_REQUIRE(outer <= hchunk || !(h && (h->ko || h->ao || h->loo > 0)), FFB_EUNSUPPORTED, "fused pass with more than 65535 outer slices");
  for (long long h0 = 0; h0 < outer; h0 += hchunk) {
    const long long cnt = std::min<long long>(hchunk, outer - h0);
    q.in = src + h0 * q.in_os2; q.out = dst + h0 * q.out_os2;
    if (h && h->w) q.p.hook.w = h->w + h0 * q.in_os2;
    if (h && h->acc) q.p.hook.acc = h->acc + h0 * q.out_os2;
    int rc = sizeof(T) == 8 ? fs_call_double(is_a, Nsub, dir, hook, &q, (int)gx, (int)(cnt * q.mod), smem, st)
                            : fs_call_float(is_a, Nsub, dir, hook, &q, (int)gx, (int)(cnt * q.mod), smem, st);
    if (rc == 1) return set_error(FFB_EUNSUPPORTED, "no four-step sub-pass kernel for N = %d", Nsub);
    if (rc) return rc;
  }
  return FFB_OK;
}

// Sub-pass A of `nv` inverse transforms of the same input (fft_fs.cuh: fs_pass_multi_kernel): variant v applies prologue pro[v] and
// writes dst[v].
template <typename T>
static int fs_pass_multi(const DimTables<T>* tb, long long inner, long long outer, const cx<T>* src, int nv, cx<T>* const* dst,
                         typename Pow2Params<T>::Fuse* pro, cudaStream_t st) {
  const int N = tb->N, N1 = tb->N1, N2 = tb->N2;
  FFB_REQUIRE(nv >= 1 && nv <= kFsMaxVariants, FFB_EINVAL, "1 to %d variants", kFsMaxVariants);
  FsMultiLaunch<T> mq;
  memset(&mq, 0, sizeof(mq));
  mq.nv = nv;
  double w_bytes = 0;
  const T* seen_w[kFsMaxVariants] = {nullptr};
  const double lines = (double)inner * (double)outer;
  for (int v = 0; v < nv; ++v) {
    FsLaunch<T>& q = mq.q[v];
    fs_params<T>(q.p, N1, (long long)N2 * inner, (long long)N2 * inner, T(1), tb->tw1);
    q.p.twN = tb->twN; q.p.twN_mask = N - 1;
    q.p.dead.on = 0;
    q.in_os = inner; q.out_os = inner; q.mod = N2;
    q.in_os2 = inner * N; q.out_os2 = inner * N; q.nlines = inner;
    pro[v].idm = N2; pro[v].ido = 1;
    q.p.hook = pro[v];
    bool dup = false;
    for (int u = 0; u < v; ++u) dup = dup || seen_w[u] == pro[v].w;
    seen_w[v] = pro[v].w;
    if (pro[v].w && !dup) w_bytes += lines * N * sizeof(T);   // a dense factor shared by several variants comes from DRAM once
    FFB_REQUIRE(!(pro[v].ko && outer > 65535 / N2), FFB_EUNSUPPORTED, "fused pass with more than 65535 outer slices");
  }
  const int W = mq.q[0].p.W;
  const size_t smem = pow2_smem_bytes<T>(N1, W, C2C_COLS);
  const long long gx = (inner + W - 1) / W;
  FFB_REQUIRE(inner < (1ll << 31) && N2 <= 65535, FFB_EUNSUPPORTED, "too many lines for one launch");
  char pname[64];
  snprintf(pname, sizeof(pname), "fft_fs_am_%s_N%d", sizeof(T) == 8 ? "f64" : "f32", N1);
  ProfScope ps(pname, lines * (1.0 + nv) * N * sizeof(cx<T>) + w_bytes);
  const long long hchunk = std::max<long long>(1, 65535 / N2);
  for (long long h0 = 0; h0 < outer; h0 += hchunk) {
    const long long cnt = std::min<long long>(hchunk, outer - h0);
    for (int v = 0; v < nv; ++v) {
      FsLaunch<T>& q = mq.q[v];
      q.in = src + h0 * q.in_os2; q.out = dst[v] + h0 * q.out_os2;
      if (pro[v].w) q.p.hook.w = pro[v].w + h0 * q.in_os2;
    }
    int rc = sizeof(T) == 8 ? fs_multi_call_double(N1, +1, &mq, (int)gx, (int)(cnt * N2), smem, st)
                            : fs_multi_call_float(N1, +1, &mq, (int)gx, (int)(cnt * N2), smem, st);
    if (rc == 1) return set_error(FFB_EUNSUPPORTED, "no four-step sub-pass kernel for N = %d", N1);
    if (rc) return rc;
  }
  return FFB_OK;
}

template <typename T>
static int l2four_pass(ffb_plan* pl, const DimTables<T>* tb, long long inner, long long outer, const cx<T>* src, cx<T>* dst, int dir, T scale,
                       cudaStream_t st, typename Pow2Params<T>::Fuse* pro, typename Pow2Params<T>::Fuse* epi) {
  const int N = tb->N, N1 = tb->N1, N2 = tb->N2;
  L2FourParams<T> q;
  memset(&q, 0, sizeof(q));
  fs_params<T>(q.a, N1, (long long)N2 * inner, 0, T(1), tb->tw1);
  fs_params<T>(q.b, N2, 0, (long long)N1 * inner, scale, tb->tw2);
  const int WA = q.a.W, WB = q.b.W, Wmax = std::max(WA, WB);
  int Wc = Wmax * std::max(1, env_int("FFB_L2_CHUNK", 2));
  FFB_REQUIRE(is_pow2((uint64_t)Wc), FFB_EINVAL, "FFB_L2_CHUNK must be a power of two");
  while (Wc > Wmax && Wc / 2 >= inner) Wc /= 2;
  // scratch chunk layout [N][Wc]: A writes k1 in place of n1 (stride N2*Wc), B reads n2 (stride Wc)
  q.a.out_es = (long long)N2 * Wc; q.a.out_ms = (long long)(N1 / fs_points_per_thread((int)sizeof(T))) * q.a.out_es;
  q.b.in_es = Wc; q.b.in_ms = (long long)(N2 / fs_points_per_thread((int)sizeof(T))) * q.b.in_es;
  q.a.twN = tb->twN; q.a.twN_mask = N - 1;
  if (q.a.dead.on) q.a.dead.on = 1;   // sub-pass A only ever skips; B zero-fills when this is the transform's last pass
  if (pro) { pro->idm = N2; pro->ido = 1; q.a.hook = *pro; }
  if (epi) { epi->idm = N1; epi->ido = 1; q.b.hook = *epi; }
  FFB_REQUIRE(!(pro && epi), FFB_EINVAL, "internal: prologue and epilogue on one transform");
  const int hook = pro ? 1 : (epi ? 2 : 0);
  q.in = src; q.out = dst; q.N = N;
  q.Wc = Wc; q.inner = inner;
  q.ncc = (int)((inner + Wc - 1) / Wc);
  FFB_REQUIRE((long long)q.ncc * outer < (1ll << 22), FFB_EUNSUPPORTED, "too many chunks for the fused four-step pass");
  q.C = (int)(q.ncc * outer);
  q.Cpad = (q.C + 3) / 4 * 4;
  q.lgWA = q.a.lgW; q.lgWB = q.b.lgW;
  q.lg_tca = ilog2((uint64_t)(Wc / WA)); q.lg_tcb = ilog2((uint64_t)(Wc / WB));
  const int tiles = (Wc / WA) * N2;   // == (Wc / WB) * N1
  q.lgT = ilog2((uint64_t)tiles);
  FFB_REQUIRE((1 << q.lgT) == tiles && (Wc / WB) * N1 == tiles, FFB_EUNSUPPORTED, "internal: tile counts of the two sub-passes differ");
  FFB_REQUIRE(((long long)q.C << (q.lgT + 1)) < (1ll << 31), FFB_EUNSUPPORTED, "too many tiles for the fused four-step pass");
  const size_t smem = std::max(pow2_smem_bytes<T>(N1, WA, C2C_COLS), pow2_smem_bytes<T>(N2, WB, C2C_COLS));
  auto call = [&](int op, int grid) {
    return sizeof(T) == 8 ? l2four_call_double(op, N1, N2, dir, hook, &q, grid, smem, st) : l2four_call_float(op, N1, N2, dir, hook, &q, grid, smem, st);
  };
  const int per_sm = call(1, 0);
  FFB_REQUIRE(per_sm >= 1, per_sm < 0 ? per_sm : FFB_EUNSUPPORTED, "fused four-step kernel does not fit an SM (N = %d x %d)", N1, N2);
  const int resident = per_sm * num_sms();
  // A runs far enough ahead of B that a chunk is complete when the first tile that needs it starts: at that moment the oldest
  // unfinished ticket is about one wave of resident CTAs behind, so the tickets between the end of A(c) and the start of B(c),
  // (2D - 1) * tiles, must exceed `ahead` waves; the same distance separates B(c) from the A(c + nslots) that reuses its slot.
  const double ahead = 0.1 * env_int("FFB_L2_AHEAD", 20);
  int D = (int)std::ceil((ahead * resident / tiles + 1.0) / 2.0);
  D = std::max(1, std::min(D, q.C));
  q.D = D;
  q.nslots = D + (int)std::ceil((ahead * resident / tiles + 1.0) / 2.0) + 1;
  q.slot_elems = (long long)N * Wc;
  const size_t ring_bytes = (size_t)q.nslots * q.slot_elems * sizeof(cx<T>);
  if (pl->ring_bytes < ring_bytes) {
    if (pl->ring) { FFB_CUDA(cudaStreamSynchronize(st)); cudaFree(pl->ring); pl->ring = nullptr; pl->ring_bytes = 0; }
    int rc = ffb_malloc(&pl->ring, ring_bytes);
    if (rc) return rc;
    pl->ring_bytes = ring_bytes;
  }
  const size_t nctr = 4 + 2 * (size_t)q.Cpad + 4;
  if (pl->ctr_count < nctr) {
    if (pl->ctr) { FFB_CUDA(cudaStreamSynchronize(st)); cudaFree(pl->ctr); pl->ctr = nullptr; pl->ctr_count = 0; }
    void* c = nullptr;
    int rc = ffb_malloc(&c, nctr * sizeof(unsigned));
    if (rc) return rc;
    pl->ctr = reinterpret_cast<unsigned*>(c); pl->ctr_count = nctr;
    pl->ctr_C = -1;
  }
  if (pl->ctr_C != q.Cpad) {
    // counter layout depends on the chunk count: start from zero (afterwards every launch leaves the counters zeroed)
    FFB_CUDA(cudaMemsetAsync(pl->ctr, 0, pl->ctr_count * sizeof(unsigned), st));
    pl->ctr_C = q.Cpad;
  }
  q.ring = reinterpret_cast<cx<T>*>(pl->ring);
  q.ctr = pl->ctr;
  q.pf = env_int("FFB_L2_PF", 0);
  q.acq = env_int("FFB_L2_ACQ", 1);
  char pname[64];
  snprintf(pname, sizeof(pname), "fft_l2four_%s_N%d", sizeof(T) == 8 ? "f64" : "f32", N);
  const double lines = (double)inner * (double)outer;
  const double fused_bytes = lines * N * ((pro && pro->w ? sizeof(T) : 0) + (epi && epi->w ? sizeof(T) : 0) + (epi && epi->acc ? sizeof(cx<T>) : 0));
  ProfScope ps(pname, lines * 2.0 * N * sizeof(cx<T>) + fused_bytes);
  const long long total = (long long)q.C << (q.lgT + 1);
  const int grid = (int)std::min<long long>(resident, total);
  static unsigned long long* dbg_dev = nullptr;
  const int dbg = env_int("FFB_L2_DEBUG", 0);
  if (dbg) {
    if (!dbg_dev) FFB_CUDA(cudaMalloc(&dbg_dev, 6 * sizeof(unsigned long long)));
    FFB_CUDA(cudaMemsetAsync(dbg_dev, 0, 6 * sizeof(unsigned long long), st));
    q.dbg = dbg_dev;
  }
  int rc = call(0, grid);
  if (dbg && rc == FFB_OK) {
    unsigned long long h[6];
    FFB_CUDA(cudaMemcpyAsync(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
    FFB_CUDA(cudaStreamSynchronize(st));
    const double loop = (double)h[3];
    fprintf(stderr, "[l2four N=%dx%d dir=%d Wc=%d C=%d D=%d slots=%d grid=%d tiles/chunk=%d] per-CTA mean: loop %.0f cyc, tiles %.1f, wait-slot %.1f%%, wait-A %.1f%%, publish %.1f%%, waits %.1f\n",
            N1, N2, dir, Wc, q.C, q.D, q.nslots, grid, tiles, loop / grid, (double)h[4] / grid, 100.0 * h[0] / loop, 100.0 * h[1] / loop, 100.0 * h[2] / loop, (double)h[5] / grid);
  }
  return rc == 1 ? set_error(FFB_EUNSUPPORTED, "no fused four-step kernel for N = %d x %d", N1, N2) : rc;
}

// strided (column) pass along dimension d: one register-resident pass, or the four-step pair A (in place allowed) + B
// (strictly out of place).  part: 0 = single pass, 1 = four-step A, 2 = four-step B.
template <typename T>
static int cols_pass(ffb_plan* pl, const DimTables<T>* tb, int part, long long inner, long long outer, const cx<T>* src, cx<T>* dst, int dir, T scale,
                     cudaStream_t st, typename Pow2Params<T>::Fuse* pro = nullptr, typename Pow2Params<T>::Fuse* epi = nullptr) {
  const int N = tb->N;
  if (part == 4) return l2four_pass<T>(pl, tb, inner, outer, src, dst, dir, scale, st, pro, epi);
  if (part == 0) {
    // single pass: transform index = t + m*Tn, the outer slice index is o_lo
    if (pro) { pro->idm = 1; pro->ido = 0; if (pro->other_from_col == 0) pro->other_from_col = 2; }
    if (epi) { epi->idm = 1; epi->ido = 0; if (epi->other_from_col == 0) epi->other_from_col = 2; }
    return pow2_pass<T>(N, C2C_COLS, dir, src, dst, inner, 1, inner * N, inner, 1, inner * N, inner, outer, scale, tb->tw, nullptr, st,
                        SegStride(), SegStride(), Outer2(), nullptr, 0, pro, epi);
  }
  return fs_pass<T>(tb, part, inner, outer, src, dst, dir, scale, st, pro, epi);
}

// public ffb_fuse -> kernel hook for the strided pass along dimension d of a (e0, e1, e2) spectral array
template <typename T>
static typename Pow2Params<T>::Fuse make_hook(const ffb_fuse* f, int d, int nd, const long long e[3], bool epilogue) {
  typename Pow2Params<T>::Fuse h;
  memset(&h, 0, sizeof(h));
  h.on = 1;
  h.cr = (T)f->cr; h.ci = (T)f->ci;
  const void* vec[3] = {f->kx, f->l, f->m};
  const void* avec[3] = {f->akx, f->al, f->am};
  const int other = (nd == 3) ? (d == 1 ? 2 : 1) : -1;
  h.k0 = (const T*)vec[0]; h.kt = (const T*)vec[d]; h.ko = other >= 0 ? (const T*)vec[other] : nullptr;
  h.w = (const T*)f->w;
  h.n0 = (int)e[0];
  h.other_from_col = (nd == 3 && d == 2) ? 1 : 0;   // z-pass: other = y = col / n0; y-pass: other = z = outer slice
  if (epilogue) {
    h.acc = (const cx<T>*)f->acc; h.ar = (T)f->ar; h.ai = (T)f->ai;
    h.a0 = (const T*)avec[0]; h.at = (const T*)avec[d]; h.ao = other >= 0 ? (const T*)avec[other] : nullptr;
    h.dealias = f->dealias;
    h.lo0 = f->alias_lo[0]; h.hi0 = f->alias_hi[0];
    h.lot = f->alias_lo[d]; h.hit = f->alias_hi[d];
    if (other >= 0) { h.loo = f->alias_lo[other]; h.hio = f->alias_hi[other]; }
  }
  return h;
}

// c2c pass along dimension d of a dense complex array with extents e[0..2] x nb (x fastest).
template <typename T>
static int c2c_dim(ffb_plan* pl, int d, const long long e[3], long long nb, const cx<T>* src, cx<T>* dst, int dir, T scale,
                   cudaStream_t st) {
  auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[d]);
  const int N = tb->N;
  long long inner = 1, outer = nb;
  for (int i = 0; i < d; ++i) inner *= e[i];
  for (int i = d + 1; i < 3; ++i) outer *= e[i];
  if (tb->pow2 && tb->tw) {
    if (d == 0)
      return pow2_pass<T>(N, C2C_ROWS, dir, src, dst, 1, N, 0, 1, N, 0, outer, 1, scale, tb->tw, nullptr, st);
    return cols_pass<T>(pl, tb, 0, inner, outer, src, dst, dir, scale, st);
  }
  FFB_REQUIRE(tb->wN, FFB_EUNSUPPORTED, "dimension %d of this plan mixes the arbitrary-size path with a four-step-only length", d);
  int rc = ensure_ws(pl, 2);
  if (rc) return rc;
  return generic_fft_axis<T>(src, dst, reinterpret_cast<cx<T>*>(pl->ws[0]), reinterpret_cast<cx<T>*>(pl->ws[1]), inner, N, outer,
                             dir, scale, tb->wN, st);
}

// All-power-of-two plans: a list of passes (x rows pass, then one single pass or a four-step pair per strided dimension)
// is scheduled over {IN (never written), OUT, WS0, WS1} so that the result lands in OUT and strictly out-of-place passes
// (r2c, c2r, four-step B) never alias.
template <typename T>
static int exec_pow2(ffb_plan* pl, const void* in, void* out, int dir, const ffb_fuse* fuse = nullptr, const void* a_done = nullptr) {
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  const int nd = pl->ndim;
  const long long nb = pl->nbatch;
  struct Op { int kind, d, part; bool strict; };  // kind: 0 = c2c rows, 1 = r2c rows, 2 = c2r rows, 3 = strided pass
  std::vector<Op> ops;
  auto push_cols = [&](int d) {
    auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[d]);
    if (tb->four && l2four_enabled(tb->N1, tb->N2)) ops.push_back({3, d, 4, false});   // fused, L2-resident intermediate, in place allowed
    else if (tb->four) { ops.push_back({3, d, 1, false}); ops.push_back({3, d, 2, true}); }
    else ops.push_back({3, d, 0, false});
  };
  if (pl->kind == FFB_C2C) { ops.push_back({0, 0, 0, false}); for (int d = 1; d < nd; ++d) push_cols(d); }
  else if (dir < 0) { ops.push_back({1, 0, 0, true}); for (int d = 1; d < nd; ++d) push_cols(d); }
  else { for (int d = nd - 1; d >= 1; --d) push_cols(d); ops.push_back({2, 0, 0, true}); }
  const int n = (int)ops.size();
  enum { IN = 0, OUT = 1, WS0 = 2, WS1 = 3 };
  std::vector<int> src(n), dst(n);
  dst[n - 1] = OUT;
  int need_ws = 0;
  for (int i = n - 1; i >= 0; --i) {
    if (i == 0) src[i] = IN;
    else if (ops[i].strict) src[i] = (dst[i] != WS0) ? WS0 : WS1;
    else src[i] = dst[i];
    if (i > 0) dst[i - 1] = src[i];
    need_ws = std::max(need_ws, std::max(src[i], dst[i]) - 1);
  }
  if (need_ws > 0) { int rc = ensure_ws(pl, need_ws); if (rc) return rc; }
  auto buf = [&](int b) -> void* { return b == IN ? const_cast<void*>(in) : b == OUT ? out : pl->ws[b - WS0]; };
  // a_done: the first op (four-step sub-pass A, prologue included) was already run into that array by exec_pow2_multi
  if (a_done) FFB_REQUIRE(n >= 2 && ops[0].kind == 3 && ops[0].part == 1, FFB_EINVAL, "internal: a_done without a four-step first pass");
  long double tot = 1;
  for (int d = 0; d < nd; ++d) tot *= (long double)pl->n[d];
  const T inv = (T)(1.0L / tot);
  const long long e[3] = {pl->nc[0], pl->nc[1], pl->nc[2]};
  auto* tb0 = reinterpret_cast<DimTables<T>*>(pl->tables[0]);
  long long rows = nb;
  for (int d = 1; d < nd; ++d) rows *= pl->n[d];
  if (fuse) {
    FFB_REQUIRE(pl->kind == FFB_R2C && nd >= 2 && nb == 1, FFB_EUNSUPPORTED, "fused transforms need an r2c plan with ndim >= 2 and one field");
    FFB_REQUIRE(dir > 0 ? ops[0].kind == 3 : ops[n - 1].kind == 3, FFB_EUNSUPPORTED, "no strided pass to fuse into");
  }
  const bool want_pro = fuse && dir > 0 && (fuse->kx || fuse->l || fuse->m || fuse->w || fuse->cr != 1.0 || fuse->ci != 0.0);
  for (int i = a_done ? 1 : 0; i < n; ++i) {
    const Op& op = ops[i];
    const void* s_ = (a_done && i == 1) ? a_done : buf(src[i]);
    void* d_ = buf(dst[i]);
    const T sc = (dir > 0 && i == n - 1) ? inv : T(1);
    g_pass_reverse = snake_enabled() ? (i & 1) : 0;
    g_pass_keep = (snake_enabled() && i + 1 < n) ? 1 : 0;
    // forward transform followed by dealias!: the aliased columns are not stored by the x pass, skipped by intermediate strided
    // passes and zero-filled, unread, by the last pass (FFB_DEAD_SKIP=0 disables)
    g_row_dead_lo = g_row_dead_hi = 0;
    g_dead = DeadCols{0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fuse && dir < 0 && fuse->dealias && env_int("FFB_DEAD_SKIP", 1)) {
      if (op.kind == 1 && fuse->alias_lo[0] > 0) { g_row_dead_lo = fuse->alias_lo[0] - 1; g_row_dead_hi = fuse->alias_hi[0]; }
      if (op.kind == 3) {
        g_dead.on = (i == n - 1 && fuse->dealias != 2) ? 2 : 1;
        g_dead.n0 = (int)e[0];
        if (fuse->alias_lo[0] > 0) { g_dead.dlo = fuse->alias_lo[0] - 1; g_dead.dhi = fuse->alias_hi[0]; }
        if (op.d == 2 && fuse->alias_lo[1] > 0) { g_dead.olo = fuse->alias_lo[1] - 1; g_dead.ohi = fuse->alias_hi[1]; }
        if (op.part == 4 && i != n - 1) g_dead.on = 1;
      }
    }
    int rc;
    if (op.kind == 0) rc = pow2_pass<T>(tb0->N, C2C_ROWS, dir, s_, d_, 1, tb0->N, 0, 1, tb0->N, 0, rows, 1, sc, tb0->tw, nullptr, st);
    else if (op.kind == 1)
      rc = pow2_pass<T>(tb0->N, R2C_ROWS, -1, s_, d_, 1, tb0->N, 0, 1, e[0], 0, rows, 1, T(1), tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(),
                        nullptr, 0, nullptr, nullptr, nullptr, (fuse && fuse->square_input) ? 1 : 0);
    else if (op.kind == 2)
      rc = pow2_pass<T>(tb0->N, C2R_ROWS, +1, s_, d_, 1, e[0], 0, 1, tb0->N, 0, rows, 1, sc, tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(),
                        nullptr, 0, nullptr, nullptr, fuse ? reinterpret_cast<const T*>(fuse->mul) : nullptr);
    else {
      long long inner = 1, outer = nb;
      for (int q = 0; q < op.d; ++q) inner *= e[q];
      for (int q = op.d + 1; q < 3; ++q) outer *= e[q];
      typename Pow2Params<T>::Fuse hook;
      typename Pow2Params<T>::Fuse* pro = nullptr;
      typename Pow2Params<T>::Fuse* epi = nullptr;
      if (want_pro && i == 0) { hook = make_hook<T>(fuse, op.d, nd, e, false); pro = &hook; }
      if (fuse && dir < 0 && i == n - 1) { hook = make_hook<T>(fuse, op.d, nd, e, true); epi = &hook; }
      rc = cols_pass<T>(pl, reinterpret_cast<DimTables<T>*>(pl->tables[op.d]), op.part, inner, outer, reinterpret_cast<const cx<T>*>(s_),
                        reinterpret_cast<cx<T>*>(d_), dir, sc, st, pro, epi);
    }
    g_pass_reverse = 0; g_pass_keep = 0;
    g_row_dead_lo = g_row_dead_hi = 0;
    g_dead = DeadCols{0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
    if (rc) return rc;
  }
  return FFB_OK;
}

// `nv` inverse transforms of the same spectral array with different prologues.  When the first strided pass is a four-step pair, its
// sub-pass A runs once for all variants (fs_pass_multi: the input is read from DRAM once) into per-variant scratch, and each variant
// then finishes through the usual pass list.  Otherwise (or with FFB_MULTI_A=0): nv independent fused inverse transforms.
template <typename T>
static int exec_pow2_multi(ffb_plan* pl, const void* in, int nv, void* const* outs, const ffb_fuse* fuses) {
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  const int nd = pl->ndim;
  auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[nd - 1]);
  const bool shared_a = nv >= 2 && nv <= kFsMaxVariants && pl->kind == FFB_R2C && nd >= 2 && pl->nbatch == 1 && tb->four &&
                        !l2four_enabled(tb->N1, tb->N2) && env_int("FFB_MULTI_A", 1) != 0;
  if (!shared_a) {
    for (int v = 0; v < nv; ++v) {
      int rc = exec_pow2<T>(pl, in, outs[v], +1, &fuses[v]);
      if (rc) return rc;
    }
    return FFB_OK;
  }
  const long long e[3] = {pl->nc[0], pl->nc[1], pl->nc[2]};
  const int d = nd - 1;
  long long inner = 1;
  for (int q = 0; q < d; ++q) inner *= e[q];
  typename Pow2Params<T>::Fuse pro[kFsMaxVariants];
  cx<T>* dst[kFsMaxVariants];
  for (int v = 0; v < nv; ++v) {
    if (!pl->wsm[v]) { int rc = ffb_malloc(&pl->wsm[v], pl->ws_bytes); if (rc) return rc; }
    dst[v] = reinterpret_cast<cx<T>*>(pl->wsm[v]);
    pro[v] = make_hook<T>(&fuses[v], d, nd, e, false);
  }
  int rc = fs_pass_multi<T>(tb, inner, 1, reinterpret_cast<const cx<T>*>(in), nv, dst, pro, st);
  if (rc) return rc;
  for (int v = 0; v < nv; ++v)
    if ((rc = exec_pow2<T>(pl, in, outs[v], +1, &fuses[v], pl->wsm[v]))) return rc;
  return FFB_OK;
}

template <typename T>
static int exec(ffb_plan* pl, const void* in, void* out, int dir) {
  {
    bool allp = true;
    for (int d = 0; d < pl->ndim; ++d) {
      auto* tb = reinterpret_cast<DimTables<T>*>(pl->tables[d]);
      allp = allp && tb->pow2 && (tb->four || tb->tw);
    }
    if (allp) return exec_pow2<T>(pl, in, out, dir);
  }
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  const int nd = pl->ndim;
  const long long nb = pl->nbatch;
  long long e[3] = {pl->nc[0], pl->nc[1], pl->nc[2]};
  long double tot = 1;
  for (int d = 0; d < nd; ++d) tot *= (long double)pl->n[d];
  const T inv = (T)(1.0L / tot);

  if (pl->kind == FFB_C2C) {
    const cx<T>* src = reinterpret_cast<const cx<T>*>(in);
    cx<T>* dst = reinterpret_cast<cx<T>*>(out);
    for (int d = 0; d < nd; ++d) {
      const T sc = (dir > 0 && d == nd - 1) ? inv : T(1);
      int rc = c2c_dim<T>(pl, d, e, nb, d == 0 ? src : dst, dst, dir, sc, st);
      if (rc) return rc;
    }
    return FFB_OK;
  }

  // ---------------- R2C plans ----------------
  auto* tb0 = reinterpret_cast<DimTables<T>*>(pl->tables[0]);
  const int N0 = tb0->N;               // nx/2
  const long long nkr = pl->nc[0];     // nx/2 + 1
  long long rows = nb;
  for (int d = 1; d < nd; ++d) rows *= pl->n[d];
  if (dir < 0) {
    // forward: x (r2c) into `out`, then y, z in place on `out`
    cx<T>* dst = reinterpret_cast<cx<T>*>(out);
    if (tb0->pow2) {
      int rc = pow2_pass<T>(N0, R2C_ROWS, -1, in, dst, 1, N0, 0, 1, nkr, 0, rows, 1, T(1), tb0->tw, tb0->twr, st);
      if (rc) return rc;
    } else {
      int rc = ensure_ws(pl, 3);
      if (rc) return rc;
      cx<T>* z = reinterpret_cast<cx<T>*>(pl->ws[2]);
      rc = generic_fft_axis<T>(reinterpret_cast<const cx<T>*>(in), z, reinterpret_cast<cx<T>*>(pl->ws[0]),
                               reinterpret_cast<cx<T>*>(pl->ws[1]), 1, N0, rows, -1, T(1), tb0->wN, st);
      if (rc) return rc;
      const long long total = rows * (N0 + 1);
      generic_r2c_post_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z, dst, N0, rows, tb0->twr);
      count_launch();
      FFB_CHECK_LAUNCH();
    }
    for (int d = 1; d < nd; ++d) {
      int rc = c2c_dim<T>(pl, d, e, nb, dst, dst, -1, T(1), st);
      if (rc) return rc;
    }
    return FFB_OK;
  }
  // inverse: z, y on a scratch copy of the spectrum (input preserved), then x (c2r) into `out`
  const cx<T>* spec = reinterpret_cast<const cx<T>*>(in);
  if (nd > 1) {
    int rc = ensure_ws(pl, tb0->pow2 ? 1 : 3);
    if (rc) return rc;
    // scratch index 2 (generic) or 0 (pow2-only plans) holds the partially transformed spectrum
    cx<T>* w = reinterpret_cast<cx<T>*>(pl->ws[tb0->pow2 ? 0 : 2]);
    bool all_pow2 = true;
    for (int d = 1; d < nd; ++d) { auto* tbd = reinterpret_cast<DimTables<T>*>(pl->tables[d]); all_pow2 = all_pow2 && tbd->pow2 && tbd->tw; }
    if (!all_pow2 && tb0->pow2) {  // generic y/z passes need ws[0], ws[1] as their own scratch
      rc = ensure_ws(pl, 3);
      if (rc) return rc;
      w = reinterpret_cast<cx<T>*>(pl->ws[2]);
    }
    for (int d = nd - 1; d >= 1; --d) {
      rc = c2c_dim<T>(pl, d, e, nb, d == nd - 1 ? spec : w, w, +1, T(1), st);
      if (rc) return rc;
    }
    spec = w;
  }
  if (tb0->pow2) return pow2_pass<T>(N0, C2R_ROWS, +1, spec, out, 1, nkr, 0, 1, N0, 0, rows, 1, inv, tb0->tw, tb0->twr, st);
  int rc = ensure_ws(pl, 3);
  if (rc) return rc;
  // generic x: pre-process into ws[2]... but ws[2] may hold `spec`; use ws[0] for Z and route the axis scratch via ws[1]/out
  cx<T>* z = reinterpret_cast<cx<T>*>(pl->ws[0]);
  const long long total = rows * N0;
  generic_c2r_pre_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(spec, z, N0, rows, tb0->twr);
  count_launch();
  FFB_CHECK_LAUNCH();
  // ws[2] is free again once the pre-process has consumed `spec` (stream ordered)
  cx<T>* tA = reinterpret_cast<cx<T>*>(pl->ws[1]);
  cx<T>* tB = reinterpret_cast<cx<T>*>(pl->ws[2]);
  return generic_fft_axis<T>(z, reinterpret_cast<cx<T>*>(out), tA, tB, 1, N0, rows, +1, inv, tb0->wN, st);
}

// ---------------------------------------------------------------- slab-decomposed 3-D r2c / c2r
// forward : x r2c + y c2c on the local z-slab, y-pass output written in destination-rank-major order (segmented stride,
//           no pack kernel) -> all-to-all over NCCL, chunked along z so chunk c's exchange overlaps chunk c+1's y-pass
//           -> z c2c on the local y-slab.   inverse: mirrored (z, exchange, y, x c2r with 1/(nx ny nz)).
template <typename T>
static int exec_dist(ffb_plan* pl, const void* in, void* out, int dir, const ffb_fuse* fuse = nullptr) {
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  ffb_dist* d = pl->dist;
  const int P = d->nranks;
  auto* tb0 = reinterpret_cast<DimTables<T>*>(pl->tables[0]);
  auto* tb1 = reinterpret_cast<DimTables<T>*>(pl->tables[1]);
  auto* tb2 = reinterpret_cast<DimTables<T>*>(pl->tables[2]);
  const long long nkr = pl->nc[0], ny = pl->n[1], nz = pl->n[2], nyl = pl->nyl, nzl = pl->nzl;
  const int N0 = tb0->N;
  const long long blk = nkr * nyl * nzl;         // complex elements exchanged with each peer
  const int nch = pl->nchunks;
  const long long zc = nzl / nch, sub = nkr * nyl * zc;
  // scratch: the NCCL exchange needs three slab-sized arrays, the copy-engine exchange w0 / w1 / w2 as well; the peer-store exchange
  // only one (forward: r2c output, inverse: y-pass output -- never live together), which is what lets 2048^3 LSRK54 fit two GPUs
  const int nws = pl->p2p == 1 ? 1 : 3;
  int rc = ensure_ws(pl, nws);
  if (rc) return rc;
  cx<T>* w0 = reinterpret_cast<cx<T>*>(pl->ws[0]);
  cx<T>* w1 = reinterpret_cast<cx<T>*>(pl->ws[nws == 1 ? 0 : 1]);
  cx<T>* w2 = reinterpret_cast<cx<T>*>(pl->ws[nws == 1 ? 0 : 2]);
  long double tot = (long double)pl->n[0] * ny * nz;
  const T inv = (T)(1.0L / tot);
  SegStride seg; seg.seg = (int)nyl; seg.stride = blk;
  // fused forward transform (ffb_fft_forward_ex on a slab-decomposed plan): the real input is squared in the x pass, the
  // spectral factor and the dealias mask are applied by the last (z) pass.  All coordinates are local: `l` and the y alias
  // range are this rank's slices (the caller passes them that way, like every other per-slab operand).
  const int rsq = (fuse && fuse->square_input) ? 1 : 0;
  typename Pow2Params<T>::Fuse ehook;
  const typename Pow2Params<T>::Fuse* epi = nullptr;
  if (fuse) {
    FFB_REQUIRE(dir < 0, FFB_EUNSUPPORTED, "fused transforms on slab-decomposed plans: forward only");
    FFB_REQUIRE(!fuse->acc && !fuse->w && !fuse->mul, FFB_EUNSUPPORTED, "fused slab-decomposed forward transform: square_input, scalar / wavenumber factors and dealias only");
    FFB_REQUIRE(pl->p2p != 2, FFB_EUNSUPPORTED, "fused transforms are not implemented for the copy-engine exchange");
    const long long e[3] = {nkr, nyl, nz};
    ehook = make_hook<T>(fuse, 2, 3, e, true);   // z pass: i0 = kx, other = y (local), transform index = z
    ehook.idm = 1; ehook.ido = 0;
    epi = &ehook;
  }
  if (pl->p2p == 2) {
    // ---- copy-engine exchange, pipelined over kx-chunks: the pass before the exchange writes chunk c (a range of kx) in
    //      destination-rank-major order, cudaMemcpyAsync pushes the blocks into the peers' receive buffers over NVLink (no SM,
    //      runs beside the passes of the neighbouring chunks), a one-element all-reduce on the communication stream tells the
    //      receiver that chunk c has landed everywhere, and the pass after the exchange starts on chunk c right away ----
    const int cur = pl->p2p_cur;
    pl->p2p_cur ^= 1;
    cx<T>* mine = reinterpret_cast<cx<T>*>(pl->recv[cur]);
    const size_t esz = sizeof(cx<T>);
    const int nc = (int)std::min<long long>(nch, nkr);
    const long long kw = nkr / nc;                       // chunk width; the last chunk takes the remainder
    cudaEvent_t landed[8];
    auto k0_of = [&](int c) { return (long long)c * kw; };
    auto kxc_of = [&](int c) { return c == nc - 1 ? nkr - (long long)c * kw : kw; };
    if (dir < 0) {
      if ((rc = pow2_pass<T>(N0, R2C_ROWS, -1, in, w0, 1, N0, 0, 1, nkr, 0, ny * nzl, 1, T(1), tb0->tw, tb0->twr, st))) return rc;
      for (int c = 0; c < nc; ++c) {
        const long long k0 = k0_of(c), kxc = kxc_of(c), blkc = kxc * nyl * nzl, off = k0 * nyl * nzl * P;
        SegStride segc; segc.seg = (int)nyl; segc.stride = blkc;
        // y on kx-chunk c: w0 (nkr, ny, nzl) -> w1 chunk [peer][kxc, nyl, nzl]
        if ((rc = pow2_pass<T>((int)ny, C2C_COLS, -1, w0 + k0, w1 + off, nkr, 1, nkr * ny, kxc, 1, kxc * nyl, kxc, nzl, T(1), tb1->tw, nullptr, st,
                               SegStride(), segc))) return rc;
        if ((rc = dist_push_blocks(d, w1 + off, pl->peers[cur], ((size_t)off + (size_t)d->rank * blkc) * esz, (size_t)blkc * esz, (size_t)blkc * esz, st))) return rc;
        if ((rc = dist_barrier(d, d->comm_stream))) return rc;
        landed[c] = dist_next_event(d);   // ring of 256 events, at most 10 per chunk and 8 chunks: no wrap-around before the waits below
        FFB_CUDA(cudaEventRecord(landed[c], d->comm_stream));
      }
      for (int c = 0; c < nc; ++c) {
        const long long k0 = k0_of(c), kxc = kxc_of(c), off = k0 * nyl * nzl * P;
        FFB_CUDA(cudaStreamWaitEvent(st, landed[c], 0));
        // z on kx-chunk c: receive chunk [kxc, nyl, nz] -> out (nkr, nyl, nz)
        if ((rc = pow2_pass<T>((int)nz, C2C_COLS, -1, mine + off, reinterpret_cast<cx<T>*>(out) + k0, kxc * nyl, 1, kxc, nkr * nyl, 1, nkr, kxc, nyl, T(1),
                               tb2->tw, nullptr, st))) return rc;
      }
      return FFB_OK;
    }
    for (int c = 0; c < nc; ++c) {
      const long long k0 = k0_of(c), kxc = kxc_of(c), blkc = kxc * nyl * nzl, off = k0 * nyl * nzl * P;
      // z on kx-chunk c: in (nkr, nyl, nz) -> w0 chunk [kxc, nyl, nz] (= [peer][kxc, nyl, nzl])
      if ((rc = pow2_pass<T>((int)nz, C2C_COLS, +1, reinterpret_cast<const cx<T>*>(in) + k0, w0 + off, nkr * nyl, 1, nkr, kxc * nyl, 1, kxc, kxc, nyl, T(1),
                             tb2->tw, nullptr, st))) return rc;
      if ((rc = dist_push_blocks(d, w0 + off, pl->peers[cur], ((size_t)off + (size_t)d->rank * blkc) * esz, (size_t)blkc * esz, (size_t)blkc * esz, st))) return rc;
      if ((rc = dist_barrier(d, d->comm_stream))) return rc;
      landed[c] = dist_next_event(d);
      FFB_CUDA(cudaEventRecord(landed[c], d->comm_stream));
    }
    for (int c = 0; c < nc; ++c) {
      const long long k0 = k0_of(c), kxc = kxc_of(c), blkc = kxc * nyl * nzl, off = k0 * nyl * nzl * P;
      SegStride segc; segc.seg = (int)nyl; segc.stride = blkc;
      FFB_CUDA(cudaStreamWaitEvent(st, landed[c], 0));
      // y on kx-chunk c: receive chunk [sender][kxc, nyl, nzl] -> w2 (nkr, ny, nzl)
      if ((rc = pow2_pass<T>((int)ny, C2C_COLS, +1, mine + off, w2 + k0, kxc, 1, kxc * nyl, nkr, 1, nkr * ny, kxc, nzl, T(1), tb1->tw, nullptr, st,
                             segc, SegStride()))) return rc;
    }
    return pow2_pass<T>(N0, C2R_ROWS, +1, w2, out, 1, nkr, 0, 1, N0, 0, ny * nzl, 1, inv, tb0->tw, tb0->twr, st);
  }
  if (pl->p2p == 1) {
    // ---- fused pass + collective: the pass before the exchange stores straight into the peers' receive buffers (NVLink),
    //      one stream-ordered barrier replaces the all-to-all; receive buffers are double-buffered across transforms.
    //      The receive layout is blocked, [kx block of B][..][..][B] with B*sizeof(complex) = 64 bytes, chosen so that the four
    //      consecutive line elements a warp stores together are contiguous in the peer's memory (256-byte NVLink writes instead
    //      of 64-byte ones: 710 vs 310 GB/s measured) while the pass after the exchange still reads 64-byte rows. ----
    const int cur = pl->p2p_cur;
    pl->p2p_cur ^= 1;
    cx<T>* mine = reinterpret_cast<cx<T>*>(pl->recv[cur]);
    constexpr int B = 64 / (int)sizeof(cx<T>);
    const long long nkt = (nkr + B - 1) / B;           // kx blocks (the last one is ragged)
    const int Tny = (int)ny / 16, Tnz = (int)nz / 16;   // line elements per register step
    long long off[16];
    cx<T>* dst[16];
    if (dir < 0) {
      // dealias-aware exchange: modes the caller zeroes afterwards (aliased kx, aliased y -- GLOBAL range -- and, at the receiver, its
      // aliased local y and aliased kx blocks) are neither stored by the x pass, nor transformed / SENT by the y pass, nor read by the
      // z pass, which writes their zeros directly.  At aliased_fraction = 1/3 the exchange moves 44 % of the bytes.
      const bool skip = fuse && fuse->dealias && env_int("FFB_DEAD_SKIP", 1);
      DeadCols dy = {0, 1 << 30, 0, 0, 0, 0, 0, 0, 0, 0}, dz = dy;
      if (skip) {
        if (fuse->alias_lo[0] > 0) { g_row_dead_lo = fuse->alias_lo[0] - 1; g_row_dead_hi = fuse->alias_hi[0]; dy.dlo = dz.dlo = fuse->alias_lo[0] - 1; dy.dhi = dz.dhi = fuse->alias_hi[0]; }
        dy.on = 1;
        if (fuse->galias_lo[1] > 0) { dy.tlo = fuse->galias_lo[1] - 1; dy.thi = fuse->galias_hi[1]; }
        dz.on = fuse->dealias == 2 ? 1 : 2;
        // the receiver may only skip what every sender skipped: its local aliased y rows need the senders to know the global range
        if (fuse->alias_lo[1] > 0 && fuse->galias_lo[1] > 0) { dz.bylo = fuse->alias_lo[1] - 1; dz.byhi = fuse->alias_hi[1]; }
      }
      rc = pow2_pass<T>(N0, R2C_ROWS, -1, in, w0, 1, N0, 0, 1, nkr, 0, ny * nzl, 1, T(1), tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(),
                        nullptr, 0, nullptr, nullptr, nullptr, rsq);
      g_row_dead_lo = g_row_dead_hi = 0;
      if (rc) return rc;
      // y: w0 (nkr, ny, nzl) -> rank (y / nyl)'s buffer, layout [kt][z][yl][B] with z = rank*nzl + zl
      for (int m = 0; m < 16; ++m) {
        const long long y = (long long)m * Tny;
        off[m] = y * nkr;
        dst[m] = reinterpret_cast<cx<T>*>(pl->peers[cur][y / nyl]) + ((long long)d->rank * nzl * nyl + (y % nyl)) * B;
      }
      { ProfScope ps("fft_y_pass_peer_store", 0);
      g_dead = dy;
      rc = lean_tile_pass<T>((int)ny, -1, B, w0, B, nkr * ny, nkr, off, dst, nz * nyl * B, nyl * B, B, nkr, nzl, T(1), tb1->tw, st);
      g_dead = DeadCols{0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
      if (rc) return rc; }
      if ((rc = dist_barrier(d, st))) return rc;
      // z: own buffer [kt][z][yl][B] -> out (nkr, nyl, nz)
      for (int m = 0; m < 16; ++m) {
        const long long z = (long long)m * Tnz;
        off[m] = z * nyl * B;
        dst[m] = reinterpret_cast<cx<T>*>(out) + z * nkr * nyl;
      }
      g_dead = dz;
      rc = lean_tile_pass<T>((int)nz, -1, B, mine, nz * nyl * B, B, nyl * B, off, dst, B, nkr, nkr * nyl, nkr, nyl, T(1), tb2->tw, st, epi);
      g_dead = DeadCols{0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
      return rc;
    }
    // z: in (nkr, nyl, nz) -> rank (z / nzl)'s buffer, layout [kt][y][zl][B] with y = rank*nyl + yl
    for (int m = 0; m < 16; ++m) {
      const long long z = (long long)m * Tnz;
      off[m] = z * nkr * nyl;
      dst[m] = reinterpret_cast<cx<T>*>(pl->peers[cur][z / nzl]) + ((long long)d->rank * nyl * nzl + (z % nzl)) * B;
    }
    { ProfScope ps("fft_z_pass_peer_store", 0);
    if ((rc = lean_tile_pass<T>((int)nz, +1, B, reinterpret_cast<const cx<T>*>(in), B, nkr, nkr * nyl, off, dst, ny * nzl * B, nzl * B, B, nkr, nyl, T(1), tb2->tw, st))) return rc; }
    if ((rc = dist_barrier(d, st))) return rc;
    // y: own buffer [kt][y][zl][B] -> w2 (nkr, ny, nzl)
    for (int m = 0; m < 16; ++m) {
      const long long y = (long long)m * Tny;
      off[m] = y * nzl * B;
      dst[m] = w2 + y * nkr;
    }
    if ((rc = lean_tile_pass<T>((int)ny, +1, B, mine, ny * nzl * B, B, nzl * B, off, dst, B, nkr * ny, nkr, nkr, nzl, T(1), tb1->tw, st))) return rc;
    (void)nkt;
    return pow2_pass<T>(N0, C2R_ROWS, +1, w2, out, 1, nkr, 0, 1, N0, 0, ny * nzl, 1, inv, tb0->tw, tb0->twr, st);
  }
  if (dir < 0) {
    cx<T>* spec = reinterpret_cast<cx<T>*>(out);
    // x: real (nx, ny, nzl) -> w0 (nkr, ny, nzl)
    if ((rc = pow2_pass<T>(N0, R2C_ROWS, -1, in, w0, 1, N0, 0, 1, nkr, 0, ny * nzl, 1, T(1), tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(),
                           nullptr, 0, nullptr, nullptr, nullptr, rsq))) return rc;
    for (int c = 0; c < nch; ++c) {
      // y on z-chunk c: w0 -> w1 laid out [peer][kx, y_local, z_local]
      if ((rc = pow2_pass<T>((int)ny, C2C_COLS, -1, w0 + c * zc * nkr * ny, w1 + c * sub, nkr, 1, nkr * ny, nkr, 1, nkr * nyl, nkr, zc, T(1), tb1->tw,
                             nullptr, st, SegStride(), seg))) return rc;
      cudaEvent_t e = dist_next_event(d);
      FFB_CUDA(cudaEventRecord(e, st));
      FFB_CUDA(cudaStreamWaitEvent(d->comm_stream, e, 0));
      { ProfScope ps("nccl_alltoall", 0);
      if ((rc = dist_alltoall_bytes(d, w1 + c * sub, spec + c * sub, (size_t)sub * sizeof(cx<T>), (size_t)blk * sizeof(cx<T>), d->comm_stream))) return rc; }
    }
    cudaEvent_t e = dist_next_event(d);
    FFB_CUDA(cudaEventRecord(e, d->comm_stream));
    FFB_CUDA(cudaStreamWaitEvent(st, e, 0));
    // z on the local y-slab, in place: (nkr*nyl) columns of length nz
    if (epi) { ehook.other_from_col = 1; ehook.n0 = (int)nkr; }   // plain layout: column = kx + nkr * y_local
    return pow2_pass<T>((int)nz, C2C_COLS, -1, spec, spec, nkr * nyl, 1, 0, nkr * nyl, 1, 0, nkr * nyl, 1, T(1), tb2->tw, nullptr, st, SegStride(), SegStride(),
                        Outer2(), nullptr, 0, nullptr, epi);
  }
  // inverse
  const cx<T>* spec = reinterpret_cast<const cx<T>*>(in);
  if ((rc = pow2_pass<T>((int)nz, C2C_COLS, +1, spec, w0, nkr * nyl, 1, 0, nkr * nyl, 1, 0, nkr * nyl, 1, T(1), tb2->tw, nullptr, st))) return rc;
  cudaEvent_t e0 = dist_next_event(d);
  FFB_CUDA(cudaEventRecord(e0, st));
  FFB_CUDA(cudaStreamWaitEvent(d->comm_stream, e0, 0));
  for (int c = 0; c < nch; ++c) {
    { ProfScope ps("nccl_alltoall", 0);
    if ((rc = dist_alltoall_bytes(d, w0 + c * sub, w1 + c * sub, (size_t)sub * sizeof(cx<T>), (size_t)blk * sizeof(cx<T>), d->comm_stream))) return rc; }
    cudaEvent_t e = dist_next_event(d);
    FFB_CUDA(cudaEventRecord(e, d->comm_stream));
    FFB_CUDA(cudaStreamWaitEvent(st, e, 0));
    // y on z-chunk c: w1 [peer][kx, y_local, z_local] -> w2 (nkr, ny, zc)
    if ((rc = pow2_pass<T>((int)ny, C2C_COLS, +1, w1 + c * sub, w2 + c * zc * nkr * ny, nkr, 1, nkr * nyl, nkr, 1, nkr * ny, nkr, zc, T(1), tb1->tw, nullptr, st,
                           seg, SegStride()))) return rc;
    // x c2r on z-chunk c
    if ((rc = pow2_pass<T>(N0, C2R_ROWS, +1, w2 + c * zc * nkr * ny, reinterpret_cast<cx<T>*>(out) + c * zc * (long long)N0 * ny, 1, nkr, 0, 1, N0, 0,
                           ny * zc, 1, inv, tb0->tw, tb0->twr, st))) return rc;
  }
  // the next call may overwrite w0 / w1 on the compute stream: all exchanges above have been waited for already
  return FFB_OK;
}

// ---------------------------------------------------------------- slab-decomposed 2-D r2c / c2r ("2-D grids beyond one GPU")
// physical (nx, ny) is split along y: rank r holds (nx, ny/P); the half spectrum is split along kx into blocks of kb = nx/(2P)
// wavenumbers, the Nyquist wavenumber riding with the last block: every rank holds a (kb + 1, ny) array whose last column is the
// Nyquist column on rank P-1 and zero padding elsewhere (equal blocks: the exchange is a plain all-to-all, no pack / unpack).
//   forward: x r2c on the local rows, each line's half spectrum stored block by block in destination-rank-major order
//            -> all-to-all straight into `out` -> y c2c in place on the (kb + 1, ny) slab (four-step for long lines, fusion hooks)
//   inverse: y c2c into scratch -> all-to-all -> x c2r reading the per-rank blocks
template <typename T>
static int exec_dist2d(ffb_plan* pl, const void* in, void* out, int dir, const ffb_fuse* fuse = nullptr) {
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  ffb_dist* d = pl->dist;
  const int P = d->nranks;
  auto* tb0 = reinterpret_cast<DimTables<T>*>(pl->tables[0]);
  auto* tb1 = reinterpret_cast<DimTables<T>*>(pl->tables[1]);
  const long long ny = pl->n[1], nyl = pl->nyl, kb = pl->kb, nkl = kb + 1;
  const int N0 = tb0->N;
  const long long blk = nkl * nyl;               // complex elements exchanged with each peer
  int rc = ensure_ws(pl, 3);
  if (rc) return rc;
  cx<T>* w0 = reinterpret_cast<cx<T>*>(pl->ws[0]);   // forward send buffer: its padding columns are zero and never written
  cx<T>* w1 = reinterpret_cast<cx<T>*>(pl->ws[1]);
  cx<T>* w2 = reinterpret_cast<cx<T>*>(pl->ws[2]);
  if (!pl->ws0_zeroed) { FFB_CUDA(cudaMemsetAsync(w0, 0, pl->ws_bytes, st)); pl->ws0_zeroed = 1; }
  const T inv = (T)(1.0L / ((long double)pl->n[0] * (long double)ny));
  const long long e[3] = {nkl, ny, 1};
  RowSeg rs; rs.seg = (int)kb; rs.stride = blk; rs.nyq = (long long)(P - 1) * blk + kb;
  // y pass over the (nkl, ny) slab: single pass, or the four-step pair through w1 (sub-pass A may run in place, B may not)
  auto y_pass = [&](const cx<T>* src, cx<T>* dst, cx<T>* tmp, int sign, T scale, typename Pow2Params<T>::Fuse* pro, typename Pow2Params<T>::Fuse* epi) -> int {
    if (!tb1->four) return cols_pass<T>(pl, tb1, 0, nkl, 1, src, dst, sign, scale, st, pro, epi);
    if (l2four_enabled(tb1->N1, tb1->N2)) return cols_pass<T>(pl, tb1, 4, nkl, 1, src, dst, sign, scale, st, pro, epi);
    int r = cols_pass<T>(pl, tb1, 1, nkl, 1, src, tmp, sign, T(1), st, pro, nullptr);
    if (r) return r;
    return cols_pass<T>(pl, tb1, 2, nkl, 1, tmp, dst, sign, scale, st, nullptr, epi);
  };
  typename Pow2Params<T>::Fuse hook;
  if (dir < 0) {
    g_row_seg = rs;
    rc = pow2_pass<T>(N0, R2C_ROWS, -1, in, w0, 1, N0, 0, 1, nkl, 0, nyl, 1, T(1), tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(), nullptr, 0,
                      nullptr, nullptr, nullptr, (fuse && fuse->square_input) ? 1 : 0);
    g_row_seg = RowSeg();
    if (rc) return rc;
    cx<T>* spec = reinterpret_cast<cx<T>*>(out);
    const bool four2 = tb1->four && !l2four_enabled(tb1->N1, tb1->N2);
    cx<T>* land = four2 ? w1 : spec;   // the two-kernel four-step needs its input outside `out`
    { ProfScope ps("nccl_alltoall", 0);
    if ((rc = dist_alltoall_bytes(d, w0, land, (size_t)blk * sizeof(cx<T>), (size_t)blk * sizeof(cx<T>), st))) return rc; }
    typename Pow2Params<T>::Fuse* epi = nullptr;
    if (fuse) { hook = make_hook<T>(fuse, 1, 2, e, true); epi = &hook; }
    if (four2) {
      if ((rc = cols_pass<T>(pl, tb1, 1, nkl, 1, w1, w1, -1, T(1), st, nullptr, nullptr))) return rc;   // A in place
      return cols_pass<T>(pl, tb1, 2, nkl, 1, w1, spec, -1, T(1), st, nullptr, epi);
    }
    return y_pass(spec, spec, w1, -1, T(1), nullptr, epi);
  }
  // inverse
  const bool want_pro = fuse && (fuse->kx || fuse->l || fuse->w || fuse->cr != 1.0 || fuse->ci != 0.0);
  typename Pow2Params<T>::Fuse* pro = nullptr;
  if (want_pro) { hook = make_hook<T>(fuse, 1, 2, e, false); pro = &hook; }
  const cx<T>* spec = reinterpret_cast<const cx<T>*>(in);
  if (tb1->four && !l2four_enabled(tb1->N1, tb1->N2)) {
    if ((rc = cols_pass<T>(pl, tb1, 1, nkl, 1, spec, w2, +1, T(1), st, pro, nullptr))) return rc;
    if ((rc = cols_pass<T>(pl, tb1, 2, nkl, 1, w2, w1, +1, T(1), st, nullptr, nullptr))) return rc;
  } else if ((rc = y_pass(spec, w1, w2, +1, T(1), pro, nullptr))) return rc;
  { ProfScope ps("nccl_alltoall", 0);
  if ((rc = dist_alltoall_bytes(d, w1, w2, (size_t)blk * sizeof(cx<T>), (size_t)blk * sizeof(cx<T>), st))) return rc; }
  g_row_seg = rs;
  rc = pow2_pass<T>(N0, C2R_ROWS, +1, w2, out, 1, nkl, 0, 1, N0, 0, nyl, 1, inv, tb0->tw, tb0->twr, st, SegStride(), SegStride(), Outer2(), nullptr, 0,
                    nullptr, nullptr, fuse ? reinterpret_cast<const T*>(fuse->mul) : nullptr);
  g_row_seg = RowSeg();
  return rc;
}

}  // namespace ffb

extern "C" {

int ffb_plan_create(ffb_plan** out, int ndim, const int64_t* n, int dtype, int kind, int nbatch, int flags) {
  FFB_REQUIRE(out && n, FFB_EINVAL, "NULL argument");
  *out = nullptr;
  FFB_REQUIRE(ndim >= 1 && ndim <= 3, FFB_EINVAL, "ndim must be 1, 2 or 3 (got %d)", ndim);
  FFB_REQUIRE(dtype == FFB_F32 || dtype == FFB_F64, FFB_EINVAL, "bad dtype %d", dtype);
  FFB_REQUIRE(kind == FFB_R2C || kind == FFB_C2C, FFB_EINVAL, "bad kind %d", kind);
  FFB_REQUIRE(nbatch >= 1, FFB_EINVAL, "nbatch must be >= 1");
  for (int d = 0; d < ndim; ++d) {
    FFB_REQUIRE(n[d] >= 2, FFB_EINVAL, "n[%d] = %lld too small", d, (long long)n[d]);
    FFB_REQUIRE(n[d] < (1ll << 30), FFB_EUNSUPPORTED, "n[%d] = %lld too large", d, (long long)n[d]);
    // grids require even sizes in every dimension: DomainError, src/domains.jl:66,179,316
    if (n[d] % 2 != 0) return set_error(FFB_EDOMAIN, "n[%d] = %lld must be even", d, (long long)n[d]);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(FFB_ECUDA, "no CUDA device"); }
  auto* pl = new ffb_plan();
  pl->ndim = ndim; pl->dtype = dtype; pl->kind = kind; pl->nbatch = nbatch; pl->flags = flags;
  pl->dist = nullptr; pl->nyl = pl->nzl = 0; pl->nchunks = 1; pl->kb = 0; pl->ws0_zeroed = 0;
  pl->recv[0] = pl->recv[1] = nullptr; pl->p2p = 0; pl->p2p_cur = 0; pl->recv_bytes = 0;
  pl->ring = nullptr; pl->ring_bytes = 0; pl->ctr = nullptr; pl->ctr_count = 0; pl->ctr_C = -1;
  for (int b = 0; b < 2; ++b) for (int q = 0; q < 8; ++q) pl->peers[b][q] = nullptr;
  for (int i = 0; i < 4; ++i) pl->wsm[i] = nullptr;
  for (int d = 0; d < 3; ++d) { pl->n[d] = d < ndim ? n[d] : 1; pl->nc[d] = pl->n[d]; pl->tables[d] = nullptr; pl->ws[d] = nullptr; }
  if (kind == FFB_R2C) pl->nc[0] = pl->n[0] / 2 + 1;
  pl->ws_bytes = (size_t)pl->nc[0] * pl->nc[1] * pl->nc[2] * nbatch * 2 * dtype_size(dtype);
  int rc = dtype == FFB_F64 ? build_tables<double>(pl) : build_tables<float>(pl);
  if (rc) { ffb_plan_destroy(pl); return rc; }
  *out = pl;
  return FFB_OK;
}

int ffb_plan_create_dist(ffb_plan** out, int ndim, const int64_t* n, int dtype, ffb_dist* dist, int nchunks) {
  FFB_REQUIRE(dist, FFB_EINVAL, "dist is NULL");
  FFB_REQUIRE(ndim == 3 || ndim == 2, FFB_EUNSUPPORTED, "slab decomposition is implemented for 2-D and 3-D r2c grids");
  const int P = dist->nranks;
  if (ndim == 2) {
    // physical y-slabs <-> spectral kx-blocks (exec_dist2d)
    FFB_REQUIRE(n[1] % P == 0 && (n[0] / 2) % P == 0, FFB_EUNSUPPORTED, "nx/2 and ny must be divisible by the number of ranks (%d)", P);
    int rc = ffb_plan_create(out, ndim, n, dtype, FFB_R2C, 1, FFB_PLAN_DEFAULT);
    if (rc) return rc;
    ffb_plan* pl = *out;
    bool ok = true;
    if (dtype == FFB_F64) { auto* t0 = reinterpret_cast<DimTables<double>*>(pl->tables[0]); auto* t1 = reinterpret_cast<DimTables<double>*>(pl->tables[1]); ok = t0->tw && t1->pow2 && (t1->four || t1->tw); }
    else { auto* t0 = reinterpret_cast<DimTables<float>*>(pl->tables[0]); auto* t1 = reinterpret_cast<DimTables<float>*>(pl->tables[1]); ok = t0->tw && t1->pow2 && (t1->four || t1->tw); }
    const long long kb = n[0] / 2 / P;
    if (!ok || !is_pow2((uint64_t)kb)) { ffb_plan_destroy(pl); *out = nullptr; return set_error(FFB_EUNSUPPORTED, "slab-decomposed 2-D plans need power-of-two sizes (nx/2/nranks too)"); }
    pl->dist = dist;
    pl->nyl = n[1] / P; pl->nzl = 1; pl->kb = kb;
    pl->ws_bytes = (size_t)(kb + 1) * (size_t)n[1] * 2 * dtype_size(dtype);
    char buf[96];
    snprintf(buf, sizeof(buf), "slab2d[rank %d/%d, kx block %lld + 1] ", dist->rank, P, kb);
    pl->desc += buf;
    return FFB_OK;
  }
  FFB_REQUIRE(n[1] % P == 0 && n[2] % P == 0, FFB_EUNSUPPORTED, "ny and nz must be divisible by the number of ranks (%d)", P);
  int rc = ffb_plan_create(out, ndim, n, dtype, FFB_R2C, 1, FFB_PLAN_DEFAULT);
  if (rc) return rc;
  ffb_plan* pl = *out;
  for (int d = 0; d < 3; ++d) {
    const bool ok = dtype == FFB_F64 ? (reinterpret_cast<DimTables<double>*>(pl->tables[d])->tw != nullptr) : (reinterpret_cast<DimTables<float>*>(pl->tables[d])->tw != nullptr);
    if (!ok) { ffb_plan_destroy(pl); *out = nullptr; return set_error(FFB_EUNSUPPORTED, "slab-decomposed plans need power-of-two sizes within the register-kernel range"); }
  }
  pl->dist = dist;
  pl->nyl = n[1] / P; pl->nzl = n[2] / P;
  if (!is_pow2((uint64_t)pl->nyl)) { ffb_plan_destroy(pl); *out = nullptr; return set_error(FFB_EUNSUPPORTED, "ny / nranks must be a power of two"); }
  int nch = nchunks > 0 ? nchunks : 4;
  while (nch > 1 && (pl->nzl % nch != 0)) nch /= 2;
  pl->nchunks = nch;
  pl->ws_bytes = (size_t)pl->nc[0] * pl->n[1] * pl->nzl * 2 * dtype_size(dtype);
  char buf[96];
  snprintf(buf, sizeof(buf), "slab[rank %d/%d, chunks %d] ", dist->rank, P, nch);
  pl->desc += buf;
  return FFB_OK;
}

// Receive buffers of the peer-memory exchanges.  Both halves live in ONE allocation that is never returned to the driver
// while the process lives: peers keep it mapped (CUDA IPC), so a later plan of the same size reuses it (pool below).
namespace {
struct RecvSlab { void* p; size_t bytes; bool in_use; };
std::vector<RecvSlab> g_recv_pool;
std::mutex g_recv_mu;   // plans may be created / destroyed from different host threads (Julia finalizers)
}  // namespace

int ffb_plan_dist_recv_buffers(ffb_plan* pl, void** buf0, void** buf1, size_t* bytes_each) {
  FFB_REQUIRE(pl && pl->dist && buf0 && buf1, FFB_EINVAL, "needs a slab-decomposed plan");
  FFB_REQUIRE(pl->ndim == 3, FFB_EUNSUPPORTED, "peer-memory exchanges are implemented for 3-D slab plans (2-D plans exchange over NCCL)");
  if (!pl->recv[0]) {
    // room for the kx-padded blocked layout of the peer-store exchange (kx rounded up to whole 64-byte blocks)
    const size_t esz = 2 * dtype_size(pl->dtype);
    const size_t nkrp = ((size_t)pl->nc[0] + 7) / 8 * 8;
    size_t each = nkrp * (size_t)pl->n[1] * (size_t)pl->nzl * esz;
    each = (each + (2u << 20) - 1) / (2u << 20) * (2u << 20);   // whole 2 MiB pages: the allocation is not shared with other buffers
    if (each < (2u << 20)) each = 2u << 20;
    void* slab = nullptr;
    std::lock_guard<std::mutex> lk(g_recv_mu);
    for (auto& r : g_recv_pool)
      if (!r.in_use && r.bytes == 2 * each) { r.in_use = true; slab = r.p; break; }
    if (!slab) {
      int rc = ffb_malloc(&slab, 2 * each);
      if (rc) return rc;
      g_recv_pool.push_back({slab, 2 * each, true});
    }
    pl->recv[0] = slab;
    pl->recv[1] = reinterpret_cast<char*>(slab) + each;
    pl->recv_bytes = each;
  }
  *buf0 = pl->recv[0]; *buf1 = pl->recv[1];
  if (bytes_each) *bytes_each = pl->recv_bytes;
  return FFB_OK;
}

int ffb_plan_dist_set_peers(ffb_plan* pl, void* const* peers0, void* const* peers1) {
  FFB_REQUIRE(pl && pl->dist && peers0 && peers1, FFB_EINVAL, "needs a slab-decomposed plan");
  FFB_REQUIRE(pl->recv[0] && pl->recv[1], FFB_EINVAL, "call ffb_plan_dist_recv_buffers first");
  const int P = pl->dist->nranks;
  FFB_REQUIRE(P <= 8, FFB_EUNSUPPORTED, "peer-store exchange supports up to 8 ranks (one NVSwitch domain)");
  FFB_REQUIRE(is_pow2((uint64_t)pl->nzl), FFB_EUNSUPPORTED, "nz / nranks must be a power of two");
  for (int q = 0; q < P; ++q) {
    FFB_REQUIRE(peers0[q] && peers1[q], FFB_EINVAL, "peer pointer %d is NULL", q);
    pl->peers[0][q] = peers0[q]; pl->peers[1][q] = peers1[q];
  }
  FFB_REQUIRE(pl->peers[0][pl->dist->rank] == pl->recv[0] && pl->peers[1][pl->dist->rank] == pl->recv[1], FFB_EINVAL,
              "the entry of the own rank must be the local receive buffer");
  return FFB_OK;
}

int ffb_plan_dist_set_exchange(ffb_plan* pl, int mode) {
  FFB_REQUIRE(pl && pl->dist, FFB_EINVAL, "needs a slab-decomposed plan");
  FFB_REQUIRE(mode == FFB_EXCHANGE_NCCL || mode == FFB_EXCHANGE_PEER_STORE || mode == FFB_EXCHANGE_COPY_ENGINE, FFB_EINVAL, "bad exchange mode %d", mode);
  FFB_REQUIRE(mode == FFB_EXCHANGE_NCCL || pl->peers[0][0], FFB_EINVAL, "call ffb_plan_dist_set_peers first");
  FFB_REQUIRE(pl->nchunks <= 8, FFB_EUNSUPPORTED, "too many chunks");
  if (mode == FFB_EXCHANGE_PEER_STORE) {
    // the blocked passes use 64-byte wide tiles: line length * tile width must fit one CTA
    const int maxn = 16 * pow2_max_threads(dtype_size(pl->dtype)) / (64 / (2 * (int)dtype_size(pl->dtype)));
    FFB_REQUIRE(pl->n[1] >= 16 && pl->n[2] >= 16 && pl->n[1] <= maxn && pl->n[2] <= maxn, FFB_EUNSUPPORTED,
                "peer-store exchange needs 16 <= ny, nz <= %d for this precision", maxn);
    FFB_REQUIRE(pl->nyl % (pl->n[1] / 16) == 0 && pl->nzl % (pl->n[2] / 16) == 0, FFB_EUNSUPPORTED, "peer-store exchange needs at most 16 ranks");
  }
  pl->p2p = mode;
  if (mode == FFB_EXCHANGE_PEER_STORE && (pl->ws[1] || pl->ws[2])) {
    // the peer-store exchange needs one scratch array: give back what an earlier exchange (e.g. the plan-time measurement) allocated
    cudaStream_t st = current_stream();
    if (st) FFB_CUDA(cudaStreamSynchronize(st));
    FFB_CUDA(cudaStreamSynchronize(pl->dist->comm_stream));
    for (int i = 1; i < 3; ++i) { if (pl->ws[i]) cudaFree(pl->ws[i]); pl->ws[i] = nullptr; }
  }
  static const char* names[3] = {"exchange=nccl ", "exchange=peer-store ", "exchange=copy-engine "};
  const size_t prev = pl->desc.find("exchange=");
  if (prev != std::string::npos) pl->desc.erase(prev);   // the description names the exchange in use, not the history
  pl->desc += names[mode];
  return FFB_OK;
}

int ffb_plan_dist_get_exchange(const ffb_plan* pl, int* mode) {
  FFB_REQUIRE(pl && pl->dist && mode, FFB_EINVAL, "needs a slab-decomposed plan");
  *mode = pl->p2p;
  return FFB_OK;
}

int ffb_plan_destroy(ffb_plan* pl) {
  if (!pl) return FFB_OK;
  if (pl->dtype == FFB_F64) free_tables<double>(pl); else free_tables<float>(pl);
  for (int i = 0; i < 3; ++i) cudaFree(pl->ws[i]);
  for (int i = 0; i < 4; ++i) cudaFree(pl->wsm[i]);
  cudaFree(pl->ring); cudaFree(pl->ctr);
  if (pl->recv[0]) {
    std::lock_guard<std::mutex> lk(g_recv_mu);
    for (auto& r : g_recv_pool)
      if (r.p == pl->recv[0]) r.in_use = false;   // stays allocated: peers may still have it mapped
  }
  delete pl;
  return FFB_OK;
}

int ffb_plan_workspace_bytes(const ffb_plan* pl, size_t* bytes) {
  FFB_REQUIRE(pl && bytes, FFB_EINVAL, "NULL argument");
  size_t b = 0;
  for (int i = 0; i < 3; ++i) if (pl->ws[i]) b += pl->ws_bytes;
  for (int i = 0; i < 4; ++i) if (pl->wsm[i]) b += pl->ws_bytes;
  *bytes = b;
  return FFB_OK;
}

int ffb_plan_describe(const ffb_plan* pl, char* buf, size_t buflen) {
  FFB_REQUIRE(pl && buf && buflen, FFB_EINVAL, "NULL argument");
  snprintf(buf, buflen, "%s", pl->desc.c_str());
  return FFB_OK;
}

int ffb_fft_forward(ffb_plan* pl, const void* in, void* out) {
  FFB_REQUIRE(pl && in && out, FFB_EINVAL, "NULL argument");
  if (pl->kind == FFB_R2C) FFB_REQUIRE(in != out, FFB_EINVAL, "r2c transforms are out of place");
  if (pl->dist && pl->ndim == 2) return pl->dtype == FFB_F64 ? exec_dist2d<double>(pl, in, out, -1) : exec_dist2d<float>(pl, in, out, -1);
  if (pl->dist) return pl->dtype == FFB_F64 ? exec_dist<double>(pl, in, out, -1) : exec_dist<float>(pl, in, out, -1);
  return pl->dtype == FFB_F64 ? exec<double>(pl, in, out, -1) : exec<float>(pl, in, out, -1);
}

static int exec_fused(ffb_plan* pl, const void* in, void* out, int dir, const ffb_fuse* fuse) {
  FFB_REQUIRE(pl && in && out && fuse, FFB_EINVAL, "NULL argument");
  FFB_REQUIRE(in != out, FFB_EINVAL, "fused transforms are out of place");
  if (pl->dist && pl->ndim == 2) return pl->dtype == FFB_F64 ? exec_dist2d<double>(pl, in, out, dir, fuse) : exec_dist2d<float>(pl, in, out, dir, fuse);
  if (pl->dist) return pl->dtype == FFB_F64 ? exec_dist<double>(pl, in, out, dir, fuse) : exec_dist<float>(pl, in, out, dir, fuse);
  bool allp = true;
  for (int d = 0; d < pl->ndim; ++d) {
    if (pl->dtype == FFB_F64) { auto* tb = reinterpret_cast<DimTables<double>*>(pl->tables[d]); allp = allp && tb->pow2 && (tb->four || tb->tw); }
    else { auto* tb = reinterpret_cast<DimTables<float>*>(pl->tables[d]); allp = allp && tb->pow2 && (tb->four || tb->tw); }
  }
  FFB_REQUIRE(allp, FFB_EUNSUPPORTED, "fused transforms need power-of-two sizes");
  return pl->dtype == FFB_F64 ? exec_pow2<double>(pl, in, out, dir, fuse) : exec_pow2<float>(pl, in, out, dir, fuse);
}

int ffb_fft_forward_ex(ffb_plan* pl, const void* in, void* out, const ffb_fuse* fuse) { return exec_fused(pl, in, out, -1, fuse); }
int ffb_fft_inverse_ex(ffb_plan* pl, const void* in, void* out, const ffb_fuse* fuse) { return exec_fused(pl, in, out, +1, fuse); }

int ffb_fft_inverse_multi(ffb_plan* pl, const void* in, int n, void* const* outs, const ffb_fuse* fuses) {
  FFB_REQUIRE(pl && in && outs && fuses, FFB_EINVAL, "NULL argument");
  FFB_REQUIRE(n >= 1, FFB_EINVAL, "n = %d", n);
  for (int v = 0; v < n; ++v) FFB_REQUIRE(outs[v] && outs[v] != in, FFB_EINVAL, "fused transforms are out of place");
  bool allp = !pl->dist;
  for (int d = 0; allp && d < pl->ndim; ++d) {
    if (pl->dtype == FFB_F64) { auto* tb = reinterpret_cast<DimTables<double>*>(pl->tables[d]); allp = tb->pow2 && (tb->four || tb->tw); }
    else { auto* tb = reinterpret_cast<DimTables<float>*>(pl->tables[d]); allp = tb->pow2 && (tb->four || tb->tw); }
  }
  if (!allp) {   // slab-decomposed or arbitrary-size plans: one fused inverse transform after the other
    for (int v = 0; v < n; ++v) {
      int rc = ffb_fft_inverse_ex(pl, in, outs[v], &fuses[v]);
      if (rc) return rc;
    }
    return FFB_OK;
  }
  return pl->dtype == FFB_F64 ? exec_pow2_multi<double>(pl, in, n, outs, fuses) : exec_pow2_multi<float>(pl, in, n, outs, fuses);
}

int ffb_fft_inverse(ffb_plan* pl, const void* in, void* out) {
  FFB_REQUIRE(pl && in && out, FFB_EINVAL, "NULL argument");
  if (pl->kind == FFB_R2C) FFB_REQUIRE(in != out, FFB_EINVAL, "c2r transforms are out of place");
  if (pl->dist && pl->ndim == 2) return pl->dtype == FFB_F64 ? exec_dist2d<double>(pl, in, out, +1) : exec_dist2d<float>(pl, in, out, +1);
  if (pl->dist) return pl->dtype == FFB_F64 ? exec_dist<double>(pl, in, out, +1) : exec_dist<float>(pl, in, out, +1);
  return pl->dtype == FFB_F64 ? exec<double>(pl, in, out, +1) : exec<float>(pl, in, out, +1);
}

}  // extern "C"
