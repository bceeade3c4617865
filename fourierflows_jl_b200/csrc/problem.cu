// C-driven problem: `Problem(eqn, stepper, dt, grid, vars, params)` (src/problem.jl:99-111), `stepforward!`
// (src/timesteppers.jl:6-35, one method per stepper :111-667) and `step_until!` (:734-760), with the calcN!
// implementations of the benchmark equations built from library kernels.  The whole step is enqueued on the
// library stream without host synchronisation.
#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>
#include "ffb_common.cuh"
#include "dist.h"

namespace ffb {

// ---------------------------------------------------------------- fused calcN! kernels (2-D vorticity)
// uh = im*l*invKrsq*sol ; vh = -im*kr*invKrsq*sol   (GeophysicalFlows TwoDNavierStokes.calcN_advection!)
template <typename T>
__global__ void vort_prep_kernel(cx<T>* uh, cx<T>* vh, const cx<T>* sol, const T* invK, const T* kr, const T* l, long long n0) {
  const long long row = blockIdx.x;
  const T lv = l[row];
  for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
    const long long idx = row * n0 + i;
    const cx<T> s = sol[idx];
    const T w = invK[idx];
    const T lw = lv * w, kw = kr[i] * w;
    uh[idx] = mk<T>(-(lw * s.y), lw * s.x);
    vh[idx] = mk<T>(kw * s.y, -(kw * s.x));
  }
}
// u *= zeta ; v *= zeta
template <typename T>
__global__ void mul2_real_kernel(T* u, T* v, const T* z, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T zz = z[i];
    u[i] = u[i] * zz;
    v[i] = v[i] * zz;
  }
}
// N = -im*kr*uh - im*l*vh ; dealias!(N, grid)
template <typename T>
__global__ void vort_combine_kernel(cx<T>* N, const cx<T>* uh, const cx<T>* vh, const T* kr, const T* l, long long n0, int lo0, int hi0,
                                    int lo1, int hi1) {
  const long long row = blockIdx.x;
  const bool whole = lo1 > 0 && row >= lo1 - 1 && row < hi1;
  const T lv = l[row];
  for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
    const long long idx = row * n0 + i;
    if (whole || (lo0 > 0 && i >= lo0 - 1 && i < hi0)) { N[idx] = mk<T>(0, 0); continue; }
    const cx<T> a = uh[idx], b = vh[idx];
    const T k = kr[i];
    N[idx] = mk<T>(k * a.y + lv * b.y, -(k * a.x) - lv * b.x);
  }
}
// c = c^2 on 16-byte vectors (n is a multiple of 4: nx is even and so is every other extent)
template <typename T>
__global__ void square_real_kernel(T* c, long long n) {
  constexpr int V = 16 / sizeof(T);
  using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
  const long long nv = n / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  V4* cv = reinterpret_cast<V4*>(c);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    V4 v = cv[i];
    if constexpr (sizeof(T) == 4) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
    else { v.x *= v.x; v.y *= v.y; }
    cv[i] = v;
  }
  for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const T v = c[i]; c[i] = v * v; }
}

}  // namespace ffb

using namespace ffb;

struct ffb_problem {
  ffb_problem_config cfg;
  ffb_desc desc;
  int32_t galias_lo[3], galias_hi[3];   // global alias ranges (desc holds this rank's local ones on slab-decomposed problems)
  int nd, dtype;
  long long n[3], nkr, nspec, nphys;
  size_t sbytes, pbytes, rbytes;  // complex spectral array, real physical array, real spectral-shaped array
  void *kr, *l, *m, *Krsq, *invKrsq, *Ldense, *filter;
  ffb_coef L, cE, cE2, cz, ca, cb, cg;
  void *E, *E2, *zeta, *alpha, *beta, *gamma;
  ffb_plan* plan;
  void* sol;
  void* arr[6];            // stepper arrays
  void *sh1, *sh2;         // spectral scratch (uh, vh / cxh)
  void *ph1, *ph2, *ph3;   // physical scratch (u, v, zeta / cx)
  double t, dt;
  int64_t step;
  size_t device_bytes;
  std::vector<void*> owned;
};

namespace ffb {

static double roundT(const ffb_problem* p, double v) { return p->dtype == FFB_F32 ? (double)(float)v : v; }

static int dalloc(ffb_problem* p, void** ptr, size_t bytes, bool zero) {
  int rc = ffb_malloc(ptr, bytes);
  if (rc) return rc;
  p->owned.push_back(*ptr);
  p->device_bytes += bytes;
  if (zero) return ffb_memset_zero(*ptr, bytes);
  return FFB_OK;
}

template <typename T>
static int calcN_impl(ffb_problem* p, void* N, const void* sol, double t) {
  cudaStream_t s = current_stream();
  const ffb_desc* d = &p->desc;
  const long long n0 = d->dims[0];
  int rc;
  switch (p->cfg.calcN) {
    case FFB_CALCN_CALLBACK:
      FFB_REQUIRE(p->cfg.callback, FFB_EINVAL, "calcN callback is NULL");
      return p->cfg.callback(N, sol, t, p->cfg.user);
    case FFB_CALCN_ZERO:
      return ffb_memset_zero(N, p->sbytes);
    case FFB_CALCN_DIFFUSION:
      // src/diffusion.jl:135-143
      if ((rc = ffb_ew_spectral_mul(p->sh1, sol, 0.0, 1.0, p->kr, 1, nullptr, 0, nullptr, 0, nullptr, 0, 0, d))) return rc;
      if ((rc = ffb_fft_inverse(p->plan, p->sh1, p->ph1))) return rc;
      if ((rc = ffb_ew_mul_real(p->ph1, p->ph1, p->cfg.kappa, p->dtype, p->nphys))) return rc;
      if ((rc = ffb_fft_forward(p->plan, p->ph1, p->sh1))) return rc;
      return ffb_ew_spectral_mul(N, p->sh1, 0.0, 1.0, p->kr, 1, nullptr, 0, nullptr, 0, nullptr, 0, 0, d);
    case FFB_CALCN_VORTICITY2D: {
      if (p->cfg.fused) {
        // Same equation with the elementwise work folded into the transforms (5 transforms, no separate kernels):
        //   zeta = irfft(sol);  u*zeta = irfft(im*l*invKrsq .* sol) .* zeta;  v*zeta = irfft(-im*kr*invKrsq .* sol) .* zeta
        //   uh = rfft(u*zeta);  N = dealias!(-im*kr .* uh - im*l .* rfft(v*zeta))
        ffb_fuse f;
        if (p->cfg.dist) {
          memset(&f, 0, sizeof(f));
          f.cr = 1.0;
          if ((rc = ffb_fft_inverse(p->plan, sol, p->ph3))) return rc;
          f.cr = 0.0; f.ci = 1.0; f.l = p->l; f.w = p->invKrsq; f.mul = p->ph3;
          if ((rc = ffb_fft_inverse_ex(p->plan, sol, p->ph1, &f))) return rc;
          f.ci = -1.0; f.l = nullptr; f.kx = p->kr;
          if ((rc = ffb_fft_inverse_ex(p->plan, sol, p->ph2, &f))) return rc;
        } else {
          // the three inverse transforms share their input: one call, `sol` and invKrsq are read from DRAM once by the first sub-pass
          ffb_fuse g[3];
          memset(g, 0, sizeof(g));
          g[0].cr = 1.0;
          g[1].ci = 1.0; g[1].l = p->l; g[1].w = p->invKrsq; g[1].mul = p->ph3;
          g[2].ci = -1.0; g[2].kx = p->kr; g[2].w = p->invKrsq; g[2].mul = p->ph3;
          void* outs[3] = {p->ph3, p->ph1, p->ph2};
          if ((rc = ffb_fft_inverse_multi(p->plan, sol, 3, outs, g))) return rc;
        }
        // uh = rfft(u*zeta) is only ever read as the accumulated operand of the dealiased transform below, which never loads it
        // inside the alias box: that box is don't-care here (dealias = 2), its columns are neither stored nor transformed
        memset(&f, 0, sizeof(f));
        f.cr = 1.0; f.dealias = d->alias_lo[0] > 0 ? 2 : 0;
        for (int q = 0; q < 3; ++q) { f.alias_lo[q] = d->alias_lo[q]; f.alias_hi[q] = d->alias_hi[q]; }
        if ((rc = ffb_fft_forward_ex(p->plan, p->ph1, p->sh1, &f))) return rc;
        memset(&f, 0, sizeof(f));
        f.cr = 0.0; f.ci = -1.0; f.l = p->l;                   // own term: -im * l * rfft(v*zeta)
        f.acc = p->sh1; f.ar = 0.0; f.ai = -1.0; f.akx = p->kr;   // accumulated term: -im * kr * uh
        f.dealias = 1;
        for (int q = 0; q < 3; ++q) { f.alias_lo[q] = d->alias_lo[q]; f.alias_hi[q] = d->alias_hi[q]; }
        return ffb_fft_forward_ex(p->plan, p->ph2, N, &f);
      }
      const unsigned rows = (unsigned)d->dims[1];
      { ProfScope ps("calcN_vort_prep", 3.0 * p->sbytes + p->rbytes);
      vort_prep_kernel<T><<<rows, 256, 0, s>>>((cx<T>*)p->sh1, (cx<T>*)p->sh2, (const cx<T>*)sol, (const T*)p->invKrsq, (const T*)p->kr, (const T*)p->l, n0);
      count_launch(); }
      FFB_CHECK_LAUNCH();
      if ((rc = ffb_fft_inverse(p->plan, p->sh1, p->ph1))) return rc;   // u
      if ((rc = ffb_fft_inverse(p->plan, p->sh2, p->ph2))) return rc;   // v
      if ((rc = ffb_fft_inverse(p->plan, sol, p->ph3))) return rc;      // zeta (zeta_h = sol; the transform preserves its input)
      const unsigned blocks = (unsigned)std::min<long long>((p->nphys + 255) / 256, (long long)num_sms() * 16);
      { ProfScope ps("calcN_vort_products", 5.0 * p->pbytes);
      mul2_real_kernel<T><<<blocks, 256, 0, s>>>((T*)p->ph1, (T*)p->ph2, (const T*)p->ph3, p->nphys);
      count_launch(); }
      FFB_CHECK_LAUNCH();
      if ((rc = ffb_fft_forward(p->plan, p->ph1, p->sh1))) return rc;
      if ((rc = ffb_fft_forward(p->plan, p->ph2, p->sh2))) return rc;
      { ProfScope ps("calcN_vort_combine", 3.0 * p->sbytes);
      vort_combine_kernel<T><<<rows, 256, 0, s>>>((cx<T>*)N, (const cx<T>*)p->sh1, (const cx<T>*)p->sh2, (const T*)p->kr, (const T*)p->l, n0,
                                                  d->alias_lo[0], d->alias_hi[0], d->alias_lo[1], d->alias_hi[1]);
      count_launch(); }
      FFB_CHECK_LAUNCH();
      return FFB_OK;
    }
    case FFB_CALCN_BURGERS3D: {
      // N = -1/2 im kr rfft(irfft(sol)^2) ; dealias!(N, grid)   (builder-defined 3-D test equation)
      if ((rc = ffb_fft_inverse(p->plan, sol, p->ph1))) return rc;
      int exch = 0;
      if (p->cfg.dist) ffb_plan_dist_get_exchange(p->plan, &exch);
      if (p->cfg.fused && exch != FFB_EXCHANGE_COPY_ENGINE) {
        // the square is folded into the x pass (r2c load), `-1/2 im kr` and the dealias mask into the last pass's store:
        // no elementwise kernel at all, also on slab-decomposed plans
        ffb_fuse f;
        memset(&f, 0, sizeof(f));
        f.cr = 0.0; f.ci = -0.5; f.kx = p->kr; f.dealias = 1; f.square_input = 1;
        for (int q = 0; q < 3; ++q) { f.alias_lo[q] = d->alias_lo[q]; f.alias_hi[q] = d->alias_hi[q]; f.galias_lo[q] = p->galias_lo[q]; f.galias_hi[q] = p->galias_hi[q]; }
        return ffb_fft_forward_ex(p->plan, p->ph1, N, &f);
      }
      const unsigned blocks = (unsigned)std::min<long long>((p->nphys / 4 + 255) / 256, (long long)num_sms() * 16);
      { ProfScope ps("calcN_square", 2.0 * p->pbytes);
      square_real_kernel<T><<<blocks, 256, 0, s>>>((T*)p->ph1, p->nphys);
      count_launch(); }
      FFB_CHECK_LAUNCH();
      if (!p->sh1 && (rc = dalloc(p, &p->sh1, p->sbytes, true))) return rc;
      if ((rc = ffb_fft_forward(p->plan, p->ph1, p->sh1))) return rc;
      return ffb_ew_spectral_mul(N, p->sh1, 0.0, -0.5, p->kr, 1, nullptr, 0, nullptr, 0, nullptr, 0, 1, d);
    }
  }
  return set_error(FFB_EINVAL, "bad calcN kind %d", p->cfg.calcN);
}

static int calcN(ffb_problem* p, void* N, const void* sol, double t) {
  return p->dtype == FFB_F64 ? calcN_impl<double>(p, N, sol, t) : calcN_impl<float>(p, N, sol, t);
}

// LSRK54 tableau (src/timesteppers.jl:335-350), rationals evaluated in Float64 then converted to T by the stage
static const double kLsrkA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                 -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
static const double kLsrkB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
                                 3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0};
static const double kLsrkC[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                                 2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};

static int step_once(ffb_problem* p) {
  const int dt_ = p->dtype;
  const int64_t n = p->nspec;
  const double t = p->t, dt = p->dt;
  const void* filt = p->cfg.filtered ? p->filter : nullptr;
  int rc;
  switch (p->cfg.stepper) {
    case FFB_FORWARD_EULER:
      if ((rc = calcN(p, p->arr[0], p->sol, t))) return rc;
      if ((rc = ffb_stage_fe(p->sol, p->arr[0], &p->L, dt, filt, dt_, n))) return rc;
      break;
    case FFB_RK4: {
      void *sol1 = p->arr[0], *r1 = p->arr[1], *r2 = p->arr[2], *r3 = p->arr[3], *r4 = p->arr[4];
      const double h = roundT(p, dt / 2);
      if ((rc = calcN(p, r1, p->sol, t))) return rc;
      if ((rc = ffb_stage_rk4_substep(sol1, r1, p->sol, p->sol, &p->L, h, dt_, n))) return rc;
      if ((rc = calcN(p, r2, sol1, roundT(p, t + h)))) return rc;
      if ((rc = ffb_stage_rk4_substep(sol1, r2, sol1, p->sol, &p->L, h, dt_, n))) return rc;
      if ((rc = calcN(p, r3, sol1, roundT(p, t + h)))) return rc;
      if ((rc = ffb_stage_rk4_substep(sol1, r3, sol1, p->sol, &p->L, dt, dt_, n))) return rc;
      if ((rc = calcN(p, r4, sol1, roundT(p, t + dt)))) return rc;
      if ((rc = ffb_stage_rk4_final(p->sol, r1, r2, r3, r4, sol1, &p->L, dt, filt, 0, dt_, n))) return rc;
      break;
    }
    case FFB_LSRK54: {
      void *S2 = p->arr[0], *rhs = p->arr[1];
      for (int i = 0; i < 5; ++i) {
        const double ci = roundT(p, kLsrkC[i]);
        if ((rc = calcN(p, rhs, p->sol, roundT(p, t + roundT(p, ci * dt))))) return rc;
        if ((rc = ffb_stage_lsrk54(p->sol, S2, rhs, &p->L, kLsrkA[i], kLsrkB[i], dt, i == 0, i == 4 ? filt : nullptr, dt_, n))) return rc;
      }
      break;
    }
    case FFB_ETDRK4: {
      void *sol1 = p->arr[0], *sol2 = p->arr[1], *N1 = p->arr[2], *N2 = p->arr[3], *N3 = p->arr[4], *N4 = p->arr[5];
      if ((rc = calcN(p, N1, p->sol, t))) return rc;
      if ((rc = ffb_stage_etdrk4_substep12(sol1, &p->cE2, p->sol, &p->cz, N1, dt_, n))) return rc;
      const double t2 = roundT(p, t + roundT(p, dt / 2));
      if ((rc = calcN(p, N2, sol1, t2))) return rc;
      if ((rc = ffb_stage_etdrk4_substep12(sol2, &p->cE2, p->sol, &p->cz, N2, dt_, n))) return rc;
      if ((rc = calcN(p, N3, sol2, t2))) return rc;
      if ((rc = ffb_stage_etdrk4_substep3(sol2, &p->cE2, sol1, &p->cz, N1, N3, dt_, n))) return rc;
      if ((rc = calcN(p, N4, sol2, roundT(p, t + dt)))) return rc;
      if ((rc = ffb_stage_etdrk4_update(p->sol, &p->cE, &p->ca, &p->cb, &p->cg, N1, N2, N3, N4, filt, dt_, n))) return rc;
      break;
    }
    case FFB_AB3: {
      void *rhs = p->arr[0], *m1 = p->arr[1], *m2 = p->arr[2];
      if ((rc = calcN(p, rhs, p->sol, t))) return rc;
      if ((rc = ffb_stage_ab3(p->sol, rhs, m1, m2, &p->L, dt, p->step, filt, dt_, n))) return rc;
      // RHS_2 <- RHS_1 ; RHS_1 <- RHS (after the clock tick, :644-648): rotate the pointers instead of copying
      p->arr[0] = m2; p->arr[1] = rhs; p->arr[2] = m1;
      break;
    }
    default:
      return set_error(FFB_EINVAL, "bad stepper %d", p->cfg.stepper);
  }
  p->t = roundT(p, p->t + p->dt);
  p->step += 1;
  return FFB_OK;
}

}  // namespace ffb

extern "C" {

// `jacobianh(a, b, grid)`, real fields (src/utils.jl:190-197), from fused transforms only
int ffb_jacobianh(ffb_plan* plan, void* out, const void* a, const void* b, const void* kr, const void* l, void* sh, void* p1, void* p2) {
  FFB_REQUIRE(plan && out && a && b && kr && l && sh && p1 && p2, FFB_EINVAL, "NULL argument");
  int rc;
  if ((rc = ffb_fft_forward(plan, b, sh))) return rc;                       // bh = rfft(b)
  ffb_fuse f;
  memset(&f, 0, sizeof(f));
  f.cr = 0.0; f.ci = 1.0; f.l = l; f.mul = a;
  if ((rc = ffb_fft_inverse_ex(plan, sh, p1, &f))) return rc;               // p1 = a .* irfft(im * l .* bh)   (= a .* by)
  f.l = nullptr; f.kx = kr;
  if ((rc = ffb_fft_inverse_ex(plan, sh, p2, &f))) return rc;               // p2 = a .* irfft(im * kr .* bh)  (= a .* bx)
  if ((rc = ffb_fft_forward(plan, p1, sh))) return rc;                      // sh = rfft(a .* by)
  memset(&f, 0, sizeof(f));
  f.cr = 0.0; f.ci = -1.0; f.l = l;                                          // own term: -im * l .* rfft(a .* bx)
  f.acc = sh; f.ar = 0.0; f.ai = 1.0; f.akx = kr;                            // accumulated term: +im * kr .* rfft(a .* by)
  return ffb_fft_forward_ex(plan, p2, out, &f);
}

int ffb_problem_destroy(ffb_problem* p) {
  if (!p) return FFB_OK;
  ffb_sync();
  for (void* q : p->owned) cudaFree(q);
  if (p->plan) ffb_plan_destroy(p->plan);
  delete p;
  return FFB_OK;
}

int ffb_problem_create(ffb_problem** out, const ffb_problem_config* cfg) {
  FFB_REQUIRE(out && cfg, FFB_EINVAL, "NULL argument");
  *out = nullptr;
  FFB_REQUIRE(cfg->ndim >= 1 && cfg->ndim <= 3, FFB_EINVAL, "ndim must be 1..3");
  FFB_REQUIRE(cfg->dtype == FFB_F32 || cfg->dtype == FFB_F64, FFB_EINVAL, "bad dtype");
  FFB_REQUIRE(cfg->stepper >= FFB_FORWARD_EULER && cfg->stepper <= FFB_AB3, FFB_EINVAL, "bad stepper");
  FFB_REQUIRE(cfg->aliased_fraction >= 0 && cfg->aliased_fraction < 1, FFB_EINVAL, "`aliased_fraction` must be in [0, 1)");
  if (cfg->calcN == FFB_CALCN_VORTICITY2D) FFB_REQUIRE(cfg->ndim == 2, FFB_EINVAL, "vorticity equation needs a 2-D grid");
  if (cfg->calcN == FFB_CALCN_DIFFUSION) FFB_REQUIRE(cfg->ndim == 1 && cfg->kappa, FFB_EINVAL, "array-kappa diffusion needs a 1-D grid and kappa");
  for (int d = 0; d < cfg->ndim; ++d)
    if (cfg->n[d] % 2 != 0) return set_error(FFB_EDOMAIN, "n[%d] = %lld must be even", d, (long long)cfg->n[d]);

  auto* p = new ffb_problem();
  p->cfg = *cfg;
  p->nd = cfg->ndim; p->dtype = cfg->dtype;
  p->plan = nullptr; p->device_bytes = 0;
  const size_t es = dtype_size(cfg->dtype);
  for (int d = 0; d < 3; ++d) p->n[d] = d < cfg->ndim ? cfg->n[d] : 1;
  p->nkr = p->n[0] / 2 + 1;
  // slab decomposition: spectral (nkr, ny/P, nz), physical (nx, ny, nz/P)
  const int P = cfg->dist ? cfg->dist->nranks : 1, rank = cfg->dist ? cfg->dist->rank : 0;
  const bool dist2d = cfg->dist && cfg->ndim == 2;
  if (cfg->dist) {
    const bool ok3 = cfg->ndim == 3 && (cfg->calcN == FFB_CALCN_BURGERS3D || cfg->calcN == FFB_CALCN_ZERO || cfg->calcN == FFB_CALCN_CALLBACK);
    const bool ok2 = cfg->ndim == 2 && (cfg->calcN == FFB_CALCN_VORTICITY2D || cfg->calcN == FFB_CALCN_ZERO || cfg->calcN == FFB_CALCN_CALLBACK);
    if (!ok3 && !ok2) {
      delete p;
      return set_error(FFB_EUNSUPPORTED, "slab-decomposed problems: 3-D grids with the Burgers / zero / callback calcN, 2-D grids with the vorticity / zero / callback calcN");
    }
    if (cfg->ndim == 3 && (p->n[1] % P || p->n[2] % P)) { delete p; return set_error(FFB_EUNSUPPORTED, "ny and nz must be divisible by the number of ranks"); }
    if (cfg->ndim == 2 && (p->n[1] % P || (p->n[0] / 2) % P)) { delete p; return set_error(FFB_EUNSUPPORTED, "nx/2 and ny must be divisible by the number of ranks"); }
  }
  // 3-D: spectral (nkr, ny/P, nz), physical (nx, ny, nz/P).  2-D: spectral (kb + 1, ny) with kb = nx/(2P) -- the Nyquist column rides with
  // the last rank, the extra column is zero padding elsewhere --, physical (nx, ny/P)
  const long long kb = dist2d ? p->n[0] / 2 / P : 0;
  const long long nk0 = dist2d ? kb + 1 : p->nkr;
  const long long nyl = dist2d ? p->n[1] : p->n[1] / ((cfg->dist && !dist2d) ? P : 1);
  const long long nzl = p->n[2] / ((cfg->dist && !dist2d) ? P : 1);
  p->nspec = nk0 * nyl * p->n[2];
  p->nphys = dist2d ? p->n[0] * (p->n[1] / P) : p->n[0] * p->n[1] * nzl;
  p->sbytes = (size_t)p->nspec * 2 * es; p->pbytes = (size_t)p->nphys * es; p->rbytes = (size_t)p->nspec * es;
  ffb_desc& D = p->desc;
  D.ndim = p->nd; D.dtype = p->dtype;
  D.dims[0] = nk0; D.dims[1] = nyl; D.dims[2] = p->n[2]; D.dims[3] = 1;
  // getaliasedwavenumbers (src/domains.jl:408-421), evaluated in Float64
  for (int d = 0; d < 3; ++d) { D.alias_lo[d] = 0; D.alias_hi[d] = 0; }
  if (cfg->aliased_fraction > 0) {
    const double Lf = (1 - cfg->aliased_fraction) / 2, Rf = (1 + cfg->aliased_fraction) / 2;
    for (int d = 0; d < p->nd; ++d) {
      D.alias_lo[d] = (int32_t)std::floor(Lf * (double)p->n[d]) + 1;
      D.alias_hi[d] = d == 0 ? (int32_t)p->nkr : (int32_t)std::ceil(Rf * (double)p->n[d]);  // kralias = iL:nkr for the half spectrum
    }
  }
  for (int d = 0; d < 3; ++d) { p->galias_lo[d] = D.alias_lo[d]; p->galias_hi[d] = D.alias_hi[d]; }
  if (dist2d && D.alias_lo[0] > 0) {
    // kralias = iL:nkr intersected with this rank's kx block [rank*kb + 1, (rank+1)*kb], in local 1-based indices; the extra column
    // (Nyquist on the last rank, zero padding elsewhere) is always inside the aliased range
    const long long lo = std::max<long long>(D.alias_lo[0], rank * kb + 1);
    D.alias_lo[0] = (int32_t)(lo <= (rank + 1) * kb ? lo - rank * kb : kb + 1);
    D.alias_hi[0] = (int32_t)(kb + 1);
  }
  if (cfg->dist && !dist2d && D.alias_lo[1] > 0) {
    // lalias intersected with this rank's y-slab [rank*nyl + 1, (rank+1)*nyl], shifted to local 1-based indices
    const long long lo = std::max<long long>(D.alias_lo[1], rank * nyl + 1), hi = std::min<long long>(D.alias_hi[1], (rank + 1) * nyl);
    if (lo > hi) { D.alias_lo[1] = 0; D.alias_hi[1] = 0; }
    else { D.alias_lo[1] = (int32_t)(lo - rank * nyl); D.alias_hi[1] = (int32_t)(hi - rank * nyl); }
  }
#define FFB_TRY(x) do { int _rc = (x); if (_rc) { ffb_problem_destroy(p); return _rc; } } while (0)
  if (cfg->dist) FFB_TRY(ffb_plan_create_dist(&p->plan, p->nd, cfg->n, p->dtype, cfg->dist, 0));
  else FFB_TRY(ffb_plan_create(&p->plan, p->nd, cfg->n, p->dtype, FFB_R2C, 1, FFB_PLAN_DEFAULT));
  // wavenumbers and Krsq / invKrsq
  p->l = p->m = nullptr;
  FFB_TRY(dalloc(p, &p->kr, (size_t)p->nkr * es, false));
  FFB_TRY(ffb_wavenumbers(p->kr, p->n[0], cfg->L[0], p->dtype, 1));
  if (dist2d) {
    // this rank's kx block followed by the Nyquist wavenumber (a harmless stand-in on the ranks whose extra column is padding)
    void* krl = nullptr;
    FFB_TRY(dalloc(p, &krl, (size_t)nk0 * es, false));
    FFB_TRY(ffb_d2d(krl, reinterpret_cast<char*>(p->kr) + (size_t)rank * kb * es, (size_t)kb * es));
    FFB_TRY(ffb_d2d(reinterpret_cast<char*>(krl) + (size_t)kb * es, reinterpret_cast<char*>(p->kr) + (size_t)(p->nkr - 1) * es, es));
    p->kr = krl;
  }
  if (p->nd >= 2) {
    FFB_TRY(dalloc(p, &p->l, (size_t)p->n[1] * es, false));
    FFB_TRY(ffb_wavenumbers(p->l, p->n[1], cfg->L[1], p->dtype, 0));
    if (cfg->dist && !dist2d) p->l = reinterpret_cast<char*>(p->l) + (size_t)rank * nyl * es;   // this rank's slice of the y-wavenumbers
  }
  if (p->nd >= 3) { FFB_TRY(dalloc(p, &p->m, (size_t)p->n[2] * es, false)); FFB_TRY(ffb_wavenumbers(p->m, p->n[2], cfg->L[2], p->dtype, 0)); }
  p->Krsq = p->invKrsq = p->Ldense = p->filter = nullptr;
  const bool need_inv = cfg->calcN == FFB_CALCN_VORTICITY2D;
  if (need_inv) FFB_TRY(dalloc(p, &p->invKrsq, p->rbytes, false));
  if (!cfg->scalar_zero_L) {
    // L = -nu * Krsq  (`@. L = -κ * kr^2`, src/diffusion.jl:84): Krsq is formed in Ldense, then scaled in place
    FFB_TRY(dalloc(p, &p->Ldense, p->rbytes, false));
    FFB_TRY(ffb_ksq(p->Ldense, p->invKrsq, p->kr, p->l, p->m, &D));
    FFB_TRY(ffb_ew_axpby(p->Ldense, -cfg->nu, p->Ldense, 0.0, nullptr, 0, p->dtype, p->nspec));
    p->L.ptr = p->Ldense; p->L.kind = FFB_COEF_REAL; p->L.dtype = p->dtype; p->L.re = p->L.im = 0;
  } else {
    if (need_inv) FFB_TRY(ffb_ksq(nullptr, p->invKrsq, p->kr, p->l, p->m, &D));
    p->L.ptr = nullptr; p->L.kind = FFB_COEF_SCALAR; p->L.dtype = p->dtype; p->L.re = p->L.im = 0;
  }
  if (cfg->filtered) {
    FFB_TRY(dalloc(p, &p->filter, p->rbytes, false));
    const double dx = roundT(p, cfg->L[0] / (double)p->n[0]), dy = roundT(p, cfg->L[1] / (double)p->n[1]), dz = roundT(p, cfg->L[2] / (double)p->n[2]);
    // every filter parameter has its own "0 = reference default" sentinel (src/domains.jl:506: order=4, innerK=2/3, outerK=1, tol=1e-15);
    // innerK = 0 is a legal value of the reference (test/test_grid.jl passes innerK=0): a NEGATIVE innerK selects the default
    const double order = cfg->filter_order > 0 ? cfg->filter_order : 4;
    const double outerK = cfg->filter_outerK > 0 ? cfg->filter_outerK : 1.0;
    const double tol = cfg->filter_tol > 0 ? cfg->filter_tol : 1e-15;
    const double innerK = (cfg->filter_innerK > 0 || (cfg->filter_innerK == 0 && (cfg->filter_outerK > 0 || cfg->filter_tol > 0))) ? cfg->filter_innerK : 2.0 / 3.0;
    FFB_TRY(ffb_make_filter(p->filter, p->kr, p->l, p->m, dx, p->nd >= 2 ? dy : 0, p->nd >= 3 ? dz : 0, order, innerK, outerK, tol, &D));
  }
  // state and stepper arrays (`zeros(dev, eqn.T, eqn.dims)`, src/problem.jl:108; @devzeros in each stepper constructor)
  FFB_TRY(dalloc(p, &p->sol, p->sbytes, true));
  static const int narr[5] = {1, 5, 2, 6, 3};
  for (int i = 0; i < 6; ++i) p->arr[i] = nullptr;
  for (int i = 0; i < narr[cfg->stepper]; ++i) FFB_TRY(dalloc(p, &p->arr[i], p->sbytes, true));
  p->dt = roundT(p, cfg->dt); p->t = 0; p->step = 0;   // Clock{T}(dt, 0, 0), src/problem.jl:104
  if (cfg->stepper == FFB_ETDRK4) {
    const int cd = cfg->coef_dtype == FFB_F64 ? FFB_F64 : p->dtype;
    ffb_coef* cs[6] = {&p->cE, &p->cE2, &p->cz, &p->ca, &p->cb, &p->cg};
    void** ps[6] = {&p->E, &p->E2, &p->zeta, &p->alpha, &p->beta, &p->gamma};
    if (p->L.kind == FFB_COEF_SCALAR) {
      double hs[12];
      FFB_TRY(ffb_etd_coeffs(cfg->dt, &p->L, p->dtype, cd, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, hs));
      for (int i = 0; i < 6; ++i) { cs[i]->ptr = nullptr; cs[i]->kind = FFB_COEF_SCALAR; cs[i]->dtype = cd; cs[i]->re = hs[2 * i]; cs[i]->im = hs[2 * i + 1]; *ps[i] = nullptr; }
    } else {
      const size_t cb = (size_t)p->nspec * dtype_size(cd);
      for (int i = 0; i < 6; ++i) { FFB_TRY(dalloc(p, ps[i], cb, false)); cs[i]->ptr = *ps[i]; cs[i]->kind = FFB_COEF_REAL; cs[i]->dtype = cd; cs[i]->re = cs[i]->im = 0; }
      FFB_TRY(ffb_etd_coeffs(cfg->dt, &p->L, p->dtype, cd, p->nspec, p->E, p->E2, p->zeta, p->alpha, p->beta, p->gamma, nullptr));
    }
  }
  // vars
  p->sh1 = p->sh2 = p->ph1 = p->ph2 = p->ph3 = nullptr;
  FFB_TRY(dalloc(p, &p->ph1, p->pbytes, true));
  if (cfg->calcN == FFB_CALCN_VORTICITY2D) {
    FFB_TRY(dalloc(p, &p->sh1, p->sbytes, true)); FFB_TRY(dalloc(p, &p->sh2, p->sbytes, true));
    FFB_TRY(dalloc(p, &p->ph2, p->pbytes, true)); FFB_TRY(dalloc(p, &p->ph3, p->pbytes, true));
  } else if (cfg->calcN == FFB_CALCN_DIFFUSION) {
    FFB_TRY(dalloc(p, &p->sh1, p->sbytes, true));
  }
  // Burgers: the spectral scratch of the unfused form (rfft output before `-1/2 im kr` + dealias!) is allocated on first use
  FFB_TRY(ffb_sync());
#undef FFB_TRY
  *out = p;
  return FFB_OK;
}

int ffb_problem_sol(ffb_problem* p, void** sol, int64_t* n) {
  FFB_REQUIRE(p, FFB_EINVAL, "prob is NULL");
  if (sol) *sol = p->sol;
  if (n) *n = p->nspec;
  return FFB_OK;
}

int ffb_problem_plan(ffb_problem* p, ffb_plan** plan) {
  FFB_REQUIRE(p && plan, FFB_EINVAL, "NULL argument");
  *plan = p->plan;
  return FFB_OK;
}

int ffb_problem_clock(ffb_problem* p, double* t, int64_t* step, double* dt) {
  FFB_REQUIRE(p, FFB_EINVAL, "prob is NULL");
  if (t) *t = p->t;
  if (step) *step = p->step;
  if (dt) *dt = p->dt;
  return FFB_OK;
}

int ffb_problem_set_dt(ffb_problem* p, double dt) {
  FFB_REQUIRE(p, FFB_EINVAL, "prob is NULL");
  p->dt = roundT(p, dt);
  return FFB_OK;
}

int ffb_problem_bytes(ffb_problem* p, size_t* b) {
  FFB_REQUIRE(p && b, FFB_EINVAL, "NULL argument");
  size_t ws = 0;
  ffb_plan_workspace_bytes(p->plan, &ws);
  *b = p->device_bytes + ws;
  return FFB_OK;
}

int ffb_problem_set_physical(ffb_problem* p, const void* host_field) {
  FFB_REQUIRE(p && host_field, FFB_EINVAL, "NULL argument");
  int rc = ffb_h2d(p->ph1, host_field, p->pbytes);
  if (rc) return rc;
  return ffb_fft_forward(p->plan, p->ph1, p->sol);
}

int ffb_problem_get_physical(ffb_problem* p, void* host_field) {
  FFB_REQUIRE(p && host_field, FFB_EINVAL, "NULL argument");
  int rc = ffb_fft_inverse(p->plan, p->sol, p->ph1);
  if (rc) return rc;
  return ffb_d2h(host_field, p->ph1, p->pbytes);
}

int ffb_step(ffb_problem* p, int64_t nsteps) {
  FFB_REQUIRE(p, FFB_EINVAL, "prob is NULL");
  for (int64_t i = 0; i < nsteps; ++i) {
    int rc = step_once(p);
    if (rc) return rc;
  }
  return FFB_OK;
}

int ffb_step_until(ffb_problem* p, double stop_time) {
  FFB_REQUIRE(p, FFB_EINVAL, "prob is NULL");
  if (p->cfg.stepper == FFB_ETDRK4)
    return set_error(FFB_ESTEPPER, "step_until! requires fully explicit time stepper; does not work with ETDRK4");
  if (!(stop_time > p->t)) return set_error(FFB_EINVAL, "stop_time must be greater than prob.clock.t");
  const double dt = p->dt;
  const double time_interval = roundT(p, stop_time - p->t);
  const int64_t nsteps = (int64_t)std::floor(roundT(p, time_interval / dt));
  int rc = ffb_step(p, nsteps);
  if (rc) return rc;
  // `t_remaining = time_interval - prob.clock.t` (src/timesteppers.jl:752): only right when the run began at t = 0
  p->dt = roundT(p, time_interval - p->t);
  rc = ffb_step(p, 1);
  p->dt = dt;
  return rc;
}

// ---------------------------------------------------------------- host-buffer pipeline (e2e path: host state in, host state out)
// The blocking form -- ffb_h2d, ffb_step, ffb_d2h -- leaves the GPU idle during both copies (C3: 2 x 9 ms of PCIe around a 15 ms
// step).  A pipeline owns `depth` slots (device staging in / out, events) and two copy streams: ffb_pipeline_submit uploads the
// caller's pinned state on the copy-in stream, the library stream waits for it, moves it into `sol`, steps and stages the result,
// and the copy-out stream downloads it -- so the upload of submission i+1 and the download of i-1 run beside the steps of i.
struct ffb_pipeline {
  ffb_problem* prob;
  int depth, next;
  size_t bytes;
  std::vector<void*> din, dout;
  std::vector<cudaEvent_t> up, stepped, down;
  std::vector<int> busy;
  cudaStream_t s_in, s_out;
};

int ffb_pipeline_destroy(ffb_pipeline* q) {
  if (!q) return FFB_OK;
  if (q->s_in) cudaStreamSynchronize(q->s_in);
  if (q->s_out) cudaStreamSynchronize(q->s_out);
  ffb_sync();
  for (int i = 0; i < q->depth; ++i) {
    if (q->din[i]) cudaFree(q->din[i]);
    if (q->dout[i]) cudaFree(q->dout[i]);
    if (q->up[i]) cudaEventDestroy(q->up[i]);
    if (q->stepped[i]) cudaEventDestroy(q->stepped[i]);
    if (q->down[i]) cudaEventDestroy(q->down[i]);
  }
  if (q->s_in) cudaStreamDestroy(q->s_in);
  if (q->s_out) cudaStreamDestroy(q->s_out);
  delete q;
  return FFB_OK;
}

int ffb_pipeline_create(ffb_pipeline** out, ffb_problem* p, int depth) {
  FFB_REQUIRE(out && p, FFB_EINVAL, "NULL argument");
  *out = nullptr;
  FFB_REQUIRE(depth >= 1 && depth <= 16, FFB_EINVAL, "depth must be 1..16");
  // every submission is an independent state: a multistep scheme would mix the histories of different submissions
  FFB_REQUIRE(p->cfg.stepper != FFB_AB3, FFB_EUNSUPPORTED, "AB3 keeps the previous right-hand sides: no independent submissions");
  auto* q = new ffb_pipeline();
  q->prob = p; q->depth = depth; q->next = 0; q->bytes = p->sbytes; q->s_in = q->s_out = nullptr;
  q->din.assign(depth, nullptr); q->dout.assign(depth, nullptr);
  q->up.assign(depth, nullptr); q->stepped.assign(depth, nullptr); q->down.assign(depth, nullptr); q->busy.assign(depth, 0);
  auto fail = [&](int rc) { ffb_pipeline_destroy(q); return rc; };
  if (cudaStreamCreateWithFlags(&q->s_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&q->s_out, cudaStreamNonBlocking) != cudaSuccess)
    return fail(set_error(FFB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError())));
  for (int i = 0; i < depth; ++i) {
    int rc;
    if ((rc = ffb_malloc(&q->din[i], q->bytes)) || (rc = ffb_malloc(&q->dout[i], q->bytes))) return fail(rc);
    if (cudaEventCreateWithFlags(&q->up[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&q->stepped[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&q->down[i], cudaEventDisableTiming) != cudaSuccess)
      return fail(set_error(FFB_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError())));
  }
  *out = q;
  return FFB_OK;
}

int ffb_pipeline_wait(ffb_pipeline* q, int ticket) {
  FFB_REQUIRE(q && ticket >= 0 && ticket < q->depth, FFB_EINVAL, "bad ticket");
  if (!q->busy[ticket]) return FFB_OK;
  FFB_CUDA(cudaEventSynchronize(q->down[ticket]));
  q->busy[ticket] = 0;
  return FFB_OK;
}

int ffb_pipeline_submit(ffb_pipeline* q, const void* host_in, void* host_out, int64_t nsteps, int* ticket) {
  FFB_REQUIRE(q && host_in && host_out && ticket, FFB_EINVAL, "NULL argument");
  FFB_REQUIRE(nsteps >= 0, FFB_EINVAL, "nsteps = %lld", (long long)nsteps);
  ffb_problem* p = q->prob;
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  const int i = q->next;
  int rc = ffb_pipeline_wait(q, i);   // ring full: the oldest submission has to land first (its host_out is complete after this)
  if (rc) return rc;
  FFB_CUDA(cudaMemcpyAsync(q->din[i], host_in, q->bytes, cudaMemcpyHostToDevice, q->s_in));
  FFB_CUDA(cudaEventRecord(q->up[i], q->s_in));
  FFB_CUDA(cudaStreamWaitEvent(st, q->up[i], 0));
  FFB_CUDA(cudaMemcpyAsync(p->sol, q->din[i], q->bytes, cudaMemcpyDeviceToDevice, st));
  if ((rc = ffb_step(p, nsteps))) return rc;
  FFB_CUDA(cudaMemcpyAsync(q->dout[i], p->sol, q->bytes, cudaMemcpyDeviceToDevice, st));
  FFB_CUDA(cudaEventRecord(q->stepped[i], st));
  FFB_CUDA(cudaStreamWaitEvent(q->s_out, q->stepped[i], 0));
  FFB_CUDA(cudaMemcpyAsync(host_out, q->dout[i], q->bytes, cudaMemcpyDeviceToHost, q->s_out));
  FFB_CUDA(cudaEventRecord(q->down[i], q->s_out));
  q->busy[i] = 1;
  q->next = (i + 1) % q->depth;
  *ticket = i;
  return FFB_OK;
}

}  // extern "C"
