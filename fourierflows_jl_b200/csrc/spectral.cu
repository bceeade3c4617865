// Grid-side kernels of src/domains.jl (wavenumbers, Ksq/invKsq, dealias!, makefilter), the ETD coefficient
// precompute of src/timesteppers.jl:673-721, the closed elementwise vocabulary used by calcN! implementations
// (src/diffusion.jl:136-140) and the Parseval reductions of src/utils.jl:113-183.
#include <complex>
#include <type_traits>
#include "ffb_common.cuh"

namespace ffb {

// ---------------------------------------------------------------- wavenumbers (src/domains.jl:77-78)
template <typename T>
__global__ void wavenumber_kernel(T* out, long long n, long long count, double mult, int real_half) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  // fftfreq: n_nonneg = ((n-1)>>1)+1; element = (i - (i < n_nonneg ? 0 : n)) * (fs/n);  rfftfreq: i * (fs/n)
  const long long nn = ((n - 1) >> 1) + 1;
  const long long s = real_half ? i : (i < nn ? i : i - n);
  out[i] = (T)((double)s * mult);
}

// ---------------------------------------------------------------- row decomposition helper
struct RowIdx { long long j, k, f; };
FFB_D RowIdx split_row(long long row, long long n1, long long n2) {
  RowIdx r;
  r.j = row % n1;
  const long long t = row / n1;
  r.k = t % n2;
  r.f = t / n2;
  return r;
}

// ---------------------------------------------------------------- Ksq / invKsq (src/domains.jl:197-203, 338-344)
template <typename T>
__global__ void ksq_kernel(T* ksq, T* inv, const T* kx, const T* l, const T* m, long long n0, long long n1, long long n2, int ndim) {
  const long long row = blockIdx.x;
  const RowIdx r = split_row(row, n1, n2);
  const T lv = ndim >= 2 ? l[r.j] : T(0), mv = ndim >= 3 ? m[r.k] : T(0);
  for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
    const T kv = kx[i];
    T v = kv * kv;
    if (ndim >= 2) v = v + lv * lv;
    if (ndim >= 3) v = v + mv * mv;
    const long long idx = row * n0 + i;
    if (ksq) ksq[idx] = v;
    if (inv) inv[idx] = (v == T(0)) ? T(0) : T(1) / v;  // `invKsq[1,1,1] = 0`: the origin is the only zero of Ksq (also true on a slab)
  }
}

// ---------------------------------------------------------------- dealias! (src/domains.jl:428-476)
template <typename T>
__global__ void dealias_kernel(cx<T>* fh, long long n0, long long n1, long long n2, int lo0, int hi0, int lo1, int hi1, int lo2, int hi2) {
  const long long row = blockIdx.x;
  const RowIdx r = split_row(row, n1, n2);
  const bool whole = (lo1 > 0 && r.j >= lo1 - 1 && r.j < hi1) || (lo2 > 0 && r.k >= lo2 - 1 && r.k < hi2);
  cx<T>* p = fh + row * n0;
  const cx<T> z = mk<T>(0, 0);
  if (whole) {
    for (long long i = threadIdx.x; i < n0; i += blockDim.x) p[i] = z;
  } else if (lo0 > 0) {
    for (long long i = lo0 - 1 + threadIdx.x; i < hi0; i += blockDim.x) p[i] = z;
  }
}

// ---------------------------------------------------------------- makefilter (src/domains.jl:506-541)
template <typename T>
__global__ void filter_kernel(T* filt, const T* kx, const T* l, const T* m, T dx, T dy, T dz, double order, double innerK, double decay,
                              long long n0, long long n1, long long n2, int ndim) {
  const long long row = blockIdx.x;
  const RowIdx r = split_row(row, n1, n2);
  const T pi = (T)3.14159265358979323846;
  const T b = ndim >= 2 ? l[r.j] * dy / pi : T(0), c = ndim >= 3 ? m[r.k] * dz / pi : T(0);
  const int iord = (int)order;
  const bool integral = ((double)iord == order) && iord >= 1;
  for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
    const T a = kx[i] * dx / pi;
    T K;
    if (ndim == 1) K = a < 0 ? -a : a;  // kr >= 0; abs(k*dx/pi) for complex variables
    else if (ndim == 2) K = sqrt(a * a + b * b);
    else K = sqrt(a * a + b * b + c * c);
    const double Kd = (double)K;
    double f;
    if (Kd < innerK) {
      f = 1.0;
    } else {
      const double d = Kd - innerK;
      double pw;
      if (integral) { pw = d; for (int q = 1; q < iord; ++q) pw *= d; }
      else pw = pow(d, order);
      f = exp(-decay * pw);
    }
    filt[row * n0 + i] = (T)f;
  }
}

// ---------------------------------------------------------------- ETD coefficients (src/timesteppers.jl:673-721)
struct zc_t { double x, y; };
FFB_HD zc_t zmul(zc_t a, zc_t b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
FFB_HD zc_t zadd(zc_t a, zc_t b) { return {a.x + b.x, a.y + b.y}; }
FFB_HD zc_t zsub(zc_t a, zc_t b) { return {a.x - b.x, a.y - b.y}; }
FFB_HD zc_t zdiv(zc_t a, zc_t b) {
  // Smith's algorithm (robust like Julia's complex division)
  if (fabs(b.x) >= fabs(b.y)) {
    const double r = b.y / b.x, den = b.x + b.y * r;
    return {(a.x + a.y * r) / den, (a.y - a.x * r) / den};
  }
  const double r = b.x / b.y, den = b.x * r + b.y;
  return {(a.x * r + a.y) / den, (a.y * r - a.x) / den};
}
FFB_HD zc_t zexp(zc_t a) {
  const double e = exp(a.x);
  double s, c;
#ifdef __CUDA_ARCH__
  sincos(a.y, &s, &c);
#else
  s = sin(a.y); c = cos(a.y);
#endif
  return {e * c, e * s};
}

struct EtdOut { zc_t E, E2, zeta, alpha, beta, gamma; };

// one element: dtL = dt*L already formed in the problem precision T (as in `dt * L .+ circ`), widened to Float64
FFB_HD EtdOut etd_element(zc_t dtL) {
  EtdOut o;
  zc_t sz = {0, 0}, sa = {0, 0}, sb = {0, 0}, sg = {0, 0};
  for (int j = 0; j < 32; ++j) {
    double s, c;
#ifdef __CUDA_ARCH__
    sincospi((2.0 * j + 1.0) / 32.0, &s, &c);
#else
    const double ang = 3.14159265358979323846 * (2.0 * j + 1.0) / 32.0;
    s = sin(ang); c = cos(ang);
#endif
    const zc_t zc = {dtL.x + c, dtL.y + s};
    const zc_t ez = zexp(zc), ez2 = zexp({zc.x / 2, zc.y / 2});
    const zc_t z2 = zmul(zc, zc), z3 = zmul(z2, zc);
    const zc_t one = {1, 0};
    // zeta = (exp(z/2) - 1)/z
    sz = zadd(sz, zdiv(zsub(ez2, one), zc));
    // alpha = (-4 - z + exp(z)(4 - 3z + z^2))/z^3
    zc_t t = {4 - 3 * zc.x + z2.x, -3 * zc.y + z2.y};
    zc_t num = zadd({-4 - zc.x, -zc.y}, zmul(ez, t));
    sa = zadd(sa, zdiv(num, z3));
    // beta = (2 + z + exp(z)(-2 + z))/z^3
    num = zadd({2 + zc.x, zc.y}, zmul(ez, {-2 + zc.x, zc.y}));
    sb = zadd(sb, zdiv(num, z3));
    // gamma = (-4 - 3z - z^2 + exp(z)(4 - z))/z^3
    num = zadd({-4 - 3 * zc.x - z2.x, -3 * zc.y - z2.y}, zmul(ez, {4 - zc.x, -zc.y}));
    sg = zadd(sg, zdiv(num, z3));
  }
  o.zeta = {sz.x / 32, sz.y / 32};
  o.alpha = {sa.x / 32, sa.y / 32};
  o.beta = {sb.x / 32, sb.y / 32};
  o.gamma = {sg.x / 32, sg.y / 32};
  o.E = zexp(dtL);
  o.E2 = zexp({dtL.x / 2, dtL.y / 2});
  return o;
}

template <typename CT> FFB_D void put_coef(void* p, long long i, zc_t v, bool cplx) {
  if (cplx) reinterpret_cast<cx<CT>*>(p)[i] = mk<CT>((CT)v.x, (CT)v.y);
  else reinterpret_cast<CT*>(p)[i] = (CT)v.x;
}

template <typename T, typename CT>
__global__ void etd_kernel(double dt, const void* L, int cplx, long long n, void* E, void* E2, void* zeta, void* alpha, void* beta, void* gamma) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T dtT = (T)dt;
  zc_t dtL;
  if (cplx) {
    const cx<T> l = reinterpret_cast<const cx<T>*>(L)[i];
    dtL = {(double)(dtT * l.x), (double)(dtT * l.y)};
  } else {
    dtL = {(double)(dtT * reinterpret_cast<const T*>(L)[i]), 0.0};
  }
  const EtdOut o = etd_element(dtL);
  const zc_t dtz = {(double)dtT, 0};
  // expLdt, exphLdt are evaluated in T (getexpLs), the contour means in Float64 and scaled by dt
  put_coef<CT>(E, i, {(double)(T)o.E.x, (double)(T)o.E.y}, cplx);
  put_coef<CT>(E2, i, {(double)(T)o.E2.x, (double)(T)o.E2.y}, cplx);
  put_coef<CT>(zeta, i, zmul(dtz, o.zeta), cplx);
  put_coef<CT>(alpha, i, zmul(dtz, o.alpha), cplx);
  put_coef<CT>(beta, i, zmul(dtz, o.beta), cplx);
  put_coef<CT>(gamma, i, zmul(dtz, o.gamma), cplx);
}

// ---------------------------------------------------------------- elementwise vocabulary
template <typename T, int CPLX>
__global__ void axpby_kernel(void* out, T a, const void* x, T b, const void* y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if constexpr (CPLX) {
      const cx<T> xv = reinterpret_cast<const cx<T>*>(x)[i];
      cx<T> r = a * xv;
      if (y) r = r + b * reinterpret_cast<const cx<T>*>(y)[i];
      reinterpret_cast<cx<T>*>(out)[i] = r;
    } else {
      T r = a * reinterpret_cast<const T*>(x)[i];
      if (y) r = r + b * reinterpret_cast<const T*>(y)[i];
      reinterpret_cast<T*>(out)[i] = r;
    }
  }
}

template <typename T>
__global__ void mul_real_kernel(T* out, const T* x, const T* y, long long n) {
  // 16-byte vectors when all three arrays are 16-byte aligned, scalar tail / fallback otherwise
  constexpr int V = 16 / sizeof(T);
  using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const long long nv = aligned ? n / V : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const V4 a = reinterpret_cast<const V4*>(x)[i], b = reinterpret_cast<const V4*>(y)[i];
    V4 r;
    if constexpr (sizeof(T) == 4) { r.x = a.x * b.x; r.y = a.y * b.y; r.z = a.z * b.z; r.w = a.w * b.w; }
    else { r.x = a.x * b.x; r.y = a.y * b.y; }
    reinterpret_cast<V4*>(out)[i] = r;
  }
  for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] * y[i];
}

template <typename T, int XC, int YC>
__global__ void mul_any_kernel(void* out, const void* x, const void* y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if constexpr (XC && YC) reinterpret_cast<cx<T>*>(out)[i] = reinterpret_cast<const cx<T>*>(x)[i] * reinterpret_cast<const cx<T>*>(y)[i];
    else if constexpr (XC) { const T r = reinterpret_cast<const T*>(y)[i]; const cx<T> c = reinterpret_cast<const cx<T>*>(x)[i]; reinterpret_cast<cx<T>*>(out)[i] = mk<T>(c.x * r, c.y * r); }
    else if constexpr (YC) { const T r = reinterpret_cast<const T*>(x)[i]; const cx<T> c = reinterpret_cast<const cx<T>*>(y)[i]; reinterpret_cast<cx<T>*>(out)[i] = mk<T>(r * c.x, r * c.y); }
    else reinterpret_cast<T*>(out)[i] = reinterpret_cast<const T*>(x)[i] * reinterpret_cast<const T*>(y)[i];
  }
}

template <typename T> FFB_D T ipow(T v, int p) {  // p >= 1; `k^2` is `k*k` in Julia (literal_pow)
  T r = v;
  for (int q = 1; q < p; ++q) r = r * v;
  return r;
}

template <typename T>
__global__ void spectral_mul_kernel(cx<T>* out, const cx<T>* in, T ar, T ai, const T* kx, int px, const T* l, int py, const T* m, int pz,
                                    const T* w, int accumulate, int dealias, long long n0, long long n1, long long n2, int lo0, int hi0,
                                    int lo1, int hi1, int lo2, int hi2) {
  const long long row = blockIdx.x;
  const RowIdx r = split_row(row, n1, n2);
  const bool whole = dealias && ((lo1 > 0 && r.j >= lo1 - 1 && r.j < hi1) || (lo2 > 0 && r.k >= lo2 - 1 && r.k < hi2));
  const long long wrow = (r.k * n1 + r.j) * n0;  // w has no field dimension
  for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
    const long long idx = row * n0 + i;
    if (whole || (dealias && lo0 > 0 && i >= lo0 - 1 && i < hi0)) { out[idx] = mk<T>(0, 0); continue; }
    // factor evaluated left to right like `im * kr * invKrsq * sol`
    T fr = ar, fi = ai;
    if (px) { const T v = ipow(kx[i], px); fr *= v; fi *= v; }
    if (py) { const T v = ipow(l[r.j], py); fr *= v; fi *= v; }
    if (pz) { const T v = ipow(m[r.k], pz); fr *= v; fi *= v; }
    if (w) { const T v = w[wrow + i]; fr *= v; fi *= v; }
    cx<T> res = mk<T>(fr, fi) * in[idx];
    if (accumulate) res = out[idx] + res;
    out[idx] = res;
  }
}

// ---------------------------------------------------------------- Parseval sums (src/utils.jl:113-183)
template <typename T, int ABS2>
__global__ void parseval_kernel(double* acc, const cx<T>* uh, long long n0, long long nrows, int half) {
  double s = 0;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const cx<T>* p = uh + row * n0;
    for (long long i = threadIdx.x; i < n0; i += blockDim.x) {
      const cx<T> v = p[i];
      const double t = ABS2 ? (double)v.x * v.x + (double)v.y * v.y : (double)v.x;
      const double wgt = (half && i > 0 && i < n0 - 1) ? 2.0 : 1.0;
      s += wgt * t;
    }
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(acc, s);
  }
}

static int check_desc(const ffb_desc* d) {
  FFB_REQUIRE(d, FFB_EINVAL, "desc is NULL");
  FFB_REQUIRE(d->ndim >= 1 && d->ndim <= 3, FFB_EINVAL, "desc.ndim must be 1..3");
  FFB_REQUIRE(d->dtype == FFB_F32 || d->dtype == FFB_F64, FFB_EINVAL, "bad desc.dtype");
  for (int i = 0; i < 4; ++i) FFB_REQUIRE(d->dims[i] >= 1, FFB_EINVAL, "desc.dims[%d] must be >= 1", i);
  for (int i = d->ndim; i < 3; ++i) FFB_REQUIRE(d->dims[i] == 1, FFB_EINVAL, "desc.dims[%d] must be 1 for ndim=%d", i, d->ndim);
  for (int i = 0; i < 3; ++i)
    if (d->alias_lo[i] > 0)
      FFB_REQUIRE(d->alias_lo[i] <= d->alias_hi[i] + 1 && d->alias_hi[i] <= d->dims[i], FFB_EINVAL, "alias range %d out of bounds", i);
  return FFB_OK;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_wavenumbers(void* out, int64_t n, double L, int dtype, int real_half) {
  FFB_REQUIRE(out && n >= 2, FFB_EINVAL, "bad argument");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  // fs = 2π/L*n ; multiplier = fs/n, evaluated in Float64 exactly like `fftfreq(nx, 2π/Lx*nx)`
  const double fs = 2.0 * 3.141592653589793 / L * (double)n;
  const double mult = fs / (double)n;
  const long long count = real_half ? n / 2 + 1 : n;
  const unsigned blocks = (unsigned)((count + 255) / 256);
  if (dtype == FFB_F64) wavenumber_kernel<double><<<blocks, 256, 0, s>>>((double*)out, n, count, mult, real_half);
  else wavenumber_kernel<float><<<blocks, 256, 0, s>>>((float*)out, n, count, mult, real_half);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_ksq(void* ksq, void* invksq, const void* kx, const void* l, const void* m, const ffb_desc* d) {
  int rc = check_desc(d); if (rc) return rc;
  FFB_REQUIRE(kx && (d->ndim < 2 || l) && (d->ndim < 3 || m), FFB_EINVAL, "missing wavenumber vector");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const long long rows = d->dims[1] * d->dims[2];
  if (d->dtype == FFB_F64)
    ksq_kernel<double><<<(unsigned)rows, 128, 0, s>>>((double*)ksq, (double*)invksq, (const double*)kx, (const double*)l, (const double*)m, d->dims[0], d->dims[1], d->dims[2], d->ndim);
  else
    ksq_kernel<float><<<(unsigned)rows, 128, 0, s>>>((float*)ksq, (float*)invksq, (const float*)kx, (const float*)l, (const float*)m, d->dims[0], d->dims[1], d->dims[2], d->ndim);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_dealias(void* fh, const ffb_desc* d) {
  int rc = check_desc(d); if (rc) return rc;
  FFB_REQUIRE(fh, FFB_EINVAL, "fh is NULL");
  if (d->alias_lo[0] <= 0 && d->alias_lo[1] <= 0 && d->alias_lo[2] <= 0) return FFB_OK;  // aliased_fraction = 0: no-op (:434)
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const long long rows = d->dims[1] * d->dims[2] * d->dims[3];
  FFB_REQUIRE(rows < (1ll << 31), FFB_EUNSUPPORTED, "too many rows");
  if (d->dtype == FFB_F64)
    dealias_kernel<double><<<(unsigned)rows, 128, 0, s>>>((cx<double>*)fh, d->dims[0], d->dims[1], d->dims[2], d->alias_lo[0], d->alias_hi[0], d->alias_lo[1], d->alias_hi[1], d->alias_lo[2], d->alias_hi[2]);
  else
    dealias_kernel<float><<<(unsigned)rows, 128, 0, s>>>((cx<float>*)fh, d->dims[0], d->dims[1], d->dims[2], d->alias_lo[0], d->alias_hi[0], d->alias_lo[1], d->alias_hi[1], d->alias_lo[2], d->alias_hi[2]);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_make_filter(void* filter, const void* kx, const void* l, const void* m, double dx, double dy, double dz, double order,
                    double innerK, double outerK, double tol, const ffb_desc* d) {
  int rc = check_desc(d); if (rc) return rc;
  FFB_REQUIRE(filter && kx && (d->ndim < 2 || l) && (d->ndim < 3 || m), FFB_EINVAL, "missing argument");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const double decay = -log(tol) / pow(outerK - innerK, order);  // src/domains.jl:510
  const long long rows = d->dims[1] * d->dims[2] * d->dims[3];
  FFB_REQUIRE(rows < (1ll << 31), FFB_EUNSUPPORTED, "too many rows");
  if (d->dtype == FFB_F64)
    filter_kernel<double><<<(unsigned)rows, 128, 0, s>>>((double*)filter, (const double*)kx, (const double*)l, (const double*)m, dx, dy, dz, order, innerK, decay, d->dims[0], d->dims[1], d->dims[2], d->ndim);
  else
    filter_kernel<float><<<(unsigned)rows, 128, 0, s>>>((float*)filter, (const float*)kx, (const float*)l, (const float*)m, (float)dx, (float)dy, (float)dz, order, innerK, decay, d->dims[0], d->dims[1], d->dims[2], d->ndim);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_etd_coeffs(double dt, const ffb_coef* L, int dtype, int coef_dtype, int64_t n, void* expLdt, void* exphLdt, void* zeta,
                   void* alpha, void* beta, void* gamma, double* host_scalars) {
  FFB_REQUIRE(L, FFB_EINVAL, "L is NULL");
  FFB_REQUIRE(dtype == FFB_F32 || dtype == FFB_F64, FFB_EINVAL, "bad dtype");
  FFB_REQUIRE(coef_dtype == FFB_F64 || coef_dtype == dtype, FFB_EINVAL, "coef_dtype must be Float64 or the state type");
  if (L->kind == FFB_COEF_SCALAR) {
    FFB_REQUIRE(host_scalars, FFB_EINVAL, "host_scalars is NULL for scalar L");
    // A scalar L arrives as a Float64 (or Int) value: Julia forms `dt * L` and `exp(dt * L)` in Float64 after
    // converting dt to the problem's float type (src/timesteppers.jl:457,674-675,696).
    const double dtT = dtype == FFB_F32 ? (double)(float)dt : dt;
    const zc_t dtL = {dtT * L->re, dtT * L->im};
    const EtdOut o = etd_element(dtL);
    const zc_t v[6] = {o.E, o.E2, {dtT * o.zeta.x, dtT * o.zeta.y}, {dtT * o.alpha.x, dtT * o.alpha.y}, {dtT * o.beta.x, dtT * o.beta.y}, {dtT * o.gamma.x, dtT * o.gamma.y}};
    for (int i = 0; i < 6; ++i) {
      double re = v[i].x, im = v[i].y;
      if (L->im == 0.0) im = 0.0;  // real L: `real.(...)` (:710-715)
      host_scalars[2 * i] = re; host_scalars[2 * i + 1] = im;
    }
    return FFB_OK;
  }
  FFB_REQUIRE(L->ptr && expLdt && exphLdt && zeta && alpha && beta && gamma, FFB_EINVAL, "NULL array");
  FFB_REQUIRE((L->dtype == FFB_F64) == (dtype == FFB_F64), FFB_EINVAL, "L must have the problem's precision");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const int cplx = L->kind == FFB_COEF_COMPLEX;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (dtype == FFB_F64) etd_kernel<double, double><<<blocks, 128, 0, s>>>(dt, L->ptr, cplx, n, expLdt, exphLdt, zeta, alpha, beta, gamma);
  else if (coef_dtype == FFB_F64) etd_kernel<float, double><<<blocks, 128, 0, s>>>(dt, L->ptr, cplx, n, expLdt, exphLdt, zeta, alpha, beta, gamma);
  else etd_kernel<float, float><<<blocks, 128, 0, s>>>(dt, L->ptr, cplx, n, expLdt, exphLdt, zeta, alpha, beta, gamma);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_ew_axpby(void* out, double a, const void* x, double b, const void* y, int is_complex, int dtype, int64_t n) {
  FFB_REQUIRE(out && x, FFB_EINVAL, "NULL array");
  if (n <= 0) return FFB_OK;
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)num_sms() * 16);
  if (dtype == FFB_F64) {
    if (is_complex) axpby_kernel<double, 1><<<blocks, 256, 0, s>>>(out, a, x, b, y, n);
    else axpby_kernel<double, 0><<<blocks, 256, 0, s>>>(out, a, x, b, y, n);
  } else {
    if (is_complex) axpby_kernel<float, 1><<<blocks, 256, 0, s>>>(out, (float)a, x, (float)b, y, n);
    else axpby_kernel<float, 0><<<blocks, 256, 0, s>>>(out, (float)a, x, (float)b, y, n);
  }
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_ew_mul_real(void* out, const void* x, const void* y, int dtype, int64_t n) {
  FFB_REQUIRE(out && x && y, FFB_EINVAL, "NULL array");
  if (n <= 0) return FFB_OK;
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)num_sms() * 16);
  if (dtype == FFB_F64) mul_real_kernel<double><<<blocks, 256, 0, s>>>((double*)out, (const double*)x, (const double*)y, n);
  else mul_real_kernel<float><<<blocks, 256, 0, s>>>((float*)out, (const float*)x, (const float*)y, n);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_ew_mul(void* out, const void* x, int xc, const void* y, int yc, int dtype, int64_t n) {
  FFB_REQUIRE(out && x && y, FFB_EINVAL, "NULL array");
  FFB_REQUIRE(dtype == FFB_F32 || dtype == FFB_F64, FFB_EINVAL, "bad dtype");
  if (n <= 0) return FFB_OK;
  if (!xc && !yc) return ffb_ew_mul_real(out, x, y, dtype, n);
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)num_sms() * 16);
#define FFB_MULANY(T)                                                                    \
  do {                                                                                   \
    if (xc && yc) mul_any_kernel<T, 1, 1><<<blocks, 256, 0, s>>>(out, x, y, n);          \
    else if (xc) mul_any_kernel<T, 1, 0><<<blocks, 256, 0, s>>>(out, x, y, n);           \
    else mul_any_kernel<T, 0, 1><<<blocks, 256, 0, s>>>(out, x, y, n);                   \
  } while (0)
  if (dtype == FFB_F64) FFB_MULANY(double); else FFB_MULANY(float);
#undef FFB_MULANY
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_ew_spectral_mul(void* out, const void* in, double ar, double ai, const void* kx, int px, const void* l, int py, const void* m,
                        int pz, const void* w, int accumulate, int dealias, const ffb_desc* d) {
  int rc = check_desc(d); if (rc) return rc;
  FFB_REQUIRE(out && in, FFB_EINVAL, "NULL array");
  FFB_REQUIRE(px >= 0 && py >= 0 && pz >= 0, FFB_EINVAL, "negative power");
  FFB_REQUIRE((!px || kx) && (!py || l) && (!pz || m), FFB_EINVAL, "missing wavenumber vector");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const long long rows = d->dims[1] * d->dims[2] * d->dims[3];
  FFB_REQUIRE(rows < (1ll << 31), FFB_EUNSUPPORTED, "too many rows");
  const int threads = d->dims[0] >= 256 ? 256 : 64;
  if (d->dtype == FFB_F64)
    spectral_mul_kernel<double><<<(unsigned)rows, threads, 0, s>>>((cx<double>*)out, (const cx<double>*)in, ar, ai, (const double*)kx, px, (const double*)l, py, (const double*)m, pz, (const double*)w, accumulate, dealias, d->dims[0], d->dims[1], d->dims[2], d->alias_lo[0], d->alias_hi[0], d->alias_lo[1], d->alias_hi[1], d->alias_lo[2], d->alias_hi[2]);
  else
    spectral_mul_kernel<float><<<(unsigned)rows, threads, 0, s>>>((cx<float>*)out, (const cx<float>*)in, (float)ar, (float)ai, (const float*)kx, px, (const float*)l, py, (const float*)m, pz, (const float*)w, accumulate, dealias, d->dims[0], d->dims[1], d->dims[2], d->alias_lo[0], d->alias_hi[0], d->alias_lo[1], d->alias_hi[1], d->alias_lo[2], d->alias_hi[2]);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

int ffb_parseval_sum(double* host_result, const void* uh, int abs2, int half, const ffb_desc* d) {
  int rc = check_desc(d); if (rc) return rc;
  FFB_REQUIRE(host_result && uh, FFB_EINVAL, "NULL argument");
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  static double* acc = nullptr;
  if (!acc) FFB_CUDA(cudaMalloc(&acc, sizeof(double)));
  FFB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
  const long long rows = d->dims[1] * d->dims[2] * d->dims[3];
  const unsigned blocks = (unsigned)std::min<long long>(rows, (long long)num_sms() * 8);
  if (d->dtype == FFB_F64) {
    if (abs2) parseval_kernel<double, 1><<<blocks, 256, 0, s>>>(acc, (const cx<double>*)uh, d->dims[0], rows, half);
    else parseval_kernel<double, 0><<<blocks, 256, 0, s>>>(acc, (const cx<double>*)uh, d->dims[0], rows, half);
  } else {
    if (abs2) parseval_kernel<float, 1><<<blocks, 256, 0, s>>>(acc, (const cx<float>*)uh, d->dims[0], rows, half);
    else parseval_kernel<float, 0><<<blocks, 256, 0, s>>>(acc, (const cx<float>*)uh, d->dims[0], rows, half);
  }
  count_launch();
  FFB_CHECK_LAUNCH();
  FFB_CUDA(cudaMemcpyAsync(host_result, acc, sizeof(double), cudaMemcpyDeviceToHost, s));
  FFB_CUDA(cudaStreamSynchronize(s));
  return FFB_OK;
}

}  // extern "C"
