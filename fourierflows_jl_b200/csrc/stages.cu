// Fused time-stepper stage kernels: every `@.` broadcast group of src/timesteppers.jl becomes one streaming kernel
// that reads each operand once (the B3 seam of SURVEY 8b).  Arithmetic follows the reference's expression order;
// -fmad=false (see Makefile) keeps nvcc from contracting it.  Coefficient operands (`L`, ETD coefficients) may be a
// scalar, a dense real array or a dense complex array, stored as T or as Float64 (getetdcoeffs returns Float64
// even for Float32 problems, src/timesteppers.jl:692,710-715); the kernel computes in the coefficient's precision
// and rounds once on store, like Julia's promotion does.
#include "ffb_common.cuh"

namespace ffb {

template <typename CT, int KIND>
struct CoefView {
  const void* p;
  CT re, im;
  FFB_D cx<CT> mul(long long i, cx<CT> z) const {
    if constexpr (KIND == FFB_COEF_REAL) {
      const CT c = reinterpret_cast<const CT*>(p)[i];
      return mk<CT>(c * z.x, c * z.y);
    } else if constexpr (KIND == FFB_COEF_COMPLEX) {
      const cx<CT> c = reinterpret_cast<const cx<CT>*>(p)[i];
      return c * z;
    } else {
      return mk<CT>(re, im) * z;
    }
  }
  // 2*coef as in `2β * (N₂ + N₃)` (src/timesteppers.jl:502): the product 2β is formed first
  FFB_D cx<CT> mul2(long long i, cx<CT> z) const {
    if constexpr (KIND == FFB_COEF_REAL) {
      const CT c = CT(2) * reinterpret_cast<const CT*>(p)[i];
      return mk<CT>(c * z.x, c * z.y);
    } else if constexpr (KIND == FFB_COEF_COMPLEX) {
      const cx<CT> c = reinterpret_cast<const cx<CT>*>(p)[i];
      return mk<CT>(CT(2) * c.x, CT(2) * c.y) * z;
    } else {
      return mk<CT>(CT(2) * re, CT(2) * im) * z;
    }
  }
};

template <typename CT, int KIND> static CoefView<CT, KIND> view(const ffb_coef* c) {
  CoefView<CT, KIND> v;
  v.p = c->ptr; v.re = (CT)c->re; v.im = (CT)c->im;
  return v;
}

template <typename T, typename CT> FFB_D cx<CT> ld(const void* p, long long i) {
  const cx<T> z = reinterpret_cast<const cx<T>*>(p)[i];
  return mk<CT>((CT)z.x, (CT)z.y);
}
template <typename T, typename CT> FFB_D void st(void* p, long long i, cx<CT> z) {
  reinterpret_cast<cx<T>*>(p)[i] = mk<T>((T)z.x, (T)z.y);
}
template <typename T, typename CT> FFB_D cx<CT> apply_filter(const void* f, long long i, cx<CT> z) {
  // `sol *= filter` / `filter * (...)`: filter has the state's real type T
  const T c = reinterpret_cast<const T*>(f)[i];
  const cx<T> r = mk<T>((T)z.x, (T)z.y);  // the unfiltered value is rounded to T first, as in the reference's two statements
  return mk<CT>((CT)(c * r.x), (CT)(c * r.y));
}

template <class Op> __global__ void __launch_bounds__(256) stage_kernel(const Op op, const long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    op(i); op(i + stride); op(i + 2 * stride); op(i + 3 * stride);
  }
  for (; i < n; i += stride) op(i);
}

// `arrays` = number of complex state arrays read + written, `coefs` = number of coefficient operands, `filt` = filter read
template <typename T, typename CT, int KIND, class Op>
static int launch(const Op& op, long long n, const char* name, int arrays, int coefs, bool filt) {
  if (n <= 0) return FFB_OK;
  const double cbytes = KIND == FFB_COEF_REAL ? sizeof(CT) : KIND == FFB_COEF_COMPLEX ? 2 * sizeof(CT) : 0;
  ProfScope ps(name, (double)n * (arrays * 2.0 * sizeof(T) + coefs * cbytes + (filt ? sizeof(T) : 0)));
  cudaStream_t s = current_stream();
  FFB_REQUIRE(s, FFB_ECUDA, "no CUDA stream (no device?)");
  const int threads = 256;
  long long blocks = (n + threads * 4 - 1) / (threads * 4);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  stage_kernel<Op><<<(unsigned)blocks, threads, 0, s>>>(op, n);
  count_launch();
  FFB_CHECK_LAUNCH();
  return FFB_OK;
}

// ---------------------------------------------------------------- operators
// ForwardEuler (src/timesteppers.jl:113): sol += dt*(L*sol + N);  Filtered (:144): sol = filter*(sol + dt*(N + L*sol))
template <typename T, typename CT, int KIND> struct OpFE {
  void* sol; const void* N; CoefView<CT, KIND> L; CT dt; const void* filter;
  FFB_D void operator()(long long i) const {
    const cx<CT> s = ld<T, CT>(sol, i), n = ld<T, CT>(N, i);
    const cx<CT> ls = L.mul(i, s);
    if (filter) {
      const cx<CT> r = s + dt * (n + ls);
      st<T, CT>(sol, i, apply_filter<T, CT>(filter, i, r));
    } else {
      st<T, CT>(sol, i, s + dt * (ls + n));
    }
  }
};

// RK4 (:225-258): rhs += L*u;  sol1 = sol + c*rhs
template <typename T, typename CT, int KIND> struct OpRK4Sub {
  void* sol1; void* rhs; const void* u; const void* sol; CoefView<CT, KIND> L; CT c;
  FFB_D void operator()(long long i) const {
    const cx<CT> uu = ld<T, CT>(u, i);
    cx<CT> r = ld<T, CT>(rhs, i) + L.mul(i, uu);
    const cx<T> rr = mk<T>((T)r.x, (T)r.y);  // value as stored (rounded to T)
    reinterpret_cast<cx<T>*>(rhs)[i] = rr;
    r = mk<CT>((CT)rr.x, (CT)rr.y);
    st<T, CT>(sol1, i, ld<T, CT>(sol, i) + c * r);
  }
};

// RK4 final (:255,261,279): rhs4 += L*sol1;  sol += dt/6*(rhs1 + 2 rhs2 + 2 rhs3 + rhs4) [; sol *= filter]
template <typename T, typename CT, int KIND> struct OpRK4Final {
  void* sol; const void* r1; const void* r2; const void* r3; void* r4; const void* sol1; CoefView<CT, KIND> L; CT dt6;
  const void* filter; int store4;
  FFB_D void operator()(long long i) const {
    cx<CT> k4 = ld<T, CT>(r4, i) + L.mul(i, ld<T, CT>(sol1, i));
    const cx<T> k4t = mk<T>((T)k4.x, (T)k4.y);
    if (store4) reinterpret_cast<cx<T>*>(r4)[i] = k4t;
    k4 = mk<CT>((CT)k4t.x, (CT)k4t.y);
    const cx<CT> k1 = ld<T, CT>(r1, i), k2 = ld<T, CT>(r2, i), k3 = ld<T, CT>(r3, i);
    const cx<CT> sum = ((k1 + CT(2) * k2) + CT(2) * k3) + k4;
    cx<CT> s = ld<T, CT>(sol, i) + dt6 * sum;
    if (filter) s = apply_filter<T, CT>(filter, i, s);
    st<T, CT>(sol, i, s);
  }
};

// LSRK54 (:386-392, 408): rhs += L*sol;  S2 = A*S2 + dt*rhs;  sol += B*S2 [; sol *= filter]
template <typename T, typename CT, int KIND> struct OpLSRK {
  void* sol; void* S2; const void* rhs; CoefView<CT, KIND> L; CT A, B, dt; int first; const void* filter;
  FFB_D void operator()(long long i) const {
    const cx<CT> s = ld<T, CT>(sol, i);
    cx<CT> r = ld<T, CT>(rhs, i) + L.mul(i, s);
    const cx<T> rt = mk<T>((T)r.x, (T)r.y);
    r = mk<CT>((CT)rt.x, (CT)rt.y);
    const cx<CT> s2old = first ? mk<CT>(0, 0) : ld<T, CT>(S2, i);
    cx<CT> s2 = A * s2old + dt * r;
    const cx<T> s2t = mk<T>((T)s2.x, (T)s2.y);
    reinterpret_cast<cx<T>*>(S2)[i] = s2t;
    s2 = mk<CT>((CT)s2t.x, (CT)s2t.y);
    cx<CT> o = s + B * s2;
    if (filter) o = apply_filter<T, CT>(filter, i, o);
    st<T, CT>(sol, i, o);
  }
};

// ETDRK4 substeps 1,2 (:507): out = exphLdt*sol + zeta*N
template <typename T, typename CT, int KIND> struct OpETD12 {
  void* out; CoefView<CT, KIND> E2; const void* sol; CoefView<CT, KIND> zeta; const void* N;
  FFB_D void operator()(long long i) const {
    st<T, CT>(out, i, E2.mul(i, ld<T, CT>(sol, i)) + zeta.mul(i, ld<T, CT>(N, i)));
  }
};
// ETDRK4 substep 3 (:513): out = exphLdt*sol1 + zeta*(2 N3 - N1)
template <typename T, typename CT, int KIND> struct OpETD3 {
  void* out; CoefView<CT, KIND> E2; const void* sol1; CoefView<CT, KIND> zeta; const void* N1; const void* N3;
  FFB_D void operator()(long long i) const {
    const cx<CT> n1 = ld<T, CT>(N1, i), n3 = ld<T, CT>(N3, i);
    // `2N₃ - N₁` is evaluated in the state type T before the (possibly Float64) coefficient multiplies it
    const cx<T> d = mk<T>((T)2 * (T)n3.x - (T)n1.x, (T)2 * (T)n3.y - (T)n1.y);
    st<T, CT>(out, i, E2.mul(i, ld<T, CT>(sol1, i)) + zeta.mul(i, mk<CT>((CT)d.x, (CT)d.y)));
  }
};
// ETDRK4 update (:502,552): sol = expLdt*sol + alpha*N1 + 2beta*(N2+N3) + gamma*N4 [; sol *= filter]
template <typename T, typename CT, int KIND> struct OpETDUpd {
  void* sol; CoefView<CT, KIND> E, al, be, ga; const void* N1; const void* N2; const void* N3; const void* N4; const void* filter;
  FFB_D void operator()(long long i) const {
    const cx<CT> n2 = ld<T, CT>(N2, i), n3 = ld<T, CT>(N3, i);
    const cx<T> s23 = mk<T>((T)n2.x + (T)n3.x, (T)n2.y + (T)n3.y);
    cx<CT> s = ((E.mul(i, ld<T, CT>(sol, i)) + al.mul(i, ld<T, CT>(N1, i))) + be.mul2(i, mk<CT>((CT)s23.x, (CT)s23.y))) +
               ga.mul(i, ld<T, CT>(N4, i));
    if (filter) s = apply_filter<T, CT>(filter, i, s);
    st<T, CT>(sol, i, s);
  }
};

// AB3 (:628-636,640,658): rhs += L*sol; Euler when step < 3 else AB3 (Float64 constants :565-567) [; sol *= filter]
template <typename T, int KIND> struct OpAB3 {
  void* sol; void* rhs; const void* m1; const void* m2; CoefView<T, KIND> L; T dt; int euler; const void* filter;
  FFB_D void operator()(long long i) const {
    const cx<T> s = ld<T, T>(sol, i);
    const cx<T> r = ld<T, T>(rhs, i) + L.mul(i, s);
    reinterpret_cast<cx<T>*>(rhs)[i] = r;
    cx<T> o;
    if (euler) {
      o = s + dt * r;
    } else {
      // ab3h1 * RHS - ab3h2 * RHS_1 + ab3h3 * RHS_2 in Float64 (the constants are Float64), then dt * (...) and the sum
      const cx<T> a = ld<T, T>(m1, i), b = ld<T, T>(m2, i);
      const double h1 = 23.0 / 12.0, h2 = 16.0 / 12.0, h3 = 5.0 / 12.0;
      const double cxr = (h1 * (double)r.x - h2 * (double)a.x) + h3 * (double)b.x;
      const double cyi = (h1 * (double)r.y - h2 * (double)a.y) + h3 * (double)b.y;
      const double ox = (double)s.x + (double)dt * cxr, oy = (double)s.y + (double)dt * cyi;
      o = mk<T>((T)ox, (T)oy);
    }
    if (filter) {
      const T c = reinterpret_cast<const T*>(filter)[i];
      o = mk<T>(c * o.x, c * o.y);
    }
    reinterpret_cast<cx<T>*>(sol)[i] = o;
  }
};

// ---------------------------------------------------------------- dispatch on (T, CT, KIND)
// F(T, CT, KIND) must be a callable template struct; implemented with a macro to keep the switch in one place.
#define FFB_DISPATCH_TCK(dtype, cdtype, kind, ...)                                                    \
  do {                                                                                                 \
    if ((dtype) == FFB_F64) {                                                                          \
      FFB_REQUIRE((cdtype) == FFB_F64, FFB_EINVAL, "Float64 state needs Float64 coefficients");       \
      switch (kind) {                                                                                  \
        case FFB_COEF_SCALAR: { using T = double; using CT = double; constexpr int K = FFB_COEF_SCALAR; __VA_ARGS__ } break;   \
        case FFB_COEF_REAL: { using T = double; using CT = double; constexpr int K = FFB_COEF_REAL; __VA_ARGS__ } break;       \
        case FFB_COEF_COMPLEX: { using T = double; using CT = double; constexpr int K = FFB_COEF_COMPLEX; __VA_ARGS__ } break; \
        default: return set_error(FFB_EINVAL, "bad coefficient kind %d", (int)(kind));                 \
      }                                                                                                \
    } else if ((cdtype) == FFB_F64) {                                                                  \
      switch (kind) {                                                                                  \
        case FFB_COEF_SCALAR: { using T = float; using CT = double; constexpr int K = FFB_COEF_SCALAR; __VA_ARGS__ } break;    \
        case FFB_COEF_REAL: { using T = float; using CT = double; constexpr int K = FFB_COEF_REAL; __VA_ARGS__ } break;        \
        case FFB_COEF_COMPLEX: { using T = float; using CT = double; constexpr int K = FFB_COEF_COMPLEX; __VA_ARGS__ } break;  \
        default: return set_error(FFB_EINVAL, "bad coefficient kind %d", (int)(kind));                 \
      }                                                                                                \
    } else {                                                                                           \
      switch (kind) {                                                                                  \
        case FFB_COEF_SCALAR: { using T = float; using CT = float; constexpr int K = FFB_COEF_SCALAR; __VA_ARGS__ } break;     \
        case FFB_COEF_REAL: { using T = float; using CT = float; constexpr int K = FFB_COEF_REAL; __VA_ARGS__ } break;         \
        case FFB_COEF_COMPLEX: { using T = float; using CT = float; constexpr int K = FFB_COEF_COMPLEX; __VA_ARGS__ } break;   \
        default: return set_error(FFB_EINVAL, "bad coefficient kind %d", (int)(kind));                 \
      }                                                                                                \
    }                                                                                                  \
  } while (0)

static int check_coef(const ffb_coef* c, const char* name) {
  FFB_REQUIRE(c, FFB_EINVAL, "%s is NULL", name);
  FFB_REQUIRE(c->kind == FFB_COEF_SCALAR || c->ptr, FFB_EINVAL, "%s: dense coefficient without data", name);
  FFB_REQUIRE(c->dtype == FFB_F32 || c->dtype == FFB_F64, FFB_EINVAL, "%s: bad dtype", name);
  return FFB_OK;
}

}  // namespace ffb

using namespace ffb;

#define FFB_STAGE_PROLOG(dtype_, n_)                                                     \
  FFB_REQUIRE((dtype_) == FFB_F32 || (dtype_) == FFB_F64, FFB_EINVAL, "bad dtype %d", (int)(dtype_)); \
  FFB_REQUIRE((n_) >= 0, FFB_EINVAL, "negative element count")

extern "C" {

int ffb_stage_fe(void* sol, const void* N, const ffb_coef* L, double dt, const void* filter, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol && N, FFB_EINVAL, "NULL array");
  int rc = check_coef(L, "L"); if (rc) return rc;
  FFB_DISPATCH_TCK(dtype, L->dtype, L->kind, {
    OpFE<T, CT, K> op{sol, N, view<CT, K>(L), (CT)dt, filter};
    return launch<T, CT, K>(op, n, "stage_fe", 3, 1, filter != nullptr);
  });
  return FFB_OK;
}

int ffb_stage_rk4_substep(void* sol1, void* rhs, const void* u, const void* sol, const ffb_coef* L, double c, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol1 && rhs && u && sol, FFB_EINVAL, "NULL array");
  int rc = check_coef(L, "L"); if (rc) return rc;
  FFB_DISPATCH_TCK(dtype, L->dtype, L->kind, {
    OpRK4Sub<T, CT, K> op{sol1, rhs, u, sol, view<CT, K>(L), (CT)c};
    return launch<T, CT, K>(op, n, "stage_rk4_substep", u == sol ? 4 : 5, 1, false);
  });
  return FFB_OK;
}

int ffb_stage_rk4_final(void* sol, const void* rhs1, const void* rhs2, const void* rhs3, void* rhs4, const void* sol1,
                        const ffb_coef* L, double dt, const void* filter, int store_rhs4, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol && rhs1 && rhs2 && rhs3 && rhs4 && sol1, FFB_EINVAL, "NULL array");
  int rc = check_coef(L, "L"); if (rc) return rc;
  FFB_DISPATCH_TCK(dtype, L->dtype, L->kind, {
    // `dt/6` is formed in the clock's type T (src/timesteppers.jl:261)
    OpRK4Final<T, CT, K> op{sol, rhs1, rhs2, rhs3, rhs4, sol1, view<CT, K>(L), (CT)((T)dt / (T)6), filter, store_rhs4};
    return launch<T, CT, K>(op, n, "stage_rk4_final", store_rhs4 ? 8 : 7, 1, filter != nullptr);
  });
  return FFB_OK;
}

int ffb_stage_lsrk54(void* sol, void* S2, void* rhs, const ffb_coef* L, double A, double B, double dt, int first,
                     const void* filter, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol && S2 && rhs, FFB_EINVAL, "NULL array");
  int rc = check_coef(L, "L"); if (rc) return rc;
  FFB_DISPATCH_TCK(dtype, L->dtype, L->kind, {
    OpLSRK<T, CT, K> op{sol, S2, rhs, view<CT, K>(L), (CT)(T)A, (CT)(T)B, (CT)dt, first, filter};
    return launch<T, CT, K>(op, n, "stage_lsrk54", first ? 4 : 5, 1, filter != nullptr);
  });
  return FFB_OK;
}

int ffb_stage_etdrk4_substep12(void* out, const ffb_coef* exphLdt, const void* sol, const ffb_coef* zeta, const void* N,
                               int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(out && sol && N, FFB_EINVAL, "NULL array");
  int rc = check_coef(exphLdt, "exphLdt"); if (rc) return rc;
  rc = check_coef(zeta, "zeta"); if (rc) return rc;
  FFB_REQUIRE(exphLdt->kind == zeta->kind && exphLdt->dtype == zeta->dtype, FFB_EINVAL, "ETD coefficients must share kind and dtype");
  FFB_DISPATCH_TCK(dtype, zeta->dtype, zeta->kind, {
    OpETD12<T, CT, K> op{out, view<CT, K>(exphLdt), sol, view<CT, K>(zeta), N};
    return launch<T, CT, K>(op, n, "stage_etdrk4_substep12", 3, 2, false);
  });
  return FFB_OK;
}

int ffb_stage_etdrk4_substep3(void* out, const ffb_coef* exphLdt, const void* sol1, const ffb_coef* zeta, const void* N1,
                              const void* N3, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(out && sol1 && N1 && N3, FFB_EINVAL, "NULL array");
  int rc = check_coef(exphLdt, "exphLdt"); if (rc) return rc;
  rc = check_coef(zeta, "zeta"); if (rc) return rc;
  FFB_REQUIRE(exphLdt->kind == zeta->kind && exphLdt->dtype == zeta->dtype, FFB_EINVAL, "ETD coefficients must share kind and dtype");
  FFB_DISPATCH_TCK(dtype, zeta->dtype, zeta->kind, {
    OpETD3<T, CT, K> op{out, view<CT, K>(exphLdt), sol1, view<CT, K>(zeta), N1, N3};
    return launch<T, CT, K>(op, n, "stage_etdrk4_substep3", 4, 2, false);
  });
  return FFB_OK;
}

int ffb_stage_etdrk4_update(void* sol, const ffb_coef* expLdt, const ffb_coef* alpha, const ffb_coef* beta, const ffb_coef* gamma,
                            const void* N1, const void* N2, const void* N3, const void* N4, const void* filter, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol && N1 && N2 && N3 && N4, FFB_EINVAL, "NULL array");
  const ffb_coef* cs[4] = {expLdt, alpha, beta, gamma};
  for (int i = 0; i < 4; ++i) {
    int rc = check_coef(cs[i], "ETD coefficient"); if (rc) return rc;
    FFB_REQUIRE(cs[i]->kind == alpha->kind && cs[i]->dtype == alpha->dtype, FFB_EINVAL, "ETD coefficients must share kind and dtype");
  }
  FFB_DISPATCH_TCK(dtype, alpha->dtype, alpha->kind, {
    OpETDUpd<T, CT, K> op{sol, view<CT, K>(expLdt), view<CT, K>(alpha), view<CT, K>(beta), view<CT, K>(gamma), N1, N2, N3, N4, filter};
    return launch<T, CT, K>(op, n, "stage_etdrk4_update", 6, 4, filter != nullptr);
  });
  return FFB_OK;
}

int ffb_stage_ab3(void* sol, void* rhs, const void* rhs_m1, const void* rhs_m2, const ffb_coef* L, double dt, int64_t step,
                  const void* filter, int dtype, int64_t n) {
  FFB_STAGE_PROLOG(dtype, n);
  FFB_REQUIRE(sol && rhs && rhs_m1 && rhs_m2, FFB_EINVAL, "NULL array");
  int rc = check_coef(L, "L"); if (rc) return rc;
  FFB_REQUIRE((L->dtype == FFB_F64) == (dtype == FFB_F64), FFB_EINVAL, "L must have the state's precision");
  const int euler = step < 3;  // three forward-Euler steps (clock.step = 0, 1, 2), src/timesteppers.jl:629
  FFB_DISPATCH_TCK(dtype, L->dtype, L->kind, {
    (void)sizeof(CT);
    OpAB3<T, K> op{sol, rhs, rhs_m1, rhs_m2, view<T, K>(L), (T)dt, euler, filter};
    return launch<T, T, K>(op, n, "stage_ab3", euler ? 4 : 6, 1, filter != nullptr);
  });
  return FFB_OK;
}

}  // extern "C"
