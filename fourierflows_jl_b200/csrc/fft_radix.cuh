// Register butterflies (radix 2/4/8/16) for the Stockham passes.  Natural-order in, natural-order out:
//   y[k] = sum_q x[q] * exp(DIR * 2*pi*i * q*k / r),   DIR = -1 forward, +1 inverse.
// All functions work on a strided view of the thread's register array so that one array of R points can hold
// R/r independent butterflies; every index is a compile-time constant after unrolling (no local memory).
#pragma once
#include "ffb_common.cuh"

namespace ffb {

template <typename T> struct consts;
template <> struct consts<double> {
  static constexpr double rsqrt2 = 0.70710678118654752440084436210485;
  static constexpr double c1_16 = 0.92387953251128675612818318939679;  // cos(pi/8)
  static constexpr double s1_16 = 0.38268343236508977172845998403040;  // sin(pi/8)
};
template <> struct consts<float> {
  static constexpr float rsqrt2 = 0.70710678118654752440084436210485f;
  static constexpr float c1_16 = 0.92387953251128675612818318939679f;
  static constexpr float s1_16 = 0.38268343236508977172845998403040f;
};

// multiply by exp(DIR * i*pi/2) = DIR * i
template <int DIR, typename T> FFB_HD cx<T> rot90(cx<T> a) { return DIR < 0 ? mul_mi(a) : mul_i(a); }
// multiply by exp(DIR * i*pi/4) = (1 + DIR*i)/sqrt2
template <int DIR, typename T> FFB_HD cx<T> rot45(cx<T> a) {
  const T h = consts<T>::rsqrt2;
  return DIR < 0 ? mk<T>(h * (a.x + a.y), h * (a.y - a.x)) : mk<T>(h * (a.x - a.y), h * (a.y + a.x));
}
// multiply by exp(DIR * i*3pi/4) = (-1 + DIR*i)/sqrt2
template <int DIR, typename T> FFB_HD cx<T> rot135(cx<T> a) {
  const T h = consts<T>::rsqrt2;
  return DIR < 0 ? mk<T>(h * (a.y - a.x), -h * (a.x + a.y)) : mk<T>(-h * (a.x + a.y), h * (a.x - a.y));
}
// multiply by exp(DIR * i * m*pi/8) for compile-time m in [0, 16)
template <int DIR, int M, typename T> FFB_HD cx<T> rot16(cx<T> a) {
  constexpr int m = M & 15;
  if constexpr (m == 0) return a;
  else if constexpr (m == 2) return rot45<DIR>(a);
  else if constexpr (m == 4) return rot90<DIR>(a);
  else if constexpr (m == 6) return rot135<DIR>(a);
  else if constexpr (m == 8) return mk<T>(-a.x, -a.y);
  else if constexpr (m > 8) { cx<T> b = rot16<DIR, m - 8>(a); return mk<T>(-b.x, -b.y); }
  else {
    // m in {1,3,5,7}: w = cos(m pi/8) + DIR*i*sin(m pi/8)
    const T c = (m == 1) ? consts<T>::c1_16 : (m == 3) ? consts<T>::s1_16 : (m == 5) ? -consts<T>::s1_16 : -consts<T>::c1_16;
    const T s0 = (m == 1) ? consts<T>::s1_16 : (m == 3) ? consts<T>::c1_16 : (m == 5) ? consts<T>::c1_16 : consts<T>::s1_16;
    const T s = DIR < 0 ? -s0 : s0;
    return mk<T>(c * a.x - s * a.y, c * a.y + s * a.x);
  }
}

template <int DIR, typename T> FFB_HD void bfly2(cx<T>& a, cx<T>& b) {
  cx<T> t = a - b;
  a = a + b;
  b = t;
}

template <int DIR, typename T> FFB_HD void bfly4(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3) {
  cx<T> t0 = x0 + x2, t1 = x0 - x2, t2 = x1 + x3, t3 = rot90<DIR>(x1 - x3);
  x0 = t0 + t2;
  x1 = t1 + t3;
  x2 = t0 - t2;
  x3 = t1 - t3;
}

// radix-8 = 2 (n2) x 4 (n1):  n = 2*n1 + n2,  k = k1 + 4*k2
template <int DIR, typename T>
FFB_HD void bfly8(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3, cx<T>& x4, cx<T>& x5, cx<T>& x6, cx<T>& x7) {
  // DFT4 over n1 for n2 = 0 (x0,x2,x4,x6) and n2 = 1 (x1,x3,x5,x7)
  bfly4<DIR>(x0, x2, x4, x6);
  bfly4<DIR>(x1, x3, x5, x7);
  // twiddle W8^(n2*k1) on the n2 = 1 set: k1 = 0..3 -> 1, W8, W8^2, W8^3
  x3 = rot45<DIR>(x3);
  x5 = rot90<DIR>(x5);
  x7 = rot135<DIR>(x7);
  // DFT2 over n2: X[k1 + 4*k2];  A[n2=0][k1] = x(2*k1), A[n2=1][k1] = x(2*k1+1)
  cx<T> y0 = x0 + x1, y4 = x0 - x1;
  cx<T> y1 = x2 + x3, y5 = x2 - x3;
  cx<T> y2 = x4 + x5, y6 = x4 - x5;
  cx<T> y3 = x6 + x7, y7 = x6 - x7;
  x0 = y0; x1 = y1; x2 = y2; x3 = y3; x4 = y4; x5 = y5; x6 = y6; x7 = y7;
}

// radix-16 = 4 (n2) x 4 (n1):  n = 4*n1 + n2,  k = k1 + 4*k2
template <int DIR, typename T> FFB_HD void bfly16(cx<T> (&x)[16]) {
  // step 1: DFT4 over n1 for each n2 -> A[n2][k1] stored at x[4*k1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) bfly4<DIR>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);
  // step 2: twiddle W16^(n2*k1)
  x[4 * 1 + 1] = rot16<DIR, 1>(x[4 * 1 + 1]);
  x[4 * 1 + 2] = rot16<DIR, 2>(x[4 * 1 + 2]);
  x[4 * 1 + 3] = rot16<DIR, 3>(x[4 * 1 + 3]);
  x[4 * 2 + 1] = rot16<DIR, 2>(x[4 * 2 + 1]);
  x[4 * 2 + 2] = rot16<DIR, 4>(x[4 * 2 + 2]);
  x[4 * 2 + 3] = rot16<DIR, 6>(x[4 * 2 + 3]);
  x[4 * 3 + 1] = rot16<DIR, 3>(x[4 * 3 + 1]);
  x[4 * 3 + 2] = rot16<DIR, 6>(x[4 * 3 + 2]);
  x[4 * 3 + 3] = rot16<DIR, 9>(x[4 * 3 + 3]);
  // step 3: DFT4 over n2 for each k1 -> X[k1 + 4*k2] lands at x[4*k1 + k2]; transpose to natural order
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) bfly4<DIR>(x[4 * k1 + 0], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
  cx<T> y[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) y[k1 + 4 * k2] = x[4 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = y[i];
}

// Butterfly number B of radix r on the strided view v[B + k*(R/r)], k = 0..r-1.
template <int DIR, int R, int r, int B, typename T> FFB_HD void bfly_at(cx<T> (&v)[R]) {
  constexpr int s = R / r;
  if constexpr (r == 2) bfly2<DIR>(v[B], v[B + s]);
  else if constexpr (r == 4) bfly4<DIR>(v[B], v[B + s], v[B + 2 * s], v[B + 3 * s]);
  else if constexpr (r == 8) bfly8<DIR>(v[B], v[B + s], v[B + 2 * s], v[B + 3 * s], v[B + 4 * s], v[B + 5 * s], v[B + 6 * s], v[B + 7 * s]);
  else if constexpr (r == 16) {
    cx<T> y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = v[B + k * s];
    bfly16<DIR>(y);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[B + k * s] = y[k];
  } else if constexpr (r == 1) {
  } else {
    static_assert(r == 2 || r == 4 || r == 8 || r == 16 || r == 1, "unsupported radix");
  }
}

}  // namespace ffb
