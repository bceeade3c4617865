// Host-side check of the register butterflies against a naive DFT (runs on the CPU; built by tests/test_native_host.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../fourierflows_jl_b200/csrc/fft_radix.cuh"
using namespace ffb;

template <int DIR, int R, int r, int B> struct Run {
  static void go(cx<double> (&v)[R]) { bfly_at<DIR, R, r, B>(v); if constexpr (B + 1 < R / r) Run<DIR, R, r, B + 1>::go(v); }
};

template <int DIR, int R, int r> double check() {
  cx<double> v[R], ref[R];
  for (int i = 0; i < R; ++i) v[i] = mk<double>(std::sin(1.0 + 3 * i) + 0.1 * i, std::cos(2.0 + 5 * i) - 0.05 * i);
  const int s = R / r;
  for (int b = 0; b < s; ++b)
    for (int k = 0; k < r; ++k) {
      double re = 0, im = 0;
      for (int q = 0; q < r; ++q) {
        double ang = DIR * 2.0 * M_PI * q * k / r;
        cx<double> x = v[b + q * s];
        re += x.x * std::cos(ang) - x.y * std::sin(ang);
        im += x.x * std::sin(ang) + x.y * std::cos(ang);
      }
      ref[b + k * s] = mk<double>(re, im);
    }
  Run<DIR, R, r, 0>::go(v);
  double err = 0, nrm = 0;
  for (int i = 0; i < R; ++i) { err += std::pow(v[i].x - ref[i].x, 2) + std::pow(v[i].y - ref[i].y, 2); nrm += ref[i].x * ref[i].x + ref[i].y * ref[i].y; }
  return std::sqrt(err / nrm);
}

int main() {
  double worst = 0;
#define CHK(R, r) { double e1 = check<-1, R, r>(), e2 = check<1, R, r>(); printf("R=%d r=%d fwd %.2e inv %.2e\n", R, r, e1, e2); worst = std::fmax(worst, std::fmax(e1, e2)); }
  CHK(16, 16) CHK(16, 8) CHK(16, 4) CHK(16, 2) CHK(8, 8) CHK(8, 4) CHK(8, 2) CHK(4, 4) CHK(4, 2) CHK(2, 2) CHK(32, 16) CHK(32, 8)
  printf("worst %.3e\n", worst);
  return worst < 1e-14 ? 0 : 1;
}
