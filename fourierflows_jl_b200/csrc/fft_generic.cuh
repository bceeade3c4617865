// Arbitrary-length FFT passes (mixed radix Stockham through global memory).  Correctness path for the sizes the
// reference's tests use that are not powers of two (6, 10, 30, 34, ... test/runtests.jl:49-51,166; test_grid.jl:204-206)
// and fallback for lines too long for the register-resident kernels.  One thread computes one output element.
#pragma once
#include <vector>
#include "ffb_common.cuh"

namespace ffb {

// Dense geometry of a 1-D pass: element (i_in, i, i_out) lives at i_in + inner*(i + N*i_out).
template <typename T, int DIR>
__global__ void generic_pass_kernel(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst, long long inner, int N,
                                    long long total, int Ns, int r, const cx<T>* __restrict__ wN, T scale) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const long long i_in = gid % inner;
  const long long rest = gid / inner;
  const int o = (int)(rest % N);
  const long long i_out = rest / N;
  const int a = o % Ns;
  const int k = (o / Ns) % r;
  const int jhi = o / (Ns * r);
  const int j = jhi * Ns + a;
  const int M = N / r;
  // exponent step per q: a*N/(Ns*r) + k*N/r   (mod N)
  const long long step = ((long long)a * (N / (Ns * r)) + (long long)k * M) % N;
  const cx<T>* base = src + i_in + inner * ((long long)N * i_out);
  T accx = 0, accy = 0;
  long long e = 0;
  for (int q = 0; q < r; ++q) {
    const cx<T> x = base[inner * (long long)(j + q * M)];
    cx<T> w = wN[e];
    if (DIR > 0) w.y = -w.y;
    accx += x.x * w.x - x.y * w.y;
    accy += x.x * w.y + x.y * w.x;
    e += step;
    if (e >= N) e -= N;
  }
  dst[i_in + inner * ((long long)o + (long long)N * i_out)] = mk<T>(accx * scale, accy * scale);
}

// X[k] (k = 0..N) from Z = FFT_N(pairs); src rows of N complex, dst rows of N+1 complex.
template <typename T>
__global__ void generic_r2c_post_kernel(const cx<T>* __restrict__ z, cx<T>* __restrict__ out, int N, long long nlines,
                                        const cx<T>* __restrict__ w2N) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nlines * (N + 1)) return;
  const long long line = gid / (N + 1);
  const int k = (int)(gid % (N + 1));
  const cx<T> zk = z[line * N + (k % N)];
  const cx<T> zc = conj(z[line * N + ((N - k) % N)]);
  const cx<T> w = w2N[k];  // exp(-i*pi*k/N), k = 0..N
  const cx<T> s = zk + zc, d = zk - zc;
  out[line * (N + 1) + k] = T(0.5) * (s + mul_mi(w * d));
}

// Z[k] (k = 0..N-1) from the half spectrum X[0..N]; src rows of N+1, dst rows of N.
template <typename T>
__global__ void generic_c2r_pre_kernel(const cx<T>* __restrict__ x, cx<T>* __restrict__ z, int N, long long nlines,
                                       const cx<T>* __restrict__ w2N) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nlines * N) return;
  const long long line = gid / N;
  const int k = (int)(gid % N);
  cx<T> xk = x[line * (N + 1) + k];
  cx<T> xc = conj(x[line * (N + 1) + (N - k)]);
  if (k == 0) { xk.y = T(0); xc.y = T(0); }  // c2r ignores Im X[0] and Im X[N] (FFTW / cuFFT / pocketfft convention)
  const cx<T> w = conj(w2N[k]);
  const cx<T> s = xk + xc, d = xk - xc;
  z[line * N + k] = s + mul_i(w * d);
}

inline std::vector<int> generic_factors(int N) {
  std::vector<int> f;
  int n = N;
  while (n % 16 == 0) { f.push_back(16); n /= 16; }
  while (n % 8 == 0) { f.push_back(8); n /= 8; }
  while (n % 4 == 0) { f.push_back(4); n /= 4; }
  while (n % 2 == 0) { f.push_back(2); n /= 2; }
  for (int p = 3; (long long)p * p <= n; p += 2)
    while (n % p == 0) { f.push_back(p); n /= p; }
  if (n > 1) f.push_back(n);
  return f;
}

// FFT of length N along the middle axis of a dense (inner, N, outer) complex array: src -> dst.
// tmpA/tmpB are scratch arrays of the same size; src may equal dst; dst must differ from tmpA/tmpB.
template <typename T>
int generic_fft_axis(const cx<T>* src, cx<T>* dst, cx<T>* tmpA, cx<T>* tmpB, long long inner, int N, long long outer,
                     int dir, T scale, const cx<T>* wN, cudaStream_t st) {
  const long long total = inner * N * outer;
  if (total == 0) return FFB_OK;
  std::vector<int> f = generic_factors(N);
  if (f.empty()) {  // N == 1
    if (src != dst || scale != T(1)) {
      f.push_back(1);
    } else {
      return FFB_OK;
    }
  }
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  FFB_REQUIRE(blocks < (1ll << 31), FFB_EUNSUPPORTED, "generic FFT pass too large");
  const cx<T>* cur = src;
  int Ns = 1;
  for (size_t i = 0; i < f.size(); ++i) {
    const bool last = (i + 1 == f.size());
    cx<T>* target;
    bool copy_back = false;
    if (last) {
      target = dst;
      if (cur == dst) { target = tmpA; copy_back = true; }  // single pass in place
    } else {
      target = (cur == tmpA) ? tmpB : tmpA;
    }
    const T sc = last ? scale : T(1);
    if (dir < 0) generic_pass_kernel<T, -1><<<(unsigned)blocks, threads, 0, st>>>(cur, target, inner, N, total, Ns, f[i], wN, sc);
    else generic_pass_kernel<T, 1><<<(unsigned)blocks, threads, 0, st>>>(cur, target, inner, N, total, Ns, f[i], wN, sc);
    count_launch();
    FFB_CHECK_LAUNCH();
    if (copy_back) FFB_CUDA(cudaMemcpyAsync(dst, tmpA, sizeof(cx<T>) * total, cudaMemcpyDeviceToDevice, st));
    cur = target;
    Ns *= f[i];
  }
  return FFB_OK;
}

}  // namespace ffb
