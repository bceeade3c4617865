// Common device/host helpers for libfourierflows_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/fourierflows_b200.h"

#define FFB_HD __host__ __device__ __forceinline__
#define FFB_D __device__ __forceinline__

namespace ffb {

// ---- complex value type (layout-compatible with float2 / double2 and Julia's Complex{T}) ----
template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <typename T> FFB_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> FFB_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> FFB_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> FFB_HD cx<T> operator*(cx<T> a, cx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename T> FFB_HD cx<T> operator*(T s, cx<T> a) { return mk<T>(s * a.x, s * a.y); }
template <typename T> FFB_HD cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by +i / -i
template <typename T> FFB_HD cx<T> mul_i(cx<T> a) { return mk<T>(-a.y, a.x); }
template <typename T> FFB_HD cx<T> mul_mi(cx<T> a) { return mk<T>(a.y, -a.x); }

// ---- error plumbing (thread-local message, no exceptions across the ABI) ----
int set_error(int code, const char* fmt, ...);
cudaStream_t current_stream();
void count_launch(int n = 1);  // bookkeeping behind ffb_launch_count()

// Optional per-kernel timing with CUDA events on the launching stream (ffb_prof_enable / ffb_prof_report).
// `bytes` is the ALGORITHMIC HBM traffic of the launch, used for the roofline figures of bench.py.
bool prof_on();
void prof_push(const char* name, double bytes);
void prof_pop();
struct ProfScope {
  bool on;
  ProfScope(const char* name, double bytes) : on(prof_on()) { if (on) prof_push(name, bytes); }
  ~ProfScope() { if (on) prof_pop(); }
};

#define FFB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ffb::set_error(FFB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
  } while (0)

#define FFB_CHECK_LAUNCH()                                                                      \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return ffb::set_error(FFB_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                \
  } while (0)

#define FFB_REQUIRE(cond, code, ...)                       \
  do {                                                     \
    if (!(cond)) return ffb::set_error(code, __VA_ARGS__); \
  } while (0)

static inline int ilog2(uint64_t v) { int l = 0; while ((1ull << l) < v) ++l; return l; }
static inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }
static inline size_t dtype_size(int dtype) { return dtype == FFB_F64 ? 8 : 4; }

int num_sms();
int max_smem_optin();

}  // namespace ffb
