// Entry points exported by fft_stream_inst.cu (one object per real type).  Return FFB_OK, a negative ffb_status, or 1 when
// the length has no streaming kernel.
#pragma once
#include <cstddef>
namespace ffb {
inline bool stream_has(int N) { return N == 256 || N == 512 || N == 1024 || N == 2048; }
}  // namespace ffb
int stream_launch_float(int N, int dir, int split, const void* params, int threads, size_t smem, void* stream);
int stream_launch_double(int N, int dir, int split, const void* params, int threads, size_t smem, void* stream);
