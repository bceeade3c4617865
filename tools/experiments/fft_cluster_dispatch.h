// Entry points exported by fft_cluster_inst.cu (one object per real type).  Return FFB_OK, a negative ffb_status, or 1 when
// the length has no cluster kernel.
#pragma once
namespace ffb {
inline bool cluster_has(int N) { return N == 1024 || N == 2048 || N == 4096 || N == 8192; }
inline void cluster_split(int N, int* N1, int* N2) {
  switch (N) {
    case 1024: *N1 = 32; *N2 = 32; break;
    case 2048: *N1 = 32; *N2 = 64; break;
    case 4096: *N1 = 64; *N2 = 64; break;
    default: *N1 = 64; *N2 = 128; break;
  }
}
inline int cluster_cols(int real_bytes) { return real_bytes == 8 ? 4 : 8; }
}  // namespace ffb
int cluster_launch_float(int N, int dir, const void* params, long long ntiles, int nslices, void* stream);
int cluster_launch_double(int N, int dir, const void* params, long long ntiles, int nslices, void* stream);
