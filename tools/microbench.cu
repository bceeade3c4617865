// Design-time microbenchmarks for the FFT pass layout decisions (not part of the product).
//   1. strided tile copy: rows of W bytes at a large row stride (the access shape of a y-pass)
//   2. cuFFT baseline (the existing Blackwell FFT the hand-written kernels must beat)
//   3. FP64 / FP32 FMA rate
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -lcufft -o tools/microbench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cufft.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// Each CTA copies a tile: `rows` rows of `wbytes` contiguous bytes, rows separated by `stride` bytes.
// Threads are laid out so that consecutive threads touch consecutive 16-byte chunks within a row.
__global__ void tile_copy(const uint4* __restrict__ in, uint4* __restrict__ out, int rows, int wchunks,
                          size_t stride_chunks, int tiles_per_row) {
  int tile = blockIdx.x;
  int tx = tile % tiles_per_row;
  size_t base = (size_t)tx * wchunks + (size_t)(tile / tiles_per_row) * rows * stride_chunks;
  for (int i = threadIdx.x; i < rows * wchunks; i += blockDim.x) {
    int r = i / wchunks, c = i % wchunks;
    size_t idx = base + (size_t)r * stride_chunks + c;
    out[idx] = in[idx];
  }
}

__global__ void fma64(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.000001, c = 0.5, d = 0.25, e = 0.125, f = 0.1, g = 0.2, h = 0.3;
  for (int i = 0; i < iters; i++) {
    a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c);
    g = fma(g, b, c); h = fma(h, b, c); a = fma(a, b, d); e = fma(e, b, f);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + g + h;
}
__global__ void fma32(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.000001f, c = 0.5f, d = 0.25f, e = 0.125f, f = 0.1f, g = 0.2f, h = 0.3f;
  for (int i = 0; i < iters; i++) {
    a = fmaf(a, b, c); d = fmaf(d, b, c); e = fmaf(e, b, c); f = fmaf(f, b, c);
    g = fmaf(g, b, c); h = fmaf(h, b, c); a = fmaf(a, b, d); e = fmaf(e, b, f);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + g + h;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs %d smem/blk optin %zu L2 %d\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin, prop.l2CacheSize);

  // ---- 1. strided tile copy, array = (4097 x 8192) 16-byte elements (8192^2 F64 spectral array) ----
  {
    size_t nkr = 4097, ny = 8192;
    size_t bytes = nkr * ny * 16;
    uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    // plain copy for reference
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    printf("memcpy d2d: %.1f GB/s\n", 2.0 * bytes / time_ms(e0, e1) / 1e6);
    int wlist[] = {1, 2, 4, 8, 16, 32};
    int rowslist[] = {8192, 1024, 128};
    for (int rows : rowslist) for (int w : wlist) {
      int tiles_per_row = (int)(nkr / w);  // ignore tail
      int ntiles = tiles_per_row * (int)(ny / rows);
      float best = 1e9;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        tile_copy<<<ntiles, 512>>>(a, b, rows, w, nkr, tiles_per_row);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      CK(cudaGetLastError());
      double moved = 2.0 * (double)tiles_per_row * w * 16 * ny;
      printf("tile_copy rows=%5d chunk=%4dB: %.3f ms  %.1f GB/s\n", rows, w * 16, best, moved / best / 1e6);
    }
    cudaFree(a); cudaFree(b);
  }
  // same for 8-byte-element array (1025 x 2048 x 256 slab of F32 complex), row stride 1025*8 bytes, z-pass-like and y-pass-like
  // ---- 2. cuFFT baselines ----
  {
    struct Case { const char* name; int rank; int n[3]; bool dbl; };
    Case cases[] = {
      {"D2Z 2D 4096^2", 2, {4096, 4096, 0}, true},
      {"D2Z 2D 8192^2", 2, {8192, 8192, 0}, true},
      {"R2C 2D 8192^2", 2, {8192, 8192, 0}, false},
      {"R2C 3D 512^3", 3, {512, 512, 512}, false},
      {"D2Z 3D 512^3", 3, {512, 512, 512}, true},
      {"R2C 3D 1024^3", 3, {1024, 1024, 1024}, false},
      {"D2Z 3D 1024^3", 3, {1024, 1024, 1024}, true},
    };
    for (auto& c : cases) {
      size_t npts = 1; for (int i = 0; i < c.rank; i++) npts *= c.n[i];
      size_t nspec = (size_t)(c.n[c.rank - 1] / 2 + 1); for (int i = 0; i < c.rank - 1; i++) nspec *= c.n[i];
      size_t es = c.dbl ? 8 : 4;
      void *re, *cx; CK(cudaMalloc(&re, npts * es)); CK(cudaMalloc(&cx, nspec * es * 2));
      CK(cudaMemset(re, 0, npts * es));
      cufftHandle pf, pi; size_t ws = 0;
      // cuFFT is row-major: last dim fastest -> same memory layout as column-major (nx fastest) with dims reversed
      cufftResult r1 = cufftCreate(&pf); cufftResult r2 = cufftMakePlanMany(pf, c.rank, c.n, nullptr, 1, 0, nullptr, 1, 0, c.dbl ? CUFFT_D2Z : CUFFT_R2C, 1, &ws);
      size_t ws2 = 0;
      cufftCreate(&pi); cufftResult r3 = cufftMakePlanMany(pi, c.rank, c.n, nullptr, 1, 0, nullptr, 1, 0, c.dbl ? CUFFT_Z2D : CUFFT_C2R, 1, &ws2);
      if (r1 || r2 || r3) { printf("cufft %s: plan failed %d %d %d\n", c.name, r1, r2, r3); cudaFree(re); cudaFree(cx); continue; }
      float bf = 1e9, bi = 1e9;
      for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (c.dbl) cufftExecD2Z(pf, (double*)re, (cufftDoubleComplex*)cx); else cufftExecR2C(pf, (float*)re, (cufftComplex*)cx);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = time_ms(e0, e1); if (rep > 0 && ms < bf) bf = ms;
        cudaEventRecord(e0);
        if (c.dbl) cufftExecZ2D(pi, (cufftDoubleComplex*)cx, (double*)re); else cufftExecC2R(pi, (cufftComplex*)cx, (float*)re);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        ms = time_ms(e0, e1); if (rep > 0 && ms < bi) bi = ms;
      }
      double P = (double)npts * es, S = (double)nspec * es * 2;
      double alg = P + (2 * c.rank - 1) * S;
      printf("cufft %-16s fwd %.3f ms (%.0f GB/s alg) inv %.3f ms (%.0f GB/s alg) workspace %.1f/%.1f MB\n", c.name, bf, alg / bf / 1e6, bi, alg / bi / 1e6, ws / 1e6, ws2 / 1e6);
      cufftDestroy(pf); cufftDestroy(pi); cudaFree(re); cudaFree(cx);
    }
  }
  // ---- 3. FMA rates ----
  {
    double* o; CK(cudaMalloc(&o, 148 * 8 * 1024 * 8));
    int iters = 20000;
    fma64<<<148 * 8, 1024>>>(o, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); fma64<<<148 * 8, 1024>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    double fl = 148.0 * 8 * 1024 * iters * 8 * 2;
    printf("fp64 fma: %.2f TFLOP/s\n", fl / time_ms(e0, e1) / 1e9);
    fma32<<<148 * 8, 1024>>>((float*)o, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); fma32<<<148 * 8, 1024>>>((float*)o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    printf("fp32 fma: %.2f TFLOP/s\n", fl / time_ms(e0, e1) / 1e9);
    cudaFree(o);
  }
  return 0;
}
