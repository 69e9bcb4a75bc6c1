// Microbenchmark 3: can the ALU pipe (FMNMX) take part of the L1 work off the FMA pipe (FADD2)?
//   sum|a-b| = sum(a) + sum(b) - 2 * sum(min(a,b))
// variants per (i,j) register-tile entry and per two dimensions:
//   D: direct   d2 = a2 - b2 ; acc2 += |d2|           -> 2 FADD2                 (FMA pipe only)
//   M: min-form m2 = min(a2, b2) ; acc2 += m2          -> 2 FMNMX + 1 FADD2       (ALU + FMA)
// MIX = number of the 8 columns j that use the min form.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MIX>
__global__ void __launch_bounds__(256, 1) k(int iters, float* out, unsigned long long* cyc, float seed) {
  float2 a2[8], b2[8], acc[8][8];
  #pragma unroll
  for (int i = 0; i < 8; ++i) {
    a2[i] = make_float2(seed * (threadIdx.x + i), seed * (threadIdx.x + 2 * i + 1));
    b2[i] = make_float2(seed * (blockIdx.x + 3 * i), seed * (blockIdx.x + 5 * i + 2));
    #pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
  }
  unsigned long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    #pragma unroll
    for (int j = 0; j < 8; ++j) {
      #pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (j < MIX) {
          acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fminf(a2[i].x, b2[j].x), fminf(a2[i].y, b2[j].y)));
        } else {
          float2 d = __fadd2_rn(a2[i], make_float2(-b2[j].x, -b2[j].y));
          acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fabsf(d.x), fabsf(d.y)));
        }
      }
    }
    #pragma unroll
    for (int i = 0; i < 8; ++i) { a2[i] = __fadd2_rn(a2[i], make_float2(0.25f, 0.125f)); }
  }
  unsigned long long t1 = clock64();
  float s = 0;
  #pragma unroll
  for (int i = 0; i < 8; ++i)
    #pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MIX>
void run(int iters, int threads = 256) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  float* out; unsigned long long* cyc; CK(cudaMalloc(&out, (size_t)nsm * 256 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
  k<MIX><<<nsm, threads>>>(iters / 8 + 1, out, cyc, 0.001f); CK(cudaDeviceSynchronize());
  k<MIX><<<nsm, threads>>>(iters, out, cyc, 0.001f); CK(cudaDeviceSynchronize());
  unsigned long long* h = (unsigned long long*)malloc(nsm * 8);
  CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int i = 0; i < nsm; ++i) if (h[i] > mx) mx = h[i];
  double pairdims = (double)iters * 64 * 2 * threads;      // per SM
  printf("threads/SM %d, min-form columns %d/8: %.1f pair-dims/clk/SM  (direct-only peak is 64)\n", threads, MIX, pairdims / mx);
  cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
  run<0>(4000, 128); run<0>(4000, 256); run<0>(4000, 512); run<2>(4000); run<4>(4000); run<5>(4000); run<6>(4000); run<8>(4000);
  return 0;
}
