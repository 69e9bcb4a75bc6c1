// Microbenchmark 2 (design input for K1 and K2 on sm_100a):
//  (a) single-copy k-mer histograms with one ATOMS per S windows (k = 3+S, S = 1..4), per warp / per CTA
//  (b) FP32 |a-b| accumulate: scalar FADD pairs vs packed FADD2 pairs (register operands only)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ---------------- (a) ----------------
template <int S, bool PER_WARP>
__global__ void __launch_bounds__(1024) kmer_hist(int iters, unsigned long long* cyc, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  constexpr int KB = 2 * (3 + S);              // bits per k-mer
  constexpr int BINS = 1 << KB;
  const int warp = threadIdx.x >> 5;
  const int total = PER_WARP ? BINS * (blockDim.x / 32) : BINS;
  for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t* h = PER_WARP ? sm + warp * BINS : sm;
  uint32_t x = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
  uint32_t cur = x * 1664525u + 1013904223u;
  unsigned long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    // 12 words = 192 bases = lcm-friendly for S = 1,2,3,4
    #pragma unroll
    for (int wd = 0; wd < 12; ++wd) {
      x = x * 1664525u + 1013904223u;
      uint32_t nxt = x ^ (x >> 13);
      #pragma unroll
      for (int p = 0; p < 16; ++p) {
        if (((wd * 16 + p) % S) == 0) {
          uint32_t f = __funnelshift_l(nxt, cur, 2 * p);          // k-mer starting at base p in the top bits
          uint32_t off = (f >> (32 - KB - 2)) & ((BINS - 1) << 2);
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        }
      }
      cur = nxt;
    }
  }
  unsigned long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s += sm[i];
  if (s == 0xdeadbeef) sink[0] = s;
}

template <int S, bool PER_WARP>
void run_hist(const char* name, int threads, int cps_req, int iters) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  size_t bins = (size_t)1 << (2 * (3 + S));
  size_t smem = PER_WARP ? bins * 4 * (threads / 32) : bins * 4;
  if (smem > 227 * 1024) { printf("%-34s too much smem\n", name); return; }
  CK(cudaFuncSetAttribute(kmer_hist<S, PER_WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kmer_hist<S, PER_WARP>, threads, smem));
  if (occ < 1) { printf("%-34s cannot launch\n", name); return; }
  int cps = occ < cps_req ? occ : cps_req;
  int grid = nsm * cps;
  unsigned long long* cyc; uint32_t* sink; CK(cudaMalloc(&cyc, grid * 8)); CK(cudaMalloc(&sink, 4));
  kmer_hist<S, PER_WARP><<<grid, threads, smem>>>(iters / 8 + 1, cyc, sink); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kmer_hist<S, PER_WARP><<<grid, threads, smem>>>(iters, cyc, sink);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = (unsigned long long*)malloc(grid * 8);
  CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double per_sm = (double)iters * 192 * threads * cps;
  printf("%-34s S=%d thr=%4d cta/SM=%2d warps/SM=%2d smem/CTA=%6zu  windows/clk/SM %.2f  (%.1f Gwin/s chip)\n", name, S, threads, cps,
         threads / 32 * cps, smem, per_sm / mx, per_sm * nsm / (ms * 1e6));
  cudaFree(cyc); cudaFree(sink); free(h);
}

// ---------------- (b) ----------------
template <int MODE>
__global__ void __launch_bounds__(256) fp_l1(int iters, float* out, unsigned long long* cyc, float seed) {
  float a[8], b[8];
  #pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i); b[i] = seed * (blockIdx.x + 3 * i); }
  unsigned long long t0 = clock64();
  if (MODE == 0) {            // scalar: d = a - b ; acc += |d|     (8x8 = 64 accumulators)
    float acc[8][8];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
          #pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += fabsf(a[i] - b[j]);
        #pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 0.25f; }      // 8 extra FADD per 128: keeps operands live
      }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) {     // packed along j: (a_i,a_i) - (b_j,b_j+1) ; acc2 += |d2|   (8x4 pair accumulators)
    float2 acc[8][4], a2[8], b2[4];
    #pragma unroll
    for (int i = 0; i < 8; ++i) { a2[i] = make_float2(a[i], a[i]);
      #pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f); }
    #pragma unroll
    for (int j = 0; j < 4; ++j) b2[j] = make_float2(-b[2 * j], -b[2 * j + 1]);
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
          #pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 d = __fadd2_rn(a2[i], b2[j]);
            acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fabsf(d.x), fabsf(d.y)));
          }
        #pragma unroll
        for (int i = 0; i < 8; ++i) a2[i] = __fadd2_rn(a2[i], make_float2(0.25f, 0.25f));
      }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 2) {     // mixed: scalar subs + packed abs-add (3 issue slots per 2 pair-dims)
    float2 acc[8][4];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
          #pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fabsf(a[i] - b[2 * j]), fabsf(a[i] - b[2 * j + 1])));
        #pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += 0.25f;
      }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {                    // max-form: acc += max(a,b)   (FMNMX on the ALU pipe + FADD)
    float acc[8][8];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
          #pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += fmaxf(a[i], b[j]);
        #pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += 0.25f;
      }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run_fp(const char* name, int threads, int cps, int iters) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int grid = nsm * cps;
  float* out; unsigned long long* cyc; CK(cudaMalloc(&out, (size_t)grid * threads * 4)); CK(cudaMalloc(&cyc, grid * 8));
  fp_l1<MODE><<<grid, threads>>>(iters / 8 + 1, out, cyc, 0.001f); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  fp_l1<MODE><<<grid, threads>>>(iters, out, cyc, 0.001f);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = (unsigned long long*)malloc(grid * 8);
  CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double pairdims_per_sm = (double)iters * 4 * 64 * threads * cps;
  printf("%-30s thr=%3d cta/SM=%d  pair-dims/clk/SM %.1f (=%.1f lane-ops/clk/SM at 2 per pair-dim)  %.1f G pairs/s chip @136 dims, %.3f ms\n", name, threads, cps,
         pairdims_per_sm / mx, 2 * pairdims_per_sm / mx, pairdims_per_sm * nsm / 136.0 / (ms * 1e6), ms);
  cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
  int it = 600;
  run_hist<1, true>("per-warp single copy", 256, 8, it);
  run_hist<1, true>("per-warp single copy", 512, 4, it);
  run_hist<2, true>("per-warp single copy", 256, 6, it);
  run_hist<2, true>("per-warp single copy", 512, 3, it);
  run_hist<2, true>("per-warp single copy", 256, 3, it);
  run_hist<3, true>("per-warp single copy", 128, 3, it);
  run_hist<3, true>("per-warp single copy", 256, 1, it);
  run_hist<3, true>("per-warp single copy", 416, 1, it);
  run_hist<1, false>("per-CTA single copy", 1024, 2, it);
  run_hist<2, false>("per-CTA single copy", 1024, 2, it);
  run_hist<2, false>("per-CTA single copy", 256, 8, it);
  run_hist<3, false>("per-CTA single copy", 1024, 2, it);
  run_hist<3, false>("per-CTA single copy", 512, 4, it);
  run_hist<3, false>("per-CTA single copy", 256, 8, it);
  run_hist<3, false>("per-CTA single copy", 256, 4, it);
  run_hist<4, false>("per-CTA single copy", 1024, 2, it);
  run_hist<4, false>("per-CTA single copy", 512, 3, it);
  run_hist<4, false>("per-CTA single copy", 1024, 1, it);
  int fi = 2000;
  run_fp<0>("scalar FADD,FADD|.|", 256, 1, fi);
  run_fp<0>("scalar FADD,FADD|.|", 256, 2, fi);
  run_fp<1>("packed FADD2,FADD2|.|", 256, 1, fi);
  run_fp<1>("packed FADD2,FADD2|.|", 256, 2, fi);
  run_fp<2>("2xFADD + FADD2|.|", 256, 1, fi);
  run_fp<2>("2xFADD + FADD2|.|", 256, 2, fi);
  run_fp<3>("FMNMX + FADD (max form)", 256, 1, fi);
  run_fp<3>("FMNMX + FADD (max form)", 256, 2, fi);
  return 0;
}
