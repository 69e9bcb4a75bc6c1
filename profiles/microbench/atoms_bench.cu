// Microbenchmark: shared-memory histogram update throughput on sm_100a, to choose the K1 design.
// Each variant performs the same number of window updates per thread with LCG-generated codes
// (4 codes per 32-bit word, extracted with PRMT exactly as the real kernel does).
// Reports updates / clk / SM from in-kernel clock64 (max over CTAs) and from CUDA events.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}

enum Mode { ATOM_LANE32 = 0, ATOM_SINGLE = 1, ATOM_COPY8 = 2, ATOM_COPY16 = 3, RMW_LANE32 = 4,
            ATOM_CTA_LANE32 = 5, ATOM_5MER_CTA = 6, ATOM_COPY4 = 7, ATOM_LANE32_PRED = 8 };

// smem: per warp region of `warp_words` u32 (or per CTA for CTA modes)
template <int MODE>
__global__ void __launch_bounds__(1024) bench(int iters, unsigned long long* cyc, uint32_t* sink, int warp_words) {
  extern __shared__ uint32_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nthr = blockDim.x;
  int total_words = (MODE == ATOM_CTA_LANE32) ? 256 * 32 : (MODE == ATOM_5MER_CTA) ? 1024 * 32 : warp_words * (nthr / 32);
  for (int i = threadIdx.x; i < total_words; i += nthr) sm[i] = 0;
  __syncthreads();
  uint32_t* h = (MODE == ATOM_CTA_LANE32 || MODE == ATOM_5MER_CTA) ? sm : sm + warp * warp_words;
  uint32_t x = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
  const uint32_t lane4 = lane * 4;
  unsigned long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    #pragma unroll
    for (int u = 0; u < 4; ++u) {
      x = x * 1664525u + 1013904223u;
      uint32_t w = x ^ (x >> 15);
      #pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == ATOM_LANE32 || MODE == ATOM_CTA_LANE32) {
          // byte address = code*128 + lane*4 : shift+mask+or
          uint32_t off = ((k == 0 ? (w << 7) : (w >> (8 * k - 7))) & 0x7F80u) | lane4;
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == ATOM_LANE32_PRED) {
          uint32_t off = ((k == 0 ? (w << 7) : (w >> (8 * k - 7))) & 0x7F80u) | lane4;
          if ((x >> (28 + k)) & 1) atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == ATOM_SINGLE) {
          uint32_t off = ((k == 0 ? (w << 2) : (w >> (8 * k - 2))) & 0x3FCu);
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == ATOM_COPY4) {
          uint32_t off = ((k == 0 ? (w << 4) : (w >> (8 * k - 4))) & 0xFF0u) | ((lane & 3) << 2);
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == ATOM_COPY8) {
          uint32_t off = ((k == 0 ? (w << 5) : (w >> (8 * k - 5))) & 0x1FE0u) | ((lane & 7) << 2);
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == ATOM_COPY16) {
          uint32_t off = ((k == 0 ? (w << 6) : (w >> (8 * k - 6))) & 0x3FC0u) | ((lane & 15) << 2);
          atomicAdd((uint32_t*)((char*)h + off), 1u);
        } else if (MODE == RMW_LANE32) {
          uint32_t off = ((k == 0 ? (w << 7) : (w >> (8 * k - 7))) & 0x7F80u) | lane4;
          uint32_t* p = (uint32_t*)((char*)h + off);
          *p = *p + 1;
        } else if (MODE == ATOM_5MER_CTA) {
          // 10-bit code, one update per TWO windows: do 2 atomics per word instead of 4
          if (k < 2) {
            uint32_t off = ((k == 0 ? (w << 7) : (w >> 9)) & 0x1FF80u) | lane4;
            atomicAdd((uint32_t*)((char*)h + off), 1u);
          }
        }
      }
    }
  }
  unsigned long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < total_words; i += nthr) s += sm[i];
  if (s == 0xdeadbeef) sink[0] = s;
}

template <int MODE>
void run(const char* name, int threads, int warp_words, int iters, int ctas_per_sm_req) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  size_t smem = (MODE == ATOM_CTA_LANE32) ? 256 * 32 * 4 : (MODE == ATOM_5MER_CTA) ? 1024 * 32 * 4 : (size_t)warp_words * 4 * (threads / 32);
  CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench<MODE>, threads, smem));
  if (occ < 1) { printf("%-28s cannot launch (smem %zu)\n", name, smem); return; }
  int cps = occ < ctas_per_sm_req ? occ : ctas_per_sm_req;
  int grid = nsm * cps;
  unsigned long long* cyc; uint32_t* sink;
  CK(cudaMalloc(&cyc, grid * 8)); CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<grid, threads, smem>>>(iters / 8, cyc, sink, warp_words);   // warm
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  bench<MODE><<<grid, threads, smem>>>(iters, cyc, sink, warp_words);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = (unsigned long long*)malloc(grid * 8);
  CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; double avg = 0;
  for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; avg += h[i]; }
  avg /= grid;
  double windows_per_thread = (double)iters * 16;       // window-equivalents (5-mer mode: 16 windows via 8 atomics)
  double per_sm = windows_per_thread * threads * cps;
  printf("%-28s thr=%4d cta/SM=%d smem=%6zuB  windows/clk/SM: %.2f (avg-cta %.2f)  time %.3f ms -> %.2f Gwin/s chip\n",
         name, threads, cps, smem, per_sm / mx, per_sm / avg, ms, per_sm * nsm / (ms * 1e6));
  cudaFree(cyc); cudaFree(sink); free(h);
}

int main() {
  int iters = 4096;
  run<ATOM_LANE32>("atoms [256][32] per warp", 192, 8192, iters, 1);       // 6 warps x 32KB
  run<ATOM_LANE32>("atoms [256][32] per warp", 128, 8192, iters, 1);       // 4 warps
  run<ATOM_LANE32_PRED>("atoms [256][32] pw 50% pred", 192, 8192, iters, 1);
  run<ATOM_CTA_LANE32>("atoms [256][32] per CTA", 256, 0, iters, 4);
  run<ATOM_CTA_LANE32>("atoms [256][32] per CTA", 512, 0, iters, 2);
  run<ATOM_CTA_LANE32>("atoms [256][32] per CTA", 1024, 0, iters, 1);
  run<ATOM_CTA_LANE32>("atoms [256][32] per CTA", 256, 0, iters, 7);
  run<ATOM_SINGLE>("atoms [256] per warp", 256, 256, iters, 8);
  run<ATOM_COPY4>("atoms [256][4] per warp", 256, 1024, iters, 8);
  run<ATOM_COPY8>("atoms [256][8] per warp", 256, 2048, iters, 3);
  run<ATOM_COPY16>("atoms [256][16] per warp", 256, 4096, iters, 1);
  run<ATOM_COPY16>("atoms [256][16] per warp", 384, 4096, iters, 1);
  run<RMW_LANE32>("lds/sts rmw [256][32] pw", 192, 8192, iters, 1);
  run<ATOM_5MER_CTA>("atoms 5mer [1024][32] CTA", 512, 0, iters, 1);
  run<ATOM_5MER_CTA>("atoms 5mer [1024][32] CTA", 1024, 0, iters, 1);
  return 0;
}
