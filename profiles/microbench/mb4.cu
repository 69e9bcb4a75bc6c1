// Microbenchmark 4 (round 2, design input for K1 v2): the ceiling of the counting stage on B200.
//
// What it measures, per SM and clock, with the histogram layouts the count kernel can use:
//   (a) one shared-memory atomic per S windows into a 4^(3+S)-bin histogram, counter formats
//         MODE 0  u32 counters, +1            (ATOMS.POPC.INC)
//         MODE 1  16-bit counters, two per word, half selected by the LOW bit of the bin   (ATOMS.ADD)
//         MODE 2  16-bit counters, half selected by the TOP bit of the bin
//         MODE 3  16-bit counters, four per 64-bit word (atomicAdd on unsigned long long)
//       on i.i.d. sequence with a given GC content (0.5 = uniform bins; 0.3 = skewed like a low-GC genome)
//   (b) the same with the fold to tetramer counts after every L bases (the real cost of short scaffolds)
// Sequence words come from a host-generated buffer (L2 resident) so that the bins have a realistic law.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb4 mb4.cu ; run: ./mb4
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int S, int MODE>
__device__ __forceinline__ void update(uint32_t* h, uint32_t bin) {
  constexpr int KB = 2 * (3 + S);
  if (MODE == 0) {
    atomicAdd(h + bin, 1u);
  } else if (MODE == 1) {
    atomicAdd(h + (bin >> 1), (bin & 1u) ? 0x10000u : 1u);
  } else if (MODE == 2) {
    atomicAdd(h + (bin & ((1u << (KB - 1)) - 1u)), (bin >> (KB - 1)) ? 0x10000u : 1u);
  } else {
    atomicAdd(reinterpret_cast<unsigned long long*>(h) + (bin >> 2), 1ull << (16 * (bin & 3u)));
  }
}

template <int S, int MODE>
__global__ void __launch_bounds__(1024) hist(const uint4* __restrict__ seq, uint32_t nblk_mask, int iters,
                                              unsigned long long* cyc, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  constexpr int KB = 2 * (3 + S);
  constexpr int BINS = 1 << KB;
  constexpr int WORDS = MODE == 0 ? BINS : BINS / 2;
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t b = (blockIdx.x * 7919u + threadIdx.x) & nblk_mask;
  uint4 nx = seq[b];
  unsigned long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint4 c = nx;
    b = (b + blockDim.x) & nblk_mask;
    nx = seq[b];
    const uint32_t hw = __shfl_down_sync(0xffffffffu, c.x, 1);
    const uint32_t w[5] = {c.x, c.y, c.z, c.w, hw};
    #pragma unroll
    for (int p = 0; p < 64; p += S) {
      const int k = p >> 4, o = p & 15;
      const int sh = 64 - 2 * o - KB;
      const uint32_t x = (sh >= 32) ? (w[k] >> (sh - 32)) : __funnelshift_r(w[k + 1], w[k], sh);
      update<S, MODE>(sm, x & (BINS - 1));
    }
  }
  unsigned long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s += sm[i];
  if (s == 0xdeadbeef) sink[0] = s;
}

template <int S, int MODE>
void run(const char* name, const uint4* d_seq, uint32_t mask, int threads, int cps_req, int iters, double gc) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  size_t bins = (size_t)1 << (2 * (3 + S));
  size_t smem = (MODE == 0 ? bins : bins / 2) * 4;
  if (smem > 227 * 1024) { printf("%-28s S=%d too much smem\n", name, S); return; }
  auto kern = hist<S, MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) { printf("%-28s cannot launch\n", name); return; }
  int cps = occ < cps_req ? occ : cps_req;
  int grid = nsm * cps;
  unsigned long long* cyc; uint32_t* sink; CK(cudaMalloc(&cyc, grid * 8)); CK(cudaMalloc(&sink, 4));
  kern<<<grid, threads, smem>>>(d_seq, mask, iters / 8 + 1, cyc, sink); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, threads, smem>>>(d_seq, mask, iters, cyc, sink);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double per_sm = (double)iters * 64 * threads * cps;
  printf("%-28s S=%d GC=%.2f thr=%4d cta/SM=%2d smem/CTA=%6zu  windows/clk/SM %6.2f  (%.1f Gwin/s chip, %.3f ms)\n", name, S, gc, threads, cps,
         smem, per_sm / mx, per_sm * nsm / (ms * 1e6), ms);
  cudaFree(cyc); cudaFree(sink);
}

static std::vector<uint32_t> make_seq(size_t nblocks, double gc, uint64_t seed) {
  std::vector<uint32_t> w(nblocks * 4);
  uint64_t x = seed;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (double)(x >> 11) / 9007199254740992.0; };
  for (auto& word : w) {
    uint32_t v = 0;
    for (int i = 0; i < 16; ++i) {
      const double r = rnd();
      // A=0 C=1 G=2 T=3
      uint32_t base = r < (1 - gc) / 2 ? 0u : r < 0.5 ? 1u : r < 0.5 + gc / 2 ? 2u : 3u;
      v = (v << 2) | base;
    }
    word = v;
  }
  return w;
}

int main() {
  const size_t nblocks = 1 << 16;                    // 4 Mbases, 1 MB: L2 resident
  for (double gc : {0.5, 0.3}) {
    std::vector<uint32_t> h = make_seq(nblocks, gc, 88172645463325252ull);
    uint4* d; CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const uint32_t mask = (uint32_t)nblocks - 1;
    const int it = 400;
#define ROW(S_, M_, name, thr, cps) run<S_, M_>(name, d, mask, thr, cps, it, gc)
    ROW(1, 0, "u32 +1", 256, 8);
    ROW(2, 0, "u32 +1", 256, 8);
    ROW(3, 0, "u32 +1", 256, 8);
    ROW(4, 0, "u32 +1", 256, 3);
    ROW(2, 1, "u16x2 low-bit half", 256, 8);
    ROW(3, 1, "u16x2 low-bit half", 256, 8);
    ROW(4, 1, "u16x2 low-bit half", 256, 6);
    ROW(5, 1, "u16x2 low-bit half", 512, 1);
    ROW(3, 2, "u16x2 top-bit half", 256, 8);
    ROW(4, 2, "u16x2 top-bit half", 256, 6);
    ROW(5, 2, "u16x2 top-bit half", 512, 1);
    ROW(3, 3, "u16x4 in u64", 256, 8);
    ROW(4, 3, "u16x4 in u64", 256, 6);
    ROW(4, 1, "u16x2 low-bit half", 128, 6);
    ROW(4, 1, "u16x2 low-bit half", 512, 3);
    ROW(4, 1, "u16x2 low-bit half", 1024, 1);
    ROW(3, 1, "u16x2 low-bit half", 128, 16);
    ROW(3, 1, "u16x2 low-bit half", 64, 24);
    // atomics in flight: how many warps does it take to fill the pipe?
    ROW(3, 0, "u32 +1   (few warps)", 64, 1);
    ROW(3, 0, "u32 +1   (few warps)", 128, 1);
    ROW(3, 0, "u32 +1   (few warps)", 256, 1);
    ROW(3, 0, "u32 +1   (few warps)", 512, 1);
    ROW(3, 0, "u32 +1   (few warps)", 1024, 1);
    ROW(3, 2, "u16x2 top (few warps)", 64, 1);
    ROW(3, 2, "u16x2 top (few warps)", 128, 1);
    ROW(3, 2, "u16x2 top (few warps)", 256, 1);
    ROW(3, 2, "u16x2 top (few warps)", 512, 1);
    ROW(3, 2, "u16x2 top (few warps)", 1024, 1);
    ROW(4, 2, "u16x2 top (few warps)", 64, 1);
    ROW(4, 2, "u16x2 top (few warps)", 128, 1);
    ROW(4, 2, "u16x2 top (few warps)", 256, 1);
    ROW(4, 2, "u16x2 top (few warps)", 512, 1);
    ROW(4, 2, "u16x2 top (few warps)", 32, 6);
    ROW(3, 2, "u16x2 top (few warps)", 32, 24);
#undef ROW
    cudaFree(d);
  }
  return 0;
}
