// Device twin of dbb_b200/synthetic.py::generate_block -- emits IDENTICAL bytes (benchmark / test
// support, not part of the reference's path).  One thread per 64-base chunk of a scaffold.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) synth_kernel(uint64_t seed, int64_t n_seq,
                                                    const int64_t* __restrict__ seq_off,
                                                    const int64_t* __restrict__ blk_off,
                                                    const int32_t* __restrict__ genome,
                                                    const uint32_t* __restrict__ cum,
                                                    const int64_t* __restrict__ n_start,
                                                    const int64_t* __restrict__ n_len,
                                                    const int64_t* __restrict__ lc_start,
                                                    const int64_t* __restrict__ lc_len,
                                                    int64_t total_chunks, uint8_t* __restrict__ ascii) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_chunks) return;
  int64_t lo = 0, hi = n_seq;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(blk_off + mid) <= g) lo = mid; else hi = mid;
  }
  const int64_t s = lo;
  const int64_t ck = g - __ldg(blk_off + s);
  const int64_t base = __ldg(seq_off + s);
  const int64_t len = __ldg(seq_off + s + 1) - base;
  const uint64_t seedk = mix64(seed + 0x1234567ull);
  const uint64_t skey = seedk ^ ((uint64_t)s * 0x9E3779B97F4A7C15ull);
  uint32_t ctx = (uint32_t)(mix64(skey ^ ((uint64_t)ck * 0x8CB92BA72F3D8DD7ull)) >> 58);
  const uint32_t* tab = cum + (size_t)__ldg(genome + s) * 64 * 3;
  const int64_t ns = __ldg(n_start + s), nl = __ldg(n_len + s);
  const int64_t ls = __ldg(lc_start + s), ll = __ldg(lc_len + s);
  for (int t = 0; t < 64; ++t) {
    const int64_t p = ck * 64 + t;
    const uint32_t u = (uint32_t)(mix64(skey + (uint64_t)p * 0xD1B54A32D192ED03ull) >> 40);
    const uint32_t* th = tab + ctx * 3;
    const uint32_t b = (u >= __ldg(th)) + (u >= __ldg(th + 1)) + (u >= __ldg(th + 2));
    ctx = ((ctx << 2) | b) & 63u;
    if (p < len) {
      uint8_t ch = (uint8_t)("ACGT"[b]);
      if (ns >= 0 && p >= ns && p < ns + nl) ch = 'N';
      if (ls >= 0 && p >= ls && p < ls + ll) ch |= 0x20;
      ascii[base + p] = ch;
    }
  }
}

}  // namespace

DBB_EXPORT int dbb_synth_ascii_dev(uint64_t seed, int64_t n_seq, const int64_t* d_seq_off,
                                   const int64_t* d_blk_off, int64_t total_blocks, const int32_t* d_genome,
                                   const uint32_t* d_cum, const int64_t* d_n_start, const int64_t* d_n_len,
                                   const int64_t* d_lc_start, const int64_t* d_lc_len, uint8_t* d_ascii,
                                   void* stream) {
  DBB_REQUIRE(n_seq >= 0 && total_blocks >= 0, "negative size");
  if (total_blocks == 0) return DBB_OK;
  DBB_REQUIRE(d_seq_off && d_blk_off && d_genome && d_cum && d_n_start && d_n_len && d_lc_start && d_lc_len && d_ascii,
              "NULL device pointer");
  int64_t grid = (total_blocks + 255) / 256;
  synth_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(seed, n_seq, d_seq_off, d_blk_off, d_genome, d_cum,
                                                              d_n_start, d_n_len, d_lc_start, d_lc_len,
                                                              total_blocks, d_ascii);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}
