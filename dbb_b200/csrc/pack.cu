// ASCII -> 2-bit code plane + validity plane (layout in include/dbb_cuda.h), and the padded
// signature copy used by stage 2.
//
// Reference contract (genomicSignatures.py:134-144): a base is usable iff it is
// one of the four UPPERCASE bytes 'A','C','G','T'; every other byte (N, IUPAC, lowercase, U ...) makes
// each window that contains it a KeyError in the reference, i.e. uncounted.
#include "common.cuh"

namespace {

// one thread per 16-base word of the code plane.  Sequence s is ascii[seq_begin[s], seq_end[s]) (for back-to-back
// sequences seq_end = seq_begin + 1 of one offsets array; for contigs cut out of scaffolds two separate arrays).
// nmask16 (optional): a third bit plane with the same geometry as the validity plane, bit = base is 'N' or 'n'
// (what splitScaffolds.py:34 tests), input of the N-run splitter in split.cu.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ ascii,
                                                   const int64_t* __restrict__ seq_begin,
                                                   const int64_t* __restrict__ seq_end,
                                                   const int64_t* __restrict__ blk_off, int64_t n_seq,
                                                   int64_t total_words, uint32_t* __restrict__ codes,
                                                   uint16_t* __restrict__ valid16, uint16_t* __restrict__ nmask16) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_words) return;
  int64_t blk = g >> 2;
  int q = (int)(g & 3);
  // last scaffold s with blk_off[s] <= blk  (empty scaffolds share an offset with their successor)
  int64_t lo = 0, hi = n_seq;            // invariant: blk_off[lo] <= blk < blk_off[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(blk_off + mid) <= blk) lo = mid; else hi = mid;
  }
  const int64_t s = lo;
  const int64_t base = __ldg(seq_begin + s);
  const int64_t len = __ldg(seq_end + s) - base;
  const int64_t pos0 = (blk - __ldg(blk_off + s)) * DBB_BLOCK_BASES + q * 16;
  uint32_t code = 0, vmask = 0, nm = 0;
  #pragma unroll
  for (int i = 0; i < 16; ++i) {
    int64_t p = pos0 + i;
    uint32_t ch = (p < len) ? (uint32_t)__ldg(ascii + base + p) : 0u;
    uint32_t ok = (ch == 'A') | (ch == 'C') | (ch == 'G') | (ch == 'T');
    uint32_t c2 = ((ch >> 1) ^ (ch >> 2)) & 3u;          // A0 C1 G2 T3 for the four valid letters
    code = (code << 2) | (ok ? c2 : 0u);
    vmask = (vmask << 1) | ok;
    nm = (nm << 1) | (uint32_t)((ch | 0x20u) == 'n');
  }
  if (codes) codes[g] = code;
  // uint64 valid word of the block: bit (63-p); as little-endian uint16[4], element 3-q holds bases 16q..16q+15
  if (valid16) valid16[(blk << 2) + (3 - q)] = (uint16_t)vmask;
  if (nmask16) nmask16[(blk << 2) + (3 - q)] = (uint16_t)nm;
}

__global__ void __launch_bounds__(256) pad_sig_kernel(const float* __restrict__ dense, int64_t n,
                                                      float* __restrict__ padded) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * DBB_SIG_STRIDE) return;
  int64_t r = g / DBB_SIG_STRIDE;
  int c = (int)(g - r * DBB_SIG_STRIDE);
  padded[g] = c < DBB_NUM_KMERS ? dense[r * DBB_NUM_KMERS + c] : 0.f;
}

}  // namespace

DBB_EXPORT int dbb_pack_ascii_dev(const uint8_t* d_ascii, const int64_t* d_seq_off,
                                  const int64_t* d_blk_off, int64_t n_seq, int64_t total_blocks,
                                  void* d_codes, void* d_valid, void* stream) {
  DBB_REQUIRE(n_seq >= 0 && total_blocks >= 0, "negative size");
  if (total_blocks == 0) return DBB_OK;
  DBB_REQUIRE(d_ascii && d_seq_off && d_blk_off && d_codes && d_valid, "NULL device pointer");
  int64_t words = total_blocks * 4;
  int64_t grid = (words + 255) / 256;
  DBB_REQUIRE(grid < (1ll << 31), "too many blocks for one launch");
  pack_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      d_ascii, d_seq_off, d_seq_off + 1, d_blk_off, n_seq, words, (uint32_t*)d_codes, (uint16_t*)d_valid, nullptr);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

DBB_EXPORT int dbb_pack_ranges_dev(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end,
                                   const int64_t* d_blk_off, int64_t n_seq, int64_t total_blocks,
                                   void* d_codes, void* d_valid, void* d_nmask, void* stream) {
  DBB_REQUIRE(n_seq >= 0 && total_blocks >= 0, "negative size");
  if (total_blocks == 0) return DBB_OK;
  DBB_REQUIRE(d_ascii && d_begin && d_end && d_blk_off, "NULL device pointer");
  DBB_REQUIRE(d_codes || d_valid || d_nmask, "no output plane requested");
  int64_t words = total_blocks * 4;
  int64_t grid = (words + 255) / 256;
  DBB_REQUIRE(grid < (1ll << 31), "too many blocks for one launch");
  pack_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      d_ascii, d_begin, d_end, d_blk_off, n_seq, words, (uint32_t*)d_codes, (uint16_t*)d_valid, (uint16_t*)d_nmask);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

DBB_EXPORT int dbb_pad_signatures_dev(const float* d_dense, int64_t n, float* d_padded, void* stream) {
  DBB_REQUIRE(n >= 0, "negative n");
  if (n == 0) return DBB_OK;
  DBB_REQUIRE(d_dense && d_padded, "NULL device pointer");
  int64_t grid = (n * DBB_SIG_STRIDE + 255) / 256;
  pad_sig_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_dense, n, d_padded);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}
