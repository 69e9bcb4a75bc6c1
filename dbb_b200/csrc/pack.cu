// ASCII -> 2-bit code plane + validity plane (layout in include/dbb_cuda.h), and the padded
// signature copy used by stage 2.
//
// Reference contract (genomicSignatures.py:134-144): a base is usable iff it is
// one of the four UPPERCASE bytes 'A','C','G','T'; every other byte (N, IUPAC, lowercase, U ...) makes
// each window that contains it a KeyError in the reference, i.e. uncounted.
#include "common.cuh"

namespace {

// four ASCII bytes (little-endian word, first base in the low byte) -> 8 code bits, 4 validity bits, 4 N bits,
// first base in the most significant position.  The multiplications gather one field per byte into the top byte;
// the partial products occupy disjoint bit fields, so no carries occur.
__device__ __forceinline__ void classify4(uint32_t w, uint32_t& code8, uint32_t& ok4, uint32_t& n4) {
  const uint32_t ok = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) |
                      __vcmpeq4(w, 0x54545454u);                       // 0xFF per byte that is A, C, G or T
  const uint32_t c2 = ((w >> 1) ^ (w >> 2)) & 0x03030303u & ok;      // A0 C1 G2 T3, 0 for everything else
  code8 = (c2 * 0x40100401u) >> 24;
  ok4 = ((ok & 0x01010101u) * 0x08040201u) >> 24;
  n4 = ((__vcmpeq4(w | 0x20202020u, 0x6e6e6e6eu) & 0x01010101u) * 0x08040201u) >> 24;
}

// One thread per 16-base word of the code plane; a warp covers 512 consecutive bases of the packed layout.
// Sequence s is ascii[seq_begin[s], seq_end[s]) (for back-to-back sequences seq_end = seq_begin + 1 of one offsets
// array; for contigs cut out of scaffolds two separate arrays).  Each lane fetches its 16 bytes as the one or two
// ALIGNED 16-byte chunks that hold them (a chunk that holds a byte of the sequence never leaves the allocation's last
// 16-byte granule) and realigns in registers, so a warp's loads are coalesced whatever the sequence's offset.
// nmask16 (optional): a third bit plane with the same geometry as the validity plane, bit = base is 'N' or 'n'
// (what splitScaffolds.py:34 tests), input of the N-run splitter in split.cu.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ ascii,
                                                   const int64_t* __restrict__ seq_begin,
                                                   const int64_t* __restrict__ seq_end,
                                                   const int64_t* __restrict__ blk_off, int64_t n_seq,
                                                   int64_t total_words, uint32_t* __restrict__ codes,
                                                   uint16_t* __restrict__ valid16, uint16_t* __restrict__ nmask16) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // scaffold of the CTA's first word: last s with blk_off[s] <= blk (empty scaffolds share an offset with their
  // successor); one binary search per CTA, the threads then walk forward from it
  __shared__ int64_t s_first;
  if (threadIdx.x == 0) {
    const int64_t blk_first = ((int64_t)blockIdx.x * blockDim.x) >> 2;
    int64_t lo = 0, hi = n_seq;          // invariant: blk_off[lo] <= blk_first < blk_off[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(blk_off + mid) <= blk_first) lo = mid; else hi = mid;
    }
    s_first = lo;
  }
  __syncthreads();
  int64_t s = s_first;
  if (g >= total_words) return;
  const int64_t blk = g >> 2;
  const int q = (int)(g & 3);
  while (__ldg(blk_off + s + 1) <= blk) ++s;
  const int64_t base = __ldg(seq_begin + s);
  const int64_t len = __ldg(seq_end + s) - base;
  const int64_t pos0 = (blk - __ldg(blk_off + s)) * DBB_BLOCK_BASES + q * 16;
  const int64_t avail = len - pos0;
  const int v = avail >= 16 ? 16 : (avail > 0 ? (int)avail : 0);     // bases of this word that exist
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (v > 0) {
    const uint8_t* src = ascii + base + pos0;
    const int sh = (int)((uintptr_t)src & 15);
    const uint4* a0 = reinterpret_cast<const uint4*>(src - sh);
    const uint4 x = __ldg(a0);
    uint4 y = make_uint4(0u, 0u, 0u, 0u);
    if (sh + v > 16) y = __ldg(a0 + 1);
    const uint32_t r[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    const int wi = sh >> 2, bs = (sh & 3) * 8;
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
      // r[wi + i], r[wi + i + 1] with wi in 0..3: selected by predication, not by indexing
      const uint32_t lo32 = wi == 0 ? r[i] : (wi == 1 ? r[i + 1] : (wi == 2 ? r[i + 2] : r[i + 3]));
      const uint32_t hi32 = wi == 0 ? r[i + 1] : (wi == 1 ? r[i + 2] : (wi == 2 ? r[i + 3] : r[i + 4]));
      w[i] = __funnelshift_r(lo32, hi32, bs);
    }
  }
  uint32_t code = 0, vmask = 0, nm = 0;
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t c8, o4, n4;
    classify4(w[i], c8, o4, n4);
    code = (code << 8) | c8;
    vmask = (vmask << 4) | o4;
    nm = (nm << 4) | n4;
  }
  if (v < 16) {                                                      // bytes past the end of the sequence are not bases
    const uint32_t keep16 = ~(0xFFFFu >> v) & 0xFFFFu;
    vmask &= keep16;
    nm &= keep16;
    code &= v == 0 ? 0u : ~(0xFFFFFFFFu >> (2 * v));
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
