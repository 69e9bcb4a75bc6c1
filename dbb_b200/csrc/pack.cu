// ASCII -> 2-bit code plane + validity plane (layout in include/dbb_cuda.h), and the padded
// signature copy used by stage 2.
//
// Reference contract (genomicSignatures.py:134-144): a base is usable iff it is
// one of the four UPPERCASE bytes 'A','C','G','T'; every other byte (N, IUPAC, lowercase, U ...) makes
// each window that contains it a KeyError in the reference, i.e. uncounted.
#include "common.cuh"

namespace {

// Exact per-byte "is zero" flags of a word: 0x80 in every byte of v that is 0x00, 0 elsewhere (no carries cross a
// byte: the low seven bits are summed separately from the top bit).
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
  const uint32_t t = (v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | v) & 0x80808080u;
}

// Classification of four ASCII bytes (little-endian word, first base in the low byte).  Every byte is hashed to two
// bits, c2 = ((b>>1)^(b>>2))&3, which separates A, C, G, T (0, 1, 2, 3); the byte is one of the four letters iff
// it equals the letter its own hash selects (one byte permute from the table "ACGT").  The results are shifted into
// running accumulators, first base towards the most significant end: the multiplications gather one field per byte
// into the top bits of the product (their partial products occupy disjoint bit fields, so no carries occur) and a
// funnel shift appends those bits.
template <bool WANT_N>
__device__ __forceinline__ void classify4(uint32_t w, uint32_t& code_acc, uint32_t& ok_acc, uint32_t& n_acc) {
  const uint32_t t = w >> 1;
  const uint32_t c2 = (t ^ (t >> 1)) & 0x03030303u;
  const uint32_t y = (c2 | (c2 >> 4)) & 0x00FF00FFu;                 // bytes -> nibbles: the permute selector
  const uint32_t sel = (y | (y >> 8)) & 0xFFFFu;
  const uint32_t expect = __byte_perm(0x54474341u, 0u, sel);         // 'A','C','G','T' indexed by c2
  const uint32_t ok = zero_bytes(w ^ expect) >> 7;                   // 0x01 per byte that is A, C, G or T
  code_acc = __funnelshift_l((c2 & (ok * 3u)) * 0x40100401u, code_acc, 8);      // 8 code bits (0 for invalid bases)
  ok_acc = __funnelshift_l(ok * 0x80402010u, ok_acc, 4);                          // 4 validity bits
  if (WANT_N) n_acc = __funnelshift_l((zero_bytes((w | 0x20202020u) ^ 0x6e6e6e6eu) >> 7) * 0x80402010u, n_acc, 4);
}

// One thread per 64-base block of the packed layout (16 bytes of codes, one validity word, one N word); a warp
// covers 2 KB of consecutive bases.  Sequence s is ascii[seq_begin[s], seq_end[s]) (for back-to-back sequences
// seq_end = seq_begin + 1 of one offsets array; for contigs cut out of scaffolds two separate arrays).  Each thread
// fetches its 64 bytes as the four or five ALIGNED 16-byte chunks that hold them (a chunk that holds a byte of the
// sequence never leaves the allocation's last 16-byte granule) and realigns them in registers with a two-level
// select + funnel shift.  nmask (optional): a third bit plane with the geometry of the validity plane, bit = base is
// 'N' or 'n' (what splitScaffolds.py:34 tests), input of the N-run splitter in split.cu.
template <bool WANT_N>
__global__ void __launch_bounds__(128) pack_kernel(const uint8_t* __restrict__ ascii,
                                                   const int64_t* __restrict__ seq_begin,
                                                   const int64_t* __restrict__ seq_end,
                                                   const int64_t* __restrict__ blk_off, int64_t n_seq,
                                                   int64_t total_blocks, uint4* __restrict__ codes,
                                                   uint64_t* __restrict__ valid, uint64_t* __restrict__ nmask) {
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // scaffold of the CTA's first block: last s with blk_off[s] <= blk (empty scaffolds share an offset with their
  // successor); one binary search per CTA, the threads then walk forward from it
  __shared__ int64_t s_first;
  if (threadIdx.x == 0) {
    const int64_t blk_first = (int64_t)blockIdx.x * blockDim.x;
    int64_t lo = 0, hi = n_seq;          // invariant: blk_off[lo] <= blk_first < blk_off[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(blk_off + mid) <= blk_first) lo = mid; else hi = mid;
    }
    s_first = lo;
  }
  __syncthreads();
  int64_t s = s_first;
  if (blk >= total_blocks) return;
  while (__ldg(blk_off + s + 1) <= blk) ++s;
  const int64_t base = __ldg(seq_begin + s);
  const int64_t len = __ldg(seq_end + s) - base;
  const int64_t pos0 = (blk - __ldg(blk_off + s)) * DBB_BLOCK_BASES;
  const int64_t avail = len - pos0;
  const int v = avail >= DBB_BLOCK_BASES ? DBB_BLOCK_BASES : (avail > 0 ? (int)avail : 0);   // bases that exist
  uint32_t r[20];
  #pragma unroll
  for (int i = 0; i < 20; ++i) r[i] = 0u;
  int sh = 0;
  if (v > 0) {
    const uint8_t* src = ascii + base + pos0;
    sh = (int)((uintptr_t)src & 15);
    const uint4* a0 = reinterpret_cast<const uint4*>(src - sh);
    #pragma unroll
    for (int c = 0; c < 5; ++c) {
      if (c * 16 < sh + v) {
        const uint4 x = __ldg(a0 + c);
        r[4 * c] = x.x; r[4 * c + 1] = x.y; r[4 * c + 2] = x.z; r[4 * c + 3] = x.w;
      }
    }
  }
  // bytes [sh, sh + 64) of the 80 fetched: word offset by a two-level select, byte offset by a funnel shift
  const bool s1 = (sh & 4) != 0, s2 = (sh & 8) != 0;
  const int bs = (sh & 3) * 8;
  uint32_t t1[19], t2[17];
  #pragma unroll
  for (int i = 0; i < 19; ++i) t1[i] = s1 ? r[i + 1] : r[i];
  #pragma unroll
  for (int i = 0; i < 17; ++i) t2[i] = s2 ? t1[i + 2] : t1[i];
  uint32_t code[4] = {0u, 0u, 0u, 0u}, okw[2] = {0u, 0u}, nw[2] = {0u, 0u};
  #pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t w = __funnelshift_r(t2[i], t2[i + 1], bs);
    classify4<WANT_N>(w, code[i >> 2], okw[i >> 3], nw[i >> 3]);
  }
  if (v < DBB_BLOCK_BASES) {                                         // bytes past the end of the sequence are not bases
    const uint64_t keep = v == 0 ? 0ull : ~(~0ull >> v);
    okw[0] &= (uint32_t)(keep >> 32); okw[1] &= (uint32_t)keep;
    nw[0] &= (uint32_t)(keep >> 32); nw[1] &= (uint32_t)keep;
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int vk = v - 16 * k;                                     // bases of code word k that exist
      code[k] &= vk >= 16 ? ~0u : (vk <= 0 ? 0u : ~(0xFFFFFFFFu >> (2 * vk)));
    }
  }
  if (codes) codes[blk] = make_uint4(code[0], code[1], code[2], code[3]);
  if (valid) valid[blk] = ((uint64_t)okw[0] << 32) | okw[1];         // bit (63 - p) = base p
  if (WANT_N) nmask[blk] = ((uint64_t)nw[0] << 32) | nw[1];
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

static int launch_pack(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end, const int64_t* d_blk_off,
                       int64_t n_seq, int64_t total_blocks, void* d_codes, void* d_valid, void* d_nmask, cudaStream_t st) {
  const int64_t grid = (total_blocks + 127) / 128;
  DBB_REQUIRE(grid < (1ll << 31), "too many blocks for one launch");
  if (d_nmask)
    pack_kernel<true><<<(unsigned)grid, 128, 0, st>>>(d_ascii, d_begin, d_end, d_blk_off, n_seq, total_blocks, (uint4*)d_codes,
                                                      (uint64_t*)d_valid, (uint64_t*)d_nmask);
  else
    pack_kernel<false><<<(unsigned)grid, 128, 0, st>>>(d_ascii, d_begin, d_end, d_blk_off, n_seq, total_blocks, (uint4*)d_codes,
                                                       (uint64_t*)d_valid, nullptr);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

DBB_EXPORT int dbb_pack_ascii_dev(const uint8_t* d_ascii, const int64_t* d_seq_off,
                                  const int64_t* d_blk_off, int64_t n_seq, int64_t total_blocks,
                                  void* d_codes, void* d_valid, void* stream) {
  DBB_REQUIRE(n_seq >= 0 && total_blocks >= 0, "negative size");
  if (total_blocks == 0) return DBB_OK;
  DBB_REQUIRE(d_ascii && d_seq_off && d_blk_off && d_codes && d_valid, "NULL device pointer");
  return launch_pack(d_ascii, d_seq_off, d_seq_off + 1, d_blk_off, n_seq, total_blocks, d_codes, d_valid, nullptr,
                     (cudaStream_t)stream);
}

DBB_EXPORT int dbb_pack_ranges_dev(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end,
                                   const int64_t* d_blk_off, int64_t n_seq, int64_t total_blocks,
                                   void* d_codes, void* d_valid, void* d_nmask, void* stream) {
  DBB_REQUIRE(n_seq >= 0 && total_blocks >= 0, "negative size");
  if (total_blocks == 0) return DBB_OK;
  DBB_REQUIRE(d_ascii && d_begin && d_end && d_blk_off, "NULL device pointer");
  DBB_REQUIRE(d_codes || d_valid || d_nmask, "no output plane requested");
  return launch_pack(d_ascii, d_begin, d_end, d_blk_off, n_seq, total_blocks, d_codes, d_valid, d_nmask, (cudaStream_t)stream);
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
