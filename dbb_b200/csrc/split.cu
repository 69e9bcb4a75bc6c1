// N2 row: the sequence side of preprocess.py's per-scaffold worker.
//   * N-run splitter: SplitScaffolds.splits, splitScaffolds.py:27-53 -- a scaffold is cut at every run of 'N'/'n'
//     LONGER than minN that is followed by another base; the contig ends where the run starts, the next one starts
//     after it.  Quirks kept: a qualifying leading run yields the empty contig [0, 0]; a trailing run of ANY length
//     is trimmed from the last contig; a last contig that would start on the final base is dropped (:46).
//   * base composition: baseCount, seqUtils.py:258-265 -- case-insensitive counts of A, C, G and T+U for arbitrary
//     ranges (scaffolds and their contigs); preprocess.py:94-100 turns them into GC = (G+C) / (A+C+G+T) on the host.
// The splitter works on the N bit plane written by pack_kernel (one bit per base, 64 per block, MSB first), so it
// reads 1/8 byte per base; runs are rare, so the scan is a warp-uniform state machine that only does work in
// 2 048-base steps that contain an N (or continue a run).
#include "common.cuh"

namespace {

// One warp per scaffold.  Pass 1 (contig_off == nullptr) counts contigs, pass 2 writes [begin, end) pairs
// (positions relative to the scaffold start) at contig_off[s].
__global__ void __launch_bounds__(256) split_kernel(const uint64_t* __restrict__ nmask, const int64_t* __restrict__ seq_len,
                                                    const int64_t* __restrict__ blk_off, int64_t n_seq, int64_t min_n,
                                                    const int64_t* __restrict__ contig_off, int32_t* __restrict__ n_contigs,
                                                    int64_t* __restrict__ out_begin, int64_t* __restrict__ out_end) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= n_seq) return;
  const int64_t len = seq_len[s];
  const int64_t b0 = blk_off[s], nb = blk_off[s + 1] - b0;
  const bool write = contig_off != nullptr;
  int64_t o = write ? contig_off[s] : 0;
  // warp-uniform state of the reference's loop
  int64_t start_index = 0, n_start = 0;
  bool in_run = false;
  int32_t count = 0;
  for (int64_t blk = 0; blk < nb; blk += 32) {
    const int64_t my = blk + lane;
    // MSB-first plane -> bit i of m = base 64*my + i
    const uint64_t m = my < nb ? __brevll(__ldg(nmask + b0 + my)) : 0ull;
    if (!__any_sync(0xffffffffu, m != 0ull) && !in_run) continue;
    #pragma unroll 1
    for (int w = 0; w < 32; ++w) {
      const uint64_t mw = __shfl_sync(0xffffffffu, m, w);
      if (mw == 0ull && !in_run) continue;
      const int64_t p0 = (blk + w) * 64;
      if (p0 >= len) break;
      uint64_t t = mw ^ ((mw << 1) | (in_run ? 1ull : 0ull));     // transitions: run starts and ends
      while (t) {
        const int i = __ffsll((long long)t) - 1;
        t &= t - 1;
        const int64_t pos = p0 + i;
        if (pos >= len) break;                                    // padding beyond the scaffold is not a base
        if ((mw >> i) & 1ull) {                                   // splitScaffolds.py:34-37
          in_run = true;
          n_start = pos;
        } else {                                                  // :38-45, first base after a run
          if (pos - n_start > min_n) {
            if (write && lane == 0) { out_begin[o] = start_index; out_end[o] = n_start; }
            ++o; ++count;
            start_index = pos;
          }
          in_run = false;
        }
      }
    }
  }
  if (start_index != len - 1) {                                   // :46-50
    if (write && lane == 0) { out_begin[o] = start_index; out_end[o] = in_run ? n_start : len; }
    ++count;
  }
  if (!write && lane == 0) n_contigs[s] = count;
}

// One warp per range: case-insensitive A, C, G, T+U counts (seqUtils.py:258-265).  The body streams 16 bytes per
// lane; per 4-byte word the letters are recognised with the hash-and-compare of pack.cu's classify4 (upper-cased
// first; 'U' hashes to T's slot and differs from 'T' in bit 0 only, which is masked there), and the four per-byte
// indicator words are summed byte-wise (SIMD within a register), flushed to 64-bit totals every 255 words.
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
  const uint32_t t = (v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | v) & 0x80808080u;
}
__device__ __forceinline__ uint32_t hsum_bytes(uint32_t x) {         // sum of the four bytes of x
  x = (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu);
  return (x & 0xFFFFu) + (x >> 16);
}

__global__ void __launch_bounds__(256) base_count_kernel(const uint8_t* __restrict__ ascii, const int64_t* __restrict__ begin,
                                                         const int64_t* __restrict__ end, int64_t n_ranges,
                                                         int64_t* __restrict__ acgt) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_ranges) return;
  const int64_t b = begin[r], e = end[r];
  // totals: letters (any of A C G T U), hash bit 0 set (C, T/U), hash bit 1 set (G, T/U), both (T/U)
  unsigned long long n_ok = 0, n_lo = 0, n_hi = 0, n_both = 0;
  uint32_t s_ok = 0, s_lo = 0, s_hi = 0, s_both = 0;                 // byte-wise partial sums
  int pending = 0;
  auto add_word = [&](uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu;                              // upper-case the letters
    const uint32_t t = u >> 1;
    const uint32_t c2 = (t ^ (t >> 1)) & 0x03030303u;
    const uint32_t y = (c2 | (c2 >> 4)) & 0x00FF00FFu;
    const uint32_t expect = __byte_perm(0x54474341u, 0u, (y | (y >> 8)) & 0xFFFFu);
    const uint32_t both = c2 & (c2 >> 1) & 0x01010101u;              // hash 3: T's slot, where 'U' (0x55) is accepted too
    const uint32_t ok = zero_bytes((u ^ expect) & ~both) >> 7;
    s_ok += ok; s_lo += ok & c2; s_hi += ok & (c2 >> 1); s_both += ok & both;
  };
  auto flush = [&]() {
    n_ok += hsum_bytes(s_ok); n_lo += hsum_bytes(s_lo); n_hi += hsum_bytes(s_hi); n_both += hsum_bytes(s_both);
    s_ok = s_lo = s_hi = s_both = 0; pending = 0;
  };
  auto add_byte = [&](uint32_t ch) { add_word(ch); };                // the three zero bytes above it count nothing
  // head up to 16-byte alignment, aligned 16-byte vectors, tail
  int64_t p = b;
  const int64_t head_end = min(e, (b + 15) & ~(int64_t)15);
  if (p + lane < head_end) add_byte(__ldg(ascii + p + lane));
  p = head_end;
  const int64_t vecs = (e - p) >> 4;
  const uint4* v128 = reinterpret_cast<const uint4*>(ascii + p);
  for (int64_t i = lane; i < vecs; i += 32) {
    const uint4 v = __ldg(v128 + i);
    add_word(v.x); add_word(v.y); add_word(v.z); add_word(v.w);
    pending += 4;
    if (pending >= 252) flush();
  }
  p += vecs * 16;
  if (p + lane < e) add_byte(__ldg(ascii + p + lane));
  flush();
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n_ok += __shfl_xor_sync(0xffffffffu, n_ok, d); n_lo += __shfl_xor_sync(0xffffffffu, n_lo, d);
    n_hi += __shfl_xor_sync(0xffffffffu, n_hi, d); n_both += __shfl_xor_sync(0xffffffffu, n_both, d);
  }
  if (lane == 0) {
    // hash 0 = A, 1 = C, 2 = G, 3 = T/U
    acgt[4 * r + 0] = (int64_t)(n_ok - n_lo - n_hi + n_both);
    acgt[4 * r + 1] = (int64_t)(n_lo - n_both);
    acgt[4 * r + 2] = (int64_t)(n_hi - n_both);
    acgt[4 * r + 3] = (int64_t)n_both;
  }
}

}  // namespace

DBB_EXPORT int dbb_split_scaffolds_dev(const void* d_nmask, const int64_t* d_seq_len, const int64_t* d_blk_off,
                                       int64_t n_seq, int64_t min_n, const int64_t* d_contig_off,
                                       int32_t* d_n_contigs, int64_t* d_begin, int64_t* d_end, void* stream) {
  DBB_REQUIRE(n_seq >= 0 && min_n >= 0, "negative size");
  if (n_seq == 0) return DBB_OK;
  DBB_REQUIRE(d_nmask && d_seq_len && d_blk_off, "NULL device pointer");
  DBB_REQUIRE(d_contig_off ? (d_begin && d_end) : (d_n_contigs != nullptr), "count pass needs d_n_contigs, write pass d_begin/d_end");
  const int64_t grid = (n_seq * 32 + 255) / 256;
  DBB_REQUIRE(grid < (1ll << 31), "too many scaffolds for one launch");
  split_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const uint64_t*)d_nmask, d_seq_len, d_blk_off, n_seq, min_n,
                                                                d_contig_off, d_n_contigs, d_begin, d_end);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

DBB_EXPORT int dbb_base_count_dev(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end, int64_t n_ranges,
                                  int64_t* d_acgt, void* stream) {
  DBB_REQUIRE(n_ranges >= 0, "negative size");
  if (n_ranges == 0) return DBB_OK;
  DBB_REQUIRE(d_ascii && d_begin && d_end && d_acgt, "NULL device pointer");
  DBB_REQUIRE(((uintptr_t)d_ascii & 15) == 0, "d_ascii must be 16-byte aligned");
  const int64_t grid = (n_ranges * 32 + 255) / 256;
  DBB_REQUIRE(grid < (1ll << 31), "too many ranges for one launch");
  base_count_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_ascii, d_begin, d_end, n_ranges, d_acgt);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}
