// Host-side packer: ASCII -> 2-bit code plane + validity plane, bit-identical to pack_kernel (csrc/pack.cu), with the
// CPU's vector units (AVX2 + BMI2, chosen at run time; a byte-table loop otherwise).
//
// Why it exists: a caller that holds its sequences in PAGEABLE memory (numpy arrays, Python bytes) cannot feed the copy
// engine directly; instead of staging the bytes through a pinned ring (two more passes over host memory),
// dbb_tetra_signatures_host lets a few host threads pack them straight into pinned planes, 0.375 bytes per base on the
// link (csrc/host_api.cu::hybrid_upload: 30.8 -> 43.6 Gbp/s on C2).  Also exported as dbb_pack_ascii_host.  Reference contract as in pack.cu (genomicSignatures.py:134-144): a base is usable iff it is one
// of the UPPERCASE bytes A C G T; codes of unusable bases and of the padding are 0.
//
// Plain C++ (no CUDA): compiled by g++ so that the target attributes and intrinsics below stay out of nvcc's front end.
#include "../../include/dbb_cuda.h"
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace {

struct Lut {
  uint8_t v[256];     // 0..3 for A C G T, 0xff otherwise
  Lut() {
    memset(v, 0xff, sizeof(v));
    v[(unsigned char)'A'] = 0; v[(unsigned char)'C'] = 1; v[(unsigned char)'G'] = 2; v[(unsigned char)'T'] = 3;
  }
};
const Lut g_lut;

// one block of `nb` (1..64) bases, any CPU
void pack_block_scalar(const uint8_t* p, int nb, uint32_t* code4, uint64_t* valid) {
  uint32_t c[4] = {0u, 0u, 0u, 0u};
  uint64_t ok = 0;
  for (int i = 0; i < nb; ++i) {
    const uint8_t x = g_lut.v[p[i]];
    if (x != 0xff) {
      c[i >> 4] |= (uint32_t)x << (30 - 2 * (i & 15));
      ok |= 1ull << (63 - i);
    }
  }
  code4[0] = c[0]; code4[1] = c[1]; code4[2] = c[2]; code4[3] = c[3];
  *valid = ok;
}

// 32 bases -> 64 code bits (bases 0..15 in the upper word, first base most significant) and 32 validity bits
// (bit 31 = base 0).  The bytes are reversed first, so that movemask / pext, which put byte 0 into the LEAST significant
// position, deliver the first base in the MOST significant one.
__attribute__((target("avx2,bmi2"))) inline void pack32_avx2(const uint8_t* p, uint64_t* codes64, uint32_t* valid32) {
  const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                       15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  const __m256i letters = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                           'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  // reverse inside the 128-bit lanes, then swap the lanes
  __m256i r = _mm256_shuffle_epi8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(p)), rev);
  r = _mm256_permute4x64_epi64(r, 0x4E);                                     // byte j = base 31 - j
  // hash to two bits: ((b >> 1) ^ (b >> 2)) & 3 separates A C G T as 0 1 2 3; the 16-bit shifts only disturb bits the
  // mask drops
  const __m256i c2 = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(r, 1), _mm256_srli_epi16(r, 2)), _mm256_set1_epi8(3));
  const __m256i ok = _mm256_cmpeq_epi8(r, _mm256_shuffle_epi8(letters, c2)); // the byte IS the letter its hash selects
  const __m256i code = _mm256_and_si256(c2, ok);
  *valid32 = (uint32_t)_mm256_movemask_epi8(ok);
  const uint64_t m = 0x0303030303030303ull;
  const uint64_t q0 = _pext_u64((uint64_t)_mm256_extract_epi64(code, 0), m);  // bases 31..24 (base 24 in bits 15:14)
  const uint64_t q1 = _pext_u64((uint64_t)_mm256_extract_epi64(code, 1), m);
  const uint64_t q2 = _pext_u64((uint64_t)_mm256_extract_epi64(code, 2), m);
  const uint64_t q3 = _pext_u64((uint64_t)_mm256_extract_epi64(code, 3), m);  // bases 7..0 (base 0 in bits 15:14)
  *codes64 = (q3 << 48) | (q2 << 32) | (q1 << 16) | q0;
}

__attribute__((target("avx2,bmi2"))) void pack_blocks_avx2(const uint8_t* p, int64_t n_full, uint32_t* codes, uint64_t* valid) {
  for (int64_t b = 0; b < n_full; ++b, p += 64) {
    uint64_t c01, c23;
    uint32_t v0, v1;
    pack32_avx2(p, &c01, &v0);
    pack32_avx2(p + 32, &c23, &v1);
    codes[4 * b + 0] = (uint32_t)(c01 >> 32); codes[4 * b + 1] = (uint32_t)c01;
    codes[4 * b + 2] = (uint32_t)(c23 >> 32); codes[4 * b + 3] = (uint32_t)c23;
    valid[b] = ((uint64_t)v0 << 32) | v1;
  }
}

bool have_avx2() {
  // DBB_HOST_PACK_SCALAR=1 forces the table loop (tests compare the two)
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") &&
                         !(getenv("DBB_HOST_PACK_SCALAR") && *getenv("DBB_HOST_PACK_SCALAR") == '1');
  return ok;
}

}  // namespace

namespace dbb {

// one scaffold of `len` bases -> its ceil(len / 64) blocks
void host_pack_sequence(const uint8_t* seq, int64_t len, uint32_t* codes, uint64_t* valid) {
  if (len <= 0) return;
  const int64_t n_full = len / 64;
  const int tail = (int)(len - n_full * 64);
  if (have_avx2()) {
    pack_blocks_avx2(seq, n_full, codes, valid);
    if (tail) {
      alignas(32) uint8_t tmp[64];
      memset(tmp, 0, sizeof(tmp));                        // 0 is not a letter: the padding comes out invalid, code 0
      memcpy(tmp, seq + n_full * 64, (size_t)tail);
      pack_blocks_avx2(tmp, 1, codes + 4 * n_full, valid + n_full);
    }
  } else {
    for (int64_t b = 0; b < n_full; ++b) pack_block_scalar(seq + 64 * b, 64, codes + 4 * b, valid + b);
    if (tail) pack_block_scalar(seq + 64 * n_full, tail, codes + 4 * n_full, valid + n_full);
  }
}

}  // namespace dbb

extern "C" __attribute__((visibility("default")))
int dbb_pack_ascii_host(const uint8_t* ascii, const int64_t* seq_off, const int64_t* blk_off, int64_t n_seq,
                        void* codes, void* valid) {
  if (n_seq < 0 || (n_seq > 0 && (!seq_off || !blk_off || !codes || !valid))) return DBB_ERR_INVALID;
  for (int64_t s = 0; s < n_seq; ++s) {
    const int64_t len = seq_off[s + 1] - seq_off[s];
    if (len < 0 || (len > 0 && !ascii)) return DBB_ERR_INVALID;
    dbb::host_pack_sequence(ascii + seq_off[s], len, (uint32_t*)codes + 4 * blk_off[s], (uint64_t*)valid + blk_off[s]);
  }
  return DBB_OK;
}
