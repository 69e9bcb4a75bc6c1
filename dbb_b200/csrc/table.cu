// Canonical tetramer table and library-wide utilities.
//
// Reference behaviour (genomicSignatures.py:44-72): enumerate all 4-mers in lexicographic order,
// canonical form = the smaller of {k-mer, reverse complement} by string compare, column index =
// order of first appearance.  With A<C<G<T encoded 0..3 (first base most significant) lexicographic
// order is numeric order of the 8-bit code, the reverse complement is "reverse the four 2-bit fields
// and complement each" and a canonical k-mer is first met exactly when code == min(code, rc(code)).
#include "common.cuh"
#include <cstring>
#include <mutex>

namespace dbb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int revcomp_code(int code) {
  int r = 0;
  for (int p = 0; p < 4; ++p) {
    int b = (code >> (2 * p)) & 3;      // base at position 3-p
    r |= (3 - b) << (2 * (3 - p));      // complement, placed mirrored
  }
  return r;
}

static KmerTable build_table() {
  KmerTable t;
  int next = 0;
  for (int c = 0; c < DBB_NUM_KMERS; ++c) t.col_code2[c] = -1;
  for (int code = 0; code < 256; ++code) {
    int rc = revcomp_code(code);
    if (code <= rc) {                  // canonical, first seen now
      t.lut256[code] = (uint8_t)next;
      t.col_code1[next] = (int16_t)code;
      for (int p = 0; p < 4; ++p) t.names[next][p] = "ACGT"[(code >> (2 * (3 - p))) & 3];
      t.names[next][4] = 0;
      ++next;
    }
  }
  for (int code = 0; code < 256; ++code) {
    int rc = revcomp_code(code);
    if (code > rc) {
      t.lut256[code] = t.lut256[rc];
      t.col_code2[t.lut256[rc]] = (int16_t)code;
    }
  }
  return t;
}

const KmerTable& kmer_table() {
  static KmerTable t = build_table();
  return t;
}

}  // namespace dbb

DBB_EXPORT const char* dbb_last_error(void) { return dbb::g_err; }

DBB_EXPORT int dbb_version(void) { return 100; }

DBB_EXPORT int dbb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* free_bytes,
                               int64_t* total_bytes) {
  int dev = 0;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DBB_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
  size_t f = 0, t = 0;
  DBB_CUDA_TRY(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return DBB_OK;
}

DBB_EXPORT int dbb_kmer_lut(uint8_t* lut256) {
  DBB_REQUIRE(lut256 != nullptr, "lut256 is NULL");
  memcpy(lut256, dbb::kmer_table().lut256, 256);
  return DBB_OK;
}

DBB_EXPORT int dbb_kmer_names(char* names680) {
  DBB_REQUIRE(names680 != nullptr, "names is NULL");
  memcpy(names680, dbb::kmer_table().names, DBB_NUM_KMERS * 5);
  return DBB_OK;
}

DBB_EXPORT int64_t dbb_packed_blocks(int64_t seq_len) {
  return seq_len <= 0 ? 0 : (seq_len + DBB_BLOCK_BASES - 1) / DBB_BLOCK_BASES;
}

DBB_EXPORT int dbb_block_offsets(const int64_t* seq_len, int64_t n_seq, int64_t* blk_off) {
  DBB_REQUIRE(n_seq >= 0 && blk_off != nullptr && (n_seq == 0 || seq_len != nullptr), "bad arrays");
  int64_t acc = 0;
  for (int64_t s = 0; s < n_seq; ++s) {
    DBB_REQUIRE(seq_len[s] >= 0, "negative sequence length");
    blk_off[s] = acc;
    acc += dbb_packed_blocks(seq_len[s]);
  }
  blk_off[n_seq] = acc;
  return DBB_OK;
}
