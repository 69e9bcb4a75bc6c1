// Shared helpers for libdbbcuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/dbb_cuda.h"

#define DBB_EXPORT extern "C" __attribute__((visibility("default")))

namespace dbb {

void set_error(const char* fmt, ...);

// canonical tetramer table (built once on the host, table.cu)
struct KmerTable {
  uint8_t lut256[256];     // code -> column
  int16_t col_code1[DBB_NUM_KMERS];   // the canonical code of the column
  int16_t col_code2[DBB_NUM_KMERS];   // its reverse complement, or -1 for palindromes
  char names[DBB_NUM_KMERS][5];
};
const KmerTable& kmer_table();

// internal: rebuild a count plan in place (count.cu), used by the host-buffer entry points
int count_plan_rebuild(dbb_count_plan* plan, const int64_t* h_blk_off, int64_t n_seq, cudaStream_t st);

}  // namespace dbb

#define DBB_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      dbb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DBB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define DBB_REQUIRE(cond, msg)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      dbb::set_error("invalid argument: %s", msg);          \
      return DBB_ERR_INVALID;                               \
    }                                                       \
  } while (0)
