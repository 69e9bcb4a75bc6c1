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

// grow-only device buffer owned by a context (no destructor: a context is freed explicitly by dbb_ctx_destroy, the
// per-device default contexts live until the process ends and are never torn down behind the CUDA runtime's back)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
};

// per-context state of the host entry points, defined where it is used
struct SigState;         // host_api.cu : dbb_tetra_signatures_host
struct PairsHostState;   // host_api.cu : dbb_pairs_host
struct KnnState;         // knn.cu      : dbb_knn_dev / dbb_knn_host
void free_sig_state(SigState*);
void free_pairs_state(PairsHostState*);
void free_knn_state(KnnState*);

}  // namespace dbb

// One context = one device: its stream, device scratch and cached plans.  The host entry points without a context
// argument use the default context of the CURRENT device (created on first use), so a process that switches devices
// between calls never launches on another device's stream or buffers.
struct dbb_ctx {
  int device = 0;
  cudaStream_t st = nullptr;
  dbb::SigState* sig = nullptr;
  dbb::PairsHostState* pairs = nullptr;
  dbb::KnnState* knn = nullptr;
  void* mu = nullptr;            // std::mutex*: one host-buffer call at a time per context
};

namespace dbb {
// NULL -> the default context of the current device; otherwise `given`, which must belong to the current device
int ctx_resolve(dbb_ctx* given, dbb_ctx** out);
struct CtxLock { explicit CtxLock(dbb_ctx* c); ~CtxLock(); void* m; };
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
