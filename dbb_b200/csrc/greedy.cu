// "Next" row N3: greedy core growth, Greedy.__greedy, greedy.py:167-251.
//
// Reference loop: walk the contigs in the given order; an unprocessed contig seeds a core; repeatedly, among the
// LATER unprocessed contigs, take the one closest to the core in coverage (|cov - core.cov|; a later contig replaces
// an earlier one at equal distance, :204-218) that passes withinDistGC(contig, core), withinDistCov(contig, core)
// and L1(contig.sig, core.sig) < boundDistTD(contig) (distributions.py:75-127, fixedSeqLen=None), add it and update
// the core's length / GC / coverage / signature as running length-weighted means (:229-234); when nothing qualifies
// the core is kept if its length exceeds minBinSize, otherwise its contigs are released (:239-246).
//
// The inner scan over N contigs is the data-parallel part: one kernel per growth step evaluates all candidates
// against the current core (scalar predicates lane-parallel, the FP64 L1 distance warp-cooperatively for the few that
// pass both) and reduces to (min |cov diff|, last index attaining it); the block that finishes last also marks the
// winner as processed.  The outer loop is inherently sequential (each step depends on the updated core); it runs on
// the host in the same IEEE double arithmetic, one launch and one 16-byte read-back per step, the core's signature
// travelling by value in the kernel parameters.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace {

struct CoreState {
  double sig[DBB_NUM_KMERS];
  double gc, cov;
  int32_t gc_mbin;
};

struct StepResult { long long idx; double diff; };

struct GreedyDev {
  int64_t n;
  const double* sig; const double* gc; const double* cov; const double* td_bound;
  const int32_t* gc_lbin; const int32_t* cov_lbin;
  const double* gc_lo; const double* gc_hi; int32_t gc_nl;
  const double* cov_lo; const double* cov_hi;
  uint8_t* processed;
  unsigned int* counter; double* part_diff; long long* part_idx; StepResult* result;
};

__device__ __forceinline__ bool better(double d, long long i, double bd, long long bi) {
  return i >= 0 && (bi < 0 || d < bd || (d == bd && i > bi));      // ties: the later contig wins (greedy.py:204-218)
}

__global__ void __launch_bounds__(256) greedy_step_kernel(const GreedyDev D, const int64_t first, const __grid_constant__ CoreState core) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = first + ((int64_t)blockIdx.x * 8 + warp) * 32 + lane;
  double cs[5];
  #pragma unroll
  for (int q = 0; q < 5; ++q) { const int k = lane + 32 * q; cs[q] = k < DBB_NUM_KMERS ? core.sig[k] : 0.0; }
  bool pass = false;
  double diff = 0.0;
  if (c < D.n && !D.processed[c]) {
    diff = fabs(D.cov[c] - core.cov);                                // :203
    if (!(diff > 1e12)) {                                            // :198,204 (closestAbsCovDiff starts at 1e12)
      const int l = D.gc_lbin[c];
      const double gdiff = D.gc[c] - core.gc;                        // distributions.py:92
      const bool p_gc = gdiff >= D.gc_lo[core.gc_mbin * D.gc_nl + l] && gdiff <= D.gc_hi[core.gc_mbin * D.gc_nl + l];
      const double eff = core.cov > 0 ? core.cov : 0.1;              // distributions.py:121-125
      const double cdiff = (D.cov[c] - eff) * 100.0 / eff;
      const int cl = D.cov_lbin[c];
      pass = p_gc && cdiff >= D.cov_lo[cl] && cdiff <= D.cov_hi[cl];
    }
  }
  uint32_t todo = __ballot_sync(0xffffffffu, pass);
  double bd = 0.0;
  long long bi = -1;
  while (todo) {                                                     // TD distance only for contigs that passed both
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t cj = c - lane + j;
    const double* s = D.sig + cj * DBB_NUM_KMERS;
    double part = 0.0;
    #pragma unroll
    for (int q = 0; q < 5; ++q) { const int k = lane + 32 * q; if (k < DBB_NUM_KMERS) part += fabs(s[k] - cs[q]); }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    const double dj = __shfl_sync(0xffffffffu, diff, j);
    if (part < D.td_bound[cj] && better(dj, cj, bd, bi)) { bd = dj; bi = cj; }    // distributions.py:103-104
  }
  __shared__ double s_d[8];
  __shared__ long long s_i[8];
  __shared__ bool s_last;
  if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) if (better(s_d[w], s_i[w], bd, bi)) { bd = s_d[w]; bi = s_i[w]; }
    D.part_diff[blockIdx.x] = bd; D.part_idx[blockIdx.x] = bi;
    __threadfence();
    s_last = atomicAdd(D.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // the last block to finish reduces the partial results and marks the winner
  __threadfence();
  bd = 0.0; bi = -1;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
    const double d = D.part_diff[b];
    const long long i = D.part_idx[b];
    if (better(d, i, bd, bi)) { bd = d; bi = i; }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(od, oi, bd, bi)) { bd = od; bi = oi; }
  }
  if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) if (better(s_d[w], s_i[w], bd, bi)) { bd = s_d[w]; bi = s_i[w]; }
    D.result->idx = bi; D.result->diff = bd;
    if (bi >= 0) D.processed[bi] = 1;
    *D.counter = 0u;
  }
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <typename T> int up(const T* h, size_t n, cudaStream_t st) {
    if (cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) { dbb::set_error("cudaMalloc failed in dbb_greedy_host"); return DBB_ERR_NOMEM; }
    if (n && cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) { dbb::set_error("H2D copy failed in dbb_greedy_host"); return DBB_ERR_CUDA; }
    return DBB_OK;
  }
  int zero(size_t bytes, cudaStream_t st) {
    if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) { dbb::set_error("cudaMalloc failed in dbb_greedy_host"); return DBB_ERR_NOMEM; }
    if (cudaMemsetAsync(p, 0, std::max<size_t>(bytes, 16), st) != cudaSuccess) { dbb::set_error("memset failed in dbb_greedy_host"); return DBB_ERR_CUDA; }
    return DBB_OK;
  }
};

}  // namespace

DBB_EXPORT int dbb_greedy_host(const dbb_greedy_input* in, int64_t min_bin_size, int32_t* bin_id, int64_t* n_cores_out,
                               int64_t* core_len, double* core_gc, double* core_cov, double* core_sig, int64_t* n_steps_out) {
  DBB_REQUIRE(in && bin_id && n_cores_out, "NULL argument");
  const int64_t n = in->n;
  DBB_REQUIRE(n >= 0, "negative size");
  *n_cores_out = 0;
  if (n_steps_out) *n_steps_out = 0;
  if (n == 0) return DBB_OK;
  DBB_REQUIRE(in->sig && in->length && in->gc && in->cov && in->td_bound && in->gc_lbin && in->cov_lbin, "incomplete contig inputs");
  DBB_REQUIRE(in->gc_mean_keys && in->gc_lo && in->gc_hi && in->cov_lo && in->cov_hi && in->gc_nm > 0 && in->gc_nl > 0 && in->cov_nl > 0,
              "incomplete tables");
  cudaStream_t st;
  DBB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  Buf sig, gc, cov, tdb, gl, cl, glo, ghi, clo, chi, proc, counter, pd, pi, res;
  GreedyDev D;
  int rc = DBB_OK;
  StepResult* h_res = nullptr;
  do {
    const size_t N = (size_t)n;
    const size_t max_blocks = (N + 255) / 256;
#define UP(buf, field, count) if ((rc = buf.up(in->field, count, st)) != DBB_OK) break;
    UP(sig, sig, N * DBB_NUM_KMERS) UP(gc, gc, N) UP(cov, cov, N) UP(tdb, td_bound, N) UP(gl, gc_lbin, N) UP(cl, cov_lbin, N)
    UP(glo, gc_lo, (size_t)in->gc_nm * in->gc_nl) UP(ghi, gc_hi, (size_t)in->gc_nm * in->gc_nl)
    UP(clo, cov_lo, (size_t)in->cov_nl) UP(chi, cov_hi, (size_t)in->cov_nl)
#undef UP
    if ((rc = proc.zero(N, st)) != DBB_OK || (rc = counter.zero(16, st)) != DBB_OK || (rc = pd.zero(max_blocks * 8, st)) != DBB_OK ||
        (rc = pi.zero(max_blocks * 8, st)) != DBB_OK || (rc = res.zero(sizeof(StepResult), st)) != DBB_OK) break;
    if (cudaMallocHost((void**)&h_res, sizeof(StepResult)) != cudaSuccess) { dbb::set_error("cudaMallocHost failed"); rc = DBB_ERR_NOMEM; break; }
    D.n = n; D.sig = (const double*)sig.p; D.gc = (const double*)gc.p; D.cov = (const double*)cov.p; D.td_bound = (const double*)tdb.p;
    D.gc_lbin = (const int32_t*)gl.p; D.cov_lbin = (const int32_t*)cl.p; D.gc_lo = (const double*)glo.p; D.gc_hi = (const double*)ghi.p;
    D.gc_nl = in->gc_nl; D.cov_lo = (const double*)clo.p; D.cov_hi = (const double*)chi.p; D.processed = (uint8_t*)proc.p;
    D.counter = (unsigned int*)counter.p; D.part_diff = (double*)pd.p; D.part_idx = (long long*)pi.p; D.result = (StepResult*)res.p;

    std::vector<uint8_t> processed(N, 0);
    std::vector<int64_t> members;
    int64_t n_cores = 0, n_steps = 0;
    for (int64_t i = 0; i < n; ++i) bin_id[i] = -1;                    // Greedy.UNBINNED
    CoreState core;
    cudaError_t e = cudaSuccess;
    for (int64_t ci = 0; ci < n && e == cudaSuccess; ++ci) {            // greedy.py:177
      if (processed[ci]) continue;
      processed[ci] = 1;                                              // (indices <= ci are never candidates again)
      bin_id[ci] = (int32_t)n_cores;
      members.assign(1, ci);
      int64_t core_length = in->length[ci];                           // :190
      core.gc = in->gc[ci]; core.cov = in->cov[ci];
      for (int k = 0; k < DBB_NUM_KMERS; ++k) core.sig[k] = in->sig[(size_t)ci * DBB_NUM_KMERS + k];
      while (true) {                                                  // :194
        // distributions.py:76 -- nearest mean-GC key of the evolving core (ascending keys, lowest wins ties)
        int m = 0;
        double best = fabs(in->gc_mean_keys[0] - core.gc);
        for (int k = 1; k < in->gc_nm; ++k) { const double d = fabs(in->gc_mean_keys[k] - core.gc); if (d < best) { best = d; m = k; } }
        core.gc_mbin = m;
        long long idx = -1;
        if (ci + 1 < n) {
          const int64_t cand = n - (ci + 1);
          const unsigned grid = (unsigned)((cand + 255) / 256);
          greedy_step_kernel<<<grid, 256, 0, st>>>(D, ci + 1, core);
          e = cudaMemcpyAsync(h_res, D.result, sizeof(StepResult), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          if (e != cudaSuccess) break;
          idx = h_res->idx;
          ++n_steps;
        }
        if (idx >= 0) {                                               // :221-234
          processed[idx] = 1;
          bin_id[idx] = (int32_t)n_cores;
          members.push_back(idx);
          core_length += in->length[idx];
          const double weight = (double)in->length[idx] / (double)core_length;
          const double rest = 1.0 - weight;
          core.gc = in->gc[idx] * weight + core.gc * rest;
          core.cov = in->cov[idx] * weight + core.cov * rest;
          const double* s = in->sig + (size_t)idx * DBB_NUM_KMERS;
          for (int k = 0; k < DBB_NUM_KMERS; ++k) core.sig[k] = s[k] * weight + core.sig[k] * rest;
        } else {                                                      // :235-246
          if (core_length > min_bin_size) {
            if (core_len) core_len[n_cores] = core_length;
            if (core_gc) core_gc[n_cores] = core.gc;
            if (core_cov) core_cov[n_cores] = core.cov;
            if (core_sig) for (int k = 0; k < DBB_NUM_KMERS; ++k) core_sig[(size_t)n_cores * DBB_NUM_KMERS + k] = core.sig[k];
            ++n_cores;
          } else {
            for (int64_t mbr : members) {
              processed[mbr] = 0;
              bin_id[mbr] = -1;
              if (mbr > ci && e == cudaSuccess) e = cudaMemsetAsync((uint8_t*)proc.p + mbr, 0, 1, st);
            }
          }
          break;
        }
      }
    }
    if (e != cudaSuccess) { dbb::set_error("dbb_greedy_host failed: %s", cudaGetErrorString(e)); rc = DBB_ERR_CUDA; break; }
    *n_cores_out = n_cores;
    if (n_steps_out) *n_steps_out = n_steps;
  } while (0);
  cudaStreamSynchronize(st);
  if (h_res) cudaFreeHost(h_res);
  cudaStreamDestroy(st);
  return rc;
}
