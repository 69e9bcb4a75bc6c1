// Host-buffer entry points: the calls a non-CUDA consumer (the reference's Python modules through
// ctypes) binds.  They own their device scratch, stream the input through the
// GPU in bounded batches and return with the results in the caller's buffers.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return DBB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { dbb::set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return DBB_ERR_NOMEM; }
    cap = bytes;
    return DBB_OK;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

struct Lane {            // one in-flight batch
  cudaStream_t st = nullptr;
  DevBuf ascii, seq_off, blk_off, codes, valid, counts, freq;
  dbb_count_plan* plan = nullptr;
  int64_t s0 = 0, s1 = 0;
  std::vector<int64_t> h_off, h_blk;
};

#define DBB_TRY(expr) do { int _rc = (expr); if (_rc != DBB_OK) return _rc; } while (0)

int submit_batch(Lane& L, const uint8_t* ascii, const int64_t* seq_off, int64_t s0, int64_t s1,
                 bool want_counts, bool want_freq) {
  const int64_t ns = s1 - s0;
  const int64_t byte0 = seq_off[s0], nbytes = seq_off[s1] - byte0;
  L.s0 = s0; L.s1 = s1;
  L.h_off.resize(ns + 1);
  L.h_blk.resize(ns + 1);
  int64_t acc = 0;
  for (int64_t i = 0; i < ns; ++i) {
    L.h_off[i] = seq_off[s0 + i] - byte0;
    L.h_blk[i] = acc;
    acc += dbb_packed_blocks(seq_off[s0 + i + 1] - seq_off[s0 + i]);
  }
  L.h_off[ns] = nbytes;
  L.h_blk[ns] = acc;
  const int64_t nblk = acc;
  DBB_TRY(L.ascii.reserve((size_t)std::max<int64_t>(nbytes, 16)));
  DBB_TRY(L.seq_off.reserve((ns + 1) * 8));
  DBB_TRY(L.blk_off.reserve((ns + 1) * 8));
  DBB_TRY(L.codes.reserve((size_t)std::max<int64_t>(nblk, 1) * 16 + 16));
  DBB_TRY(L.valid.reserve((size_t)std::max<int64_t>(nblk, 1) * 8 + 16));
  if (want_counts) DBB_TRY(L.counts.reserve((size_t)ns * DBB_NUM_KMERS * 4));
  if (want_freq) DBB_TRY(L.freq.reserve((size_t)ns * DBB_NUM_KMERS * 4));
  if (!L.plan) DBB_TRY(dbb_count_plan_create(L.h_blk.data(), ns, &L.plan));
  else DBB_TRY(dbb::count_plan_rebuild(L.plan, L.h_blk.data(), ns, L.st));
  if (nbytes) DBB_CUDA_TRY(cudaMemcpyAsync(L.ascii.p, ascii + byte0, nbytes, cudaMemcpyHostToDevice, L.st));
  DBB_CUDA_TRY(cudaMemcpyAsync(L.seq_off.p, L.h_off.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, L.st));
  DBB_CUDA_TRY(cudaMemcpyAsync(L.blk_off.p, L.h_blk.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, L.st));
  DBB_TRY(dbb_pack_ascii_dev((const uint8_t*)L.ascii.p, (const int64_t*)L.seq_off.p, (const int64_t*)L.blk_off.p, ns,
                             nblk, L.codes.p, L.valid.p, L.st));
  DBB_TRY(dbb_tetra_count_dev(L.plan, L.codes.p, L.valid.p, want_counts ? (uint32_t*)L.counts.p : nullptr,
                              want_freq ? (float*)L.freq.p : nullptr, DBB_NUM_KMERS, L.st));
  return DBB_OK;
}

int collect_batch(Lane& L, uint32_t* counts, float* freq) {
  const int64_t ns = L.s1 - L.s0;
  if (counts) DBB_CUDA_TRY(cudaMemcpyAsync(counts + L.s0 * DBB_NUM_KMERS, L.counts.p, (size_t)ns * DBB_NUM_KMERS * 4, cudaMemcpyDeviceToHost, L.st));
  if (freq) DBB_CUDA_TRY(cudaMemcpyAsync(freq + L.s0 * DBB_NUM_KMERS, L.freq.p, (size_t)ns * DBB_NUM_KMERS * 4, cudaMemcpyDeviceToHost, L.st));
  DBB_CUDA_TRY(cudaStreamSynchronize(L.st));
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_tetra_signatures_host(const uint8_t* ascii, const int64_t* seq_off, int64_t n_seq,
                                         uint32_t* counts, float* freq) {
  DBB_REQUIRE(n_seq >= 0, "negative n_seq");
  if (n_seq == 0) return DBB_OK;
  DBB_REQUIRE(seq_off != nullptr && (counts || freq), "NULL argument");
  DBB_REQUIRE(ascii != nullptr || seq_off[n_seq] == seq_off[0], "ascii is NULL");
  for (int64_t s = 0; s < n_seq; ++s) DBB_REQUIRE(seq_off[s + 1] >= seq_off[s], "seq_off must be non-decreasing");

  // batches of ~batch_bytes of sequence, two in flight (copy of batch b+1 overlaps compute of batch b).  The
  // lanes (streams, device buffers, count plans) live for the whole process and only grow: a per-sequence caller
  // such as GenomicSignatures.seqSignature pays no allocation after its first call.
  static std::mutex mu;
  static Lane lanes[2];
  std::lock_guard<std::mutex> lock(mu);
  const char* env = getenv("DBB_HOST_BATCH_MB");
  const int64_t batch_bytes = (env && *env ? atoll(env) : 256) << 20;
  int rc = DBB_OK;
  for (auto& L : lanes) {
    if (L.st) continue;
    cudaError_t e = cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { dbb::set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); L.st = nullptr; return DBB_ERR_CUDA; }
  }
  int64_t s = 0;
  int cur = 0;
  bool pending[2] = {false, false};
  while (rc == DBB_OK && s < n_seq) {
    int64_t e = s;
    const int64_t b0 = seq_off[s];
    while (e < n_seq && (e == s || seq_off[e + 1] - b0 <= batch_bytes)) ++e;
    Lane& L = lanes[cur];
    if (pending[cur]) { rc = collect_batch(L, counts, freq); pending[cur] = false; if (rc != DBB_OK) break; }
    rc = submit_batch(L, ascii, seq_off, s, e, counts != nullptr, freq != nullptr);
    if (rc != DBB_OK) break;
    pending[cur] = true;
    cur ^= 1;
    s = e;
  }
  for (int i = 0; i < 2; ++i) {
    Lane& L = lanes[cur ^ i];
    if (pending[cur ^ i]) {
      if (rc == DBB_OK) rc = collect_batch(L, counts, freq);
      else cudaStreamSynchronize(L.st);
    }
  }
  return rc;
}

namespace {

template <typename T>
int to_dev(DevBuf& b, const T* h, size_t count, cudaStream_t st) {
  int rc = b.reserve(std::max<size_t>(count * sizeof(T), 16));
  if (rc != DBB_OK) return rc;
  if (count) DBB_CUDA_TRY(cudaMemcpyAsync(b.p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_pairs_host(const dbb_pairs_input* in, int64_t row0, int64_t row1, const dbb_pairs_output* out) {
  DBB_REQUIRE(in && out, "NULL argument struct");
  DBB_REQUIRE(in->n >= 0 && row0 >= 0 && row1 >= row0 && row1 <= in->n, "bad row range");
  if (row1 == row0 || in->n == 0) return DBB_OK;
  DBB_REQUIRE(in->sig && in->length, "sig/length are required");
  const size_t n = (size_t)in->n, rows = (size_t)(row1 - row0);
  // process-lifetime scratch (grow-only), one host-buffer call at a time
  static std::mutex mu;
  static cudaStream_t st = nullptr;
  static DevBuf dense, sig, len, tdb, gc, gcm, gcl, gclo, gchi, cov, covl, covlo, covhi;
  static DevBuf o_idx, o_dist, o_cnt, o_bp, o_td, o_gc, o_cov;
  std::lock_guard<std::mutex> lock(mu);
  if (!st) DBB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = DBB_OK;
  dbb_pairs_input d = *in;
  dbb_pairs_output o = *out;
  d.col_id = nullptr; d.item_order = nullptr; d.tile_off = nullptr; d.tile_list = nullptr;   // device-path options
  d.split_all = 0; d.n_active_row_tiles = 0;
  do {
    // NOTE: on the host side `sig` is dense [n][136]
    if ((rc = to_dev(dense, in->sig, n * DBB_NUM_KMERS, st)) != DBB_OK) break;
    if ((rc = sig.reserve(n * DBB_SIG_STRIDE * 4)) != DBB_OK) break;
    if ((rc = dbb_pad_signatures_dev((const float*)dense.p, in->n, (float*)sig.p, st)) != DBB_OK) break;
    d.sig = (const float*)sig.p;
    if ((rc = to_dev(len, in->length, n, st)) != DBB_OK) break;
    d.length = (const int32_t*)len.p;
    if (in->td_bound) { if ((rc = to_dev(tdb, in->td_bound, n, st)) != DBB_OK) break; d.td_bound = (const float*)tdb.p; }
    if (in->gc) {
      if (!(in->gc_mbin && in->gc_lbin && in->gc_lo && in->gc_hi && in->gc_nm > 0 && in->gc_nl > 0)) { dbb::set_error("incomplete GC inputs"); rc = DBB_ERR_INVALID; break; }
      if ((rc = to_dev(gc, in->gc, n, st)) != DBB_OK) break;
      if ((rc = to_dev(gcm, in->gc_mbin, n, st)) != DBB_OK) break;
      if ((rc = to_dev(gcl, in->gc_lbin, n, st)) != DBB_OK) break;
      if ((rc = to_dev(gclo, in->gc_lo, (size_t)in->gc_nm * in->gc_nl, st)) != DBB_OK) break;
      if ((rc = to_dev(gchi, in->gc_hi, (size_t)in->gc_nm * in->gc_nl, st)) != DBB_OK) break;
      d.gc = (const double*)gc.p; d.gc_mbin = (const int32_t*)gcm.p; d.gc_lbin = (const int32_t*)gcl.p;
      d.gc_lo = (const double*)gclo.p; d.gc_hi = (const double*)gchi.p;
    }
    if (in->cov) {
      if (!(in->cov_lbin && in->cov_lo && in->cov_hi && in->cov_nl > 0)) { dbb::set_error("incomplete coverage inputs"); rc = DBB_ERR_INVALID; break; }
      if ((rc = to_dev(cov, in->cov, n, st)) != DBB_OK) break;
      if ((rc = to_dev(covl, in->cov_lbin, n, st)) != DBB_OK) break;
      if ((rc = to_dev(covlo, in->cov_lo, (size_t)in->cov_nl, st)) != DBB_OK) break;
      if ((rc = to_dev(covhi, in->cov_hi, (size_t)in->cov_nl, st)) != DBB_OK) break;
      d.cov = (const double*)cov.p; d.cov_lbin = (const int32_t*)covl.p;
      d.cov_lo = (const double*)covlo.p; d.cov_hi = (const double*)covhi.p;
    }
    const size_t kk = (size_t)std::max(out->k, 0);
    if (out->knn_idx) { if ((rc = o_idx.reserve(std::max<size_t>(rows * kk * 4, 16))) != DBB_OK) break; o.knn_idx = (int32_t*)o_idx.p; }
    if (out->knn_dist) { if ((rc = o_dist.reserve(std::max<size_t>(rows * kk * 4, 16))) != DBB_OK) break; o.knn_dist = (float*)o_dist.p; }
    if (out->count) { if ((rc = o_cnt.reserve(rows * 4)) != DBB_OK) break; o.count = (int32_t*)o_cnt.p; }
    if (out->neighbour_bp) { if ((rc = o_bp.reserve(rows * 8)) != DBB_OK) break; o.neighbour_bp = (int64_t*)o_bp.p; }
    const size_t mw = (size_t)out->mask_words;
    if (out->td_mask) { if ((rc = o_td.reserve(rows * mw * 4)) != DBB_OK) break; o.td_mask = (uint32_t*)o_td.p; }
    if (out->gc_mask) { if ((rc = o_gc.reserve(rows * mw * 4)) != DBB_OK) break; o.gc_mask = (uint32_t*)o_gc.p; }
    if (out->cov_mask) { if ((rc = o_cov.reserve(rows * mw * 4)) != DBB_OK) break; o.cov_mask = (uint32_t*)o_cov.p; }
    if ((rc = dbb_pairs_dev(&d, row0, row1, &o, st)) != DBB_OK) break;
    cudaError_t e = cudaSuccess;
    auto back = [&](void* h, const DevBuf& b, size_t bytes) {
      if (h && bytes && e == cudaSuccess) e = cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st);
    };
    back(out->knn_idx, o_idx, rows * kk * 4);
    back(out->knn_dist, o_dist, rows * kk * 4);
    back(out->count, o_cnt, rows * 4);
    back(out->neighbour_bp, o_bp, rows * 8);
    back(out->td_mask, o_td, rows * mw * 4);
    back(out->gc_mask, o_gc, rows * mw * 4);
    back(out->cov_mask, o_cov, rows * mw * 4);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { dbb::set_error("pairs_host failed: %s", cudaGetErrorString(e)); rc = DBB_ERR_CUDA; }
  } while (0);
  cudaStreamSynchronize(st);
  return rc;
}
