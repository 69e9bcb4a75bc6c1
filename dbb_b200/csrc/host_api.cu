// Host-buffer entry points: the calls a non-CUDA consumer (the reference's Python modules through
// ctypes) binds.  They own their device scratch (inside a dbb_ctx, one per device), stream
// the input through the GPU in bounded batches and return with the results in the caller's buffers.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

namespace dbb {

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return DBB_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return DBB_ERR_NOMEM; }
  cap = bytes;
  return DBB_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

constexpr size_t PIN_CHUNK = 4u << 20;   // staging ring for pageable input: 8 chunks of 4 MB per lane
constexpr int PIN_SLOTS = 8;

struct Lane {            // one in-flight batch of stage 1
  cudaStream_t st = nullptr;
  DevBuf ascii, seq_off, blk_off, codes, valid, counts, freq;
  uint8_t* pin = nullptr;               // pinned ring (lazily allocated, only when a caller passes pageable memory)
  cudaEvent_t pin_ev[PIN_SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  dbb_count_plan* plan = nullptr;
  int64_t s0 = 0, s1 = 0;
  std::vector<int64_t> h_off, h_blk;
};
struct SigState { Lane lanes[2]; };
struct PairsHostState {
  DevBuf dense, sig, len, tdb, gc, gcm, gcl, gclo, gchi, cov, covl, covlo, covhi;
  DevBuf o_idx, o_dist, o_cnt, o_bp, o_td, o_gc, o_cov;
};

void free_sig_state(SigState* s) {
  if (!s) return;
  for (auto& L : s->lanes) {
    for (DevBuf* b : {&L.ascii, &L.seq_off, &L.blk_off, &L.codes, &L.valid, &L.counts, &L.freq}) b->release();
    if (L.plan) dbb_count_plan_destroy(L.plan);
    if (L.pin) cudaFreeHost(L.pin);
    for (auto& e : L.pin_ev) if (e) cudaEventDestroy(e);
    if (L.st) cudaStreamDestroy(L.st);
  }
  delete s;
}
void free_pairs_state(PairsHostState* s) {
  if (!s) return;
  for (DevBuf* b : {&s->dense, &s->sig, &s->len, &s->tdb, &s->gc, &s->gcm, &s->gcl, &s->gclo, &s->gchi, &s->cov, &s->covl, &s->covlo,
                    &s->covhi, &s->o_idx, &s->o_dist, &s->o_cnt, &s->o_bp, &s->o_td, &s->o_gc, &s->o_cov}) b->release();
  delete s;
}

CtxLock::CtxLock(dbb_ctx* c) : m(c->mu) { static_cast<std::mutex*>(m)->lock(); }
CtxLock::~CtxLock() { static_cast<std::mutex*>(m)->unlock(); }

static int make_ctx(int device, dbb_ctx** out) {
  dbb_ctx* c = new dbb_ctx();
  c->device = device;
  c->mu = new std::mutex();
  cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    delete static_cast<std::mutex*>(c->mu); delete c;
    return DBB_ERR_CUDA;
  }
  *out = c;
  return DBB_OK;
}

int ctx_resolve(dbb_ctx* given, dbb_ctx** out) {
  int dev = 0;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  if (given) {
    if (given->device != dev) {
      set_error("invalid argument: the context belongs to device %d, the current device is %d", given->device, dev);
      return DBB_ERR_INVALID;
    }
    *out = given;
    return DBB_OK;
  }
  if (dev < 0 || dev >= 64) { set_error("device ordinal out of range"); return DBB_ERR_UNSUPPORTED; }
  static std::mutex g_mu;
  static dbb_ctx* g_default[64] = {nullptr};       // never destroyed: no cudaFree after the runtime is torn down
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_default[dev]) {
    int rc = make_ctx(dev, &g_default[dev]);
    if (rc != DBB_OK) return rc;
  }
  *out = g_default[dev];
  return DBB_OK;
}

}  // namespace dbb

DBB_EXPORT int dbb_ctx_create(int device, dbb_ctx** ctx) {
  DBB_REQUIRE(ctx != nullptr, "NULL argument");
  *ctx = nullptr;
  int cur = 0;
  DBB_CUDA_TRY(cudaGetDevice(&cur));
  if (device < 0) device = cur;
  int n_dev = 0;
  DBB_CUDA_TRY(cudaGetDeviceCount(&n_dev));
  DBB_REQUIRE(device < n_dev, "no such device");
  if (device != cur) DBB_CUDA_TRY(cudaSetDevice(device));
  int rc = dbb::make_ctx(device, ctx);
  if (device != cur) cudaSetDevice(cur);
  return rc;
}

DBB_EXPORT int dbb_ctx_destroy(dbb_ctx* ctx) {
  if (!ctx) return DBB_OK;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != ctx->device) cudaSetDevice(ctx->device);
  if (ctx->st) cudaStreamSynchronize(ctx->st);
  dbb::free_sig_state(ctx->sig);
  dbb::free_pairs_state(ctx->pairs);
  dbb::free_knn_state(ctx->knn);
  if (ctx->st) cudaStreamDestroy(ctx->st);
  if (cur != ctx->device) cudaSetDevice(cur);
  delete static_cast<std::mutex*>(ctx->mu);
  delete ctx;
  return DBB_OK;
}

namespace {

using dbb::DevBuf;
using dbb::Lane;
using dbb::PIN_CHUNK;
using dbb::PIN_SLOTS;

#define DBB_TRY(expr) do { int _rc = (expr); if (_rc != DBB_OK) return _rc; } while (0)

// Host -> device copy of a range that may be PAGEABLE (a numpy array, a Python bytes object).  A pageable
// cudaMemcpyAsync is staged by the driver on the calling thread at ~10 GB/s; here a few threads copy 4 MB chunks into
// the lane's pinned ring and the chunks go up with asynchronous copies as they become ready, so the link stays busy.
// Pinned / registered memory takes the direct path.
int h2d_any(Lane& L, int device, void* dst, const uint8_t* src, size_t nbytes) {
  if (nbytes == 0) return DBB_OK;
  cudaPointerAttributes attr;
  const bool pageable = cudaPointerGetAttributes(&attr, src) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();                                        // (an unknown host pointer is reported as an error on old drivers)
  static const int n_threads = [] {
    const char* e = getenv("DBB_HOST_THREADS");
    int t = (e && *e) ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    return std::max(1, std::min(t, 32));
  }();
  if (!pageable || nbytes < 2 * PIN_CHUNK || n_threads <= 1) {
    DBB_CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, L.st));
    return DBB_OK;
  }
  if (!L.pin) {
    DBB_CUDA_TRY(cudaHostAlloc((void**)&L.pin, PIN_CHUNK * PIN_SLOTS, cudaHostAllocDefault));
    for (auto& e : L.pin_ev) DBB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int n_chunks = (int)((nbytes + PIN_CHUNK - 1) / PIN_CHUNK);
  std::vector<std::atomic<int>> ready(n_chunks), issued(n_chunks);
  for (int c = 0; c < n_chunks; ++c) { ready[c].store(0); issued[c].store(0); }
  std::atomic<int> next{0}, failed{0};
  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      const int c = next.fetch_add(1);
      if (c >= n_chunks || failed.load()) break;
      const int slot = c % PIN_SLOTS;
      if (c >= PIN_SLOTS) {                                  // the slot's previous chunk must have left for the device
        while (!issued[c - PIN_SLOTS].load(std::memory_order_acquire)) { if (failed.load()) return; std::this_thread::yield(); }
      }
      if (cudaEventSynchronize(L.pin_ev[slot]) != cudaSuccess) { failed.store(1); return; }
      const size_t o = (size_t)c * PIN_CHUNK, len = std::min(PIN_CHUNK, nbytes - o);
      memcpy(L.pin + (size_t)slot * PIN_CHUNK, src + o, len);
      ready[c].store(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < std::min(n_threads, n_chunks); ++t) pool.emplace_back(worker);
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < n_chunks && e == cudaSuccess; ++c) {
    while (!ready[c].load(std::memory_order_acquire)) { if (failed.load()) break; std::this_thread::yield(); }
    if (failed.load()) break;
    const int slot = c % PIN_SLOTS;
    const size_t o = (size_t)c * PIN_CHUNK, len = std::min(PIN_CHUNK, nbytes - o);
    e = cudaMemcpyAsync((uint8_t*)dst + o, L.pin + (size_t)slot * PIN_CHUNK, len, cudaMemcpyHostToDevice, L.st);
    if (e == cudaSuccess) e = cudaEventRecord(L.pin_ev[slot], L.st);
    issued[c].store(1, std::memory_order_release);
  }
  if (e != cudaSuccess) failed.store(1);
  for (int c = 0; c < n_chunks; ++c) issued[c].store(1, std::memory_order_release);     // release any waiting worker
  for (auto& th : pool) th.join();
  if (e != cudaSuccess || failed.load()) {
    dbb::set_error("staged host-to-device copy failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    return DBB_ERR_CUDA;
  }
  return DBB_OK;
}

int submit_batch(Lane& L, int device, const uint8_t* ascii, const int64_t* seq_off, int64_t s0, int64_t s1,
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
  DBB_TRY(h2d_any(L, device, L.ascii.p, ascii + byte0, (size_t)nbytes));
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
  // lanes (streams, device buffers, count plans) belong to the default context of the current device and only grow:
  // a per-sequence caller such as GenomicSignatures.seqSignature pays no allocation after its first call.
  dbb_ctx* ctx = nullptr;
  int rc = dbb::ctx_resolve(nullptr, &ctx);
  if (rc != DBB_OK) return rc;
  dbb::CtxLock lock(ctx);
  if (!ctx->sig) ctx->sig = new dbb::SigState();
  Lane* lanes = ctx->sig->lanes;
  const char* env = getenv("DBB_HOST_BATCH_MB");
  const int64_t batch_bytes = (env && *env ? atoll(env) : 256) << 20;
  for (int i = 0; i < 2; ++i) {
    Lane& L = lanes[i];
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
    rc = submit_batch(L, ctx->device, ascii, seq_off, s, e, counts != nullptr, freq != nullptr);
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
  // scratch of the current device's default context (grow-only), one host-buffer call at a time
  dbb_ctx* ctx = nullptr;
  int rc0 = dbb::ctx_resolve(nullptr, &ctx);
  if (rc0 != DBB_OK) return rc0;
  dbb::CtxLock lock(ctx);
  if (!ctx->pairs) ctx->pairs = new dbb::PairsHostState();
  dbb::PairsHostState& B = *ctx->pairs;
  DevBuf &dense = B.dense, &sig = B.sig, &len = B.len, &tdb = B.tdb, &gc = B.gc, &gcm = B.gcm, &gcl = B.gcl, &gclo = B.gclo, &gchi = B.gchi,
         &cov = B.cov, &covl = B.covl, &covlo = B.covlo, &covhi = B.covhi;
  DevBuf &o_idx = B.o_idx, &o_dist = B.o_dist, &o_cnt = B.o_cnt, &o_bp = B.o_bp, &o_td = B.o_td, &o_gc = B.o_gc, &o_cov = B.o_cov;
  cudaStream_t st = ctx->st;
  int rc = DBB_OK;
  dbb_pairs_input d = *in;
  dbb_pairs_output o = *out;
  d.col_id = nullptr; d.item_order = nullptr; d.tile_off = nullptr; d.tile_list = nullptr;   // device-path options
  d.split_all = 0; d.n_active_row_tiles = 0; d.split_tiles = 0;
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
