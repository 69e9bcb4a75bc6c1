// Host-buffer entry points: the calls a non-CUDA consumer (the reference's Python modules through
// ctypes) binds.  They own their device scratch (inside a dbb_ctx, one per device), stream
// the input through the GPU in bounded batches and return with the results in the caller's buffers.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
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

void host_pack_sequence(const uint8_t* seq, int64_t len, uint32_t* codes, uint64_t* valid);     // hostpack.cpp

struct PinBuf {          // grow-only pinned host buffer
  uint8_t* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return DBB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaHostAlloc((void**)&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return DBB_ERR_NOMEM; }
    cap = want;
    return DBB_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A few persistent host threads (the packers of submit_batch).  start() hands every thread the same job, wait() returns
// when all of them are back.
struct HostPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv, cv_done;
  std::function<void(int)> job;
  uint64_t gen = 0;
  int running = 0;
  bool stop = false;
  explicit HostPool(int n) {
    for (int t = 0; t < n; ++t) th.emplace_back([this, t] {
      uint64_t seen = 0;
      for (;;) {
        std::function<void(int)> j;
        {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [&] { return stop || gen != seen; });
          if (stop) return;
          seen = gen;
          j = job;
        }
        j(t);
        { std::lock_guard<std::mutex> lk(m); if (--running == 0) cv_done.notify_all(); }
      }
    });
  }
  void start(std::function<void(int)> j) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(j); running = (int)th.size(); ++gen;
    cv.notify_all();
  }
  void wait() { std::unique_lock<std::mutex> lk(m); cv_done.wait(lk, [&] { return running == 0; }); }
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
};

constexpr int DMA_DEPTH = 3;             // ASCII copies in flight on the copy engine

struct Lane {            // one in-flight batch of stage 1
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;           // copies of the host-packed planes (runs beside the ASCII copies of st)
  cudaEvent_t ev2 = nullptr, dma_ev[DMA_DEPTH] = {nullptr, nullptr, nullptr};
  PinBuf pk_codes, pk_valid;            // host-packed planes of the batch, same geometry as codes / valid
  std::vector<int64_t> unit;            // scaffold boundaries of the ~1 MB work units of the batch
  std::unique_ptr<std::atomic<uint8_t>[]> unit_done;
  size_t unit_cap = 0;
  DevBuf ascii, seq_off, blk_off, codes, valid, counts, freq;
  uint8_t* pin = nullptr;               // pinned ring (lazily allocated, only when a caller passes pageable memory)
  cudaEvent_t pin_ev[PIN_SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  dbb_count_plan* plan = nullptr;
  int64_t s0 = 0, s1 = 0;
  std::vector<int64_t> h_off, h_blk;
};
struct SigState {
  Lane lanes[2];
  HostPool* pool = nullptr;
  int pool_threads = -1;
  int64_t ascii_bytes = 0, plane_bytes = 0;      // what the last dbb_tetra_signatures_host call sent up by either route
};
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
    L.pk_codes.release(); L.pk_valid.release();
    for (auto& e : L.dma_ev) if (e) cudaEventDestroy(e);
    if (L.ev2) cudaEventDestroy(L.ev2);
    if (L.st2) cudaStreamDestroy(L.st2);
    if (L.st) cudaStreamDestroy(L.st);
  }
  delete s->pool;
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
using dbb::DMA_DEPTH;
using dbb::HostPool;
using dbb::Lane;
using dbb::SigState;
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

// Packer threads of dbb_tetra_signatures_host: DBB_HOST_PACK_THREADS, else the hardware threads this process can count on
// (all of them divided by LOCAL_WORLD_SIZE when a launcher started one process per GPU) minus one for the caller,
// at most 16.  0 switches the host packing off.
int pack_threads() {
  static const int n = [] {
    const char* e = getenv("DBB_HOST_PACK_THREADS");
    if (e && *e) return std::max(0, std::min(atoi(e), 64));
    const char* m = getenv("DBB_HOST_PACK");
    if (m && *m == '0') return 0;
    int hw = (int)std::thread::hardware_concurrency();
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const int ranks = (lw && atoi(lw) > 0) ? atoi(lw) : 1;
    return std::max(0, std::min(hw / ranks - 1, 16));
  }();
  return n;
}

constexpr int64_t UNIT_BYTES = 1 << 20;        // work unit of the hybrid upload
constexpr int64_t HYBRID_MIN_BYTES = 8 << 20;  // smaller batches go up as ASCII in one copy

// Upload of one batch by BOTH routes at once.  The batch is cut into units of whole scaffolds (~1 MB of ASCII).  The
// calling thread feeds the copy engine with ASCII units from the FRONT of the batch, DMA_DEPTH copies in flight, so it
// claims units exactly as fast as the link moves them; the packer threads claim units from the BACK, pack them into the
// pinned planes (0.375 bytes per base) and the calling thread sends finished runs of them up on a second stream.  When
// the two fronts meet the batch is up: scaffolds [0, *n_ascii) arrived as ASCII (the pack kernel handles them), the rest
// as planes.
// Used for PAGEABLE input (numpy arrays, Python bytes: what GenomicSignatures.signatures hands over), where the ASCII
// route is the driver's staged copy (~10 GB/s): C2 from pageable buffers 30.8 -> 43.6 Gbp/s with 15 packer threads
// (the pinned ring of h2d_any, which this replaces for large batches, copies every byte twice on the host).  PINNED
// input goes up as ASCII in one copy: measured on the same box, the mixed upload only moves the split (4 / 8 / 16
// packers carry 33 / 48 / 60 % of the bases) while the sum of the two routes stays at ~50 GB/s of input read from host
// memory, i.e. 49-50 Gbp/s against 51.7 for the plain copy -- the pinned line is bound by reading the input once, not
// by the bytes on the link.  DBB_HOST_PACK=0: ASCII only; =all: packers only; =hybrid: mixed upload for pinned input too.
int hybrid_upload(Lane& L, SigState& S, int device, const uint8_t* base, const uint8_t* const* ptrs, int64_t ns, int64_t nbytes,
                  int64_t* n_ascii) {
  *n_ascii = ns;
  const int n_thr = pack_threads();
  static const int mode = [] {
    const char* m = getenv("DBB_HOST_PACK");
    return !m ? 0 : strcmp(m, "all") == 0 ? 2 : strcmp(m, "hybrid") == 0 ? 1 : 0;
  }();
  // scattered sequences (ptrs: one host pointer per scaffold, dbb_tetra_signatures_host_ptrs) have no ASCII route: a copy
  // per scaffold would be thousands of small transfers; they all go through the packers
  const bool pack_all = mode == 2 || ptrs != nullptr;
  if (!ptrs) {
    cudaPointerAttributes attr;
    const bool pageable = cudaPointerGetAttributes(&attr, base) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();                                        // (an unknown host pointer is reported as an error on old drivers)
    if (n_thr <= 0 || (!pack_all && nbytes < HYBRID_MIN_BYTES) || (!pageable && mode == 0))
      return h2d_any(L, device, L.ascii.p, base, (size_t)nbytes);
  }
  if (n_thr > 0 && (!S.pool || S.pool_threads != n_thr)) { delete S.pool; S.pool = new HostPool(n_thr); S.pool_threads = n_thr; }
  if (!L.st2) {
    DBB_CUDA_TRY(cudaStreamCreateWithFlags(&L.st2, cudaStreamNonBlocking));
    DBB_CUDA_TRY(cudaEventCreateWithFlags(&L.ev2, cudaEventDisableTiming));
    for (auto& e : L.dma_ev) DBB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int64_t nblk = L.h_blk[ns];
  DBB_TRY(L.pk_codes.reserve((size_t)std::max<int64_t>(nblk, 1) * 16));
  DBB_TRY(L.pk_valid.reserve((size_t)std::max<int64_t>(nblk, 1) * 8));
  if (ptrs && (n_thr <= 0 || nbytes < (1 << 20))) {
    // no packer threads (or a batch too small to be worth waking them): the calling thread packs
    for (int64_t s = 0; s < ns; ++s)
      dbb::host_pack_sequence(ptrs[s], L.h_off[s + 1] - L.h_off[s], reinterpret_cast<uint32_t*>(L.pk_codes.p + 16 * L.h_blk[s]),
                              reinterpret_cast<uint64_t*>(L.pk_valid.p + 8 * L.h_blk[s]));
    if (nblk > 0) {
      DBB_CUDA_TRY(cudaMemcpyAsync(L.codes.p, L.pk_codes.p, (size_t)nblk * 16, cudaMemcpyHostToDevice, L.st));
      DBB_CUDA_TRY(cudaMemcpyAsync(L.valid.p, L.pk_valid.p, (size_t)nblk * 8, cudaMemcpyHostToDevice, L.st));
    }
    *n_ascii = 0;
    S.ascii_bytes -= nbytes;
    S.plane_bytes += nblk * 24;
    return DBB_OK;
  }
  // units
  L.unit.clear();
  L.unit.push_back(0);
  for (int64_t s = 0, acc0 = 0; s < ns; ++s) {
    if (L.h_off[s + 1] - acc0 >= UNIT_BYTES || s + 1 == ns) { L.unit.push_back(s + 1); acc0 = L.h_off[s + 1]; }
  }
  const int64_t U = (int64_t)L.unit.size() - 1;
  if ((size_t)U > L.unit_cap) { L.unit_cap = (size_t)U + 64; L.unit_done.reset(new std::atomic<uint8_t>[L.unit_cap]); }
  for (int64_t u = 0; u < U; ++u) L.unit_done[u].store(0, std::memory_order_relaxed);

  std::mutex mu;
  int64_t front = 0, back = U - 1;               // units [front, back] are unclaimed
  const int64_t* h_off = L.h_off.data();
  const int64_t* h_blk = L.h_blk.data();
  const int64_t* unit = L.unit.data();
  std::atomic<uint8_t>* done = L.unit_done.get();
  uint8_t* pkc = L.pk_codes.p;
  uint8_t* pkv = L.pk_valid.p;
  S.pool->start([&, h_off, h_blk, unit, done, pkc, pkv, base, ptrs](int) {
    for (;;) {
      int64_t u;
      { std::lock_guard<std::mutex> lk(mu); if (front > back) return; u = back--; }
      for (int64_t s = unit[u]; s < unit[u + 1]; ++s)
        dbb::host_pack_sequence(ptrs ? ptrs[s] : base + h_off[s], h_off[s + 1] - h_off[s], reinterpret_cast<uint32_t*>(pkc + 16 * h_blk[s]),
                                reinterpret_cast<uint64_t*>(pkv + 8 * h_blk[s]));
      done[u].store(1, std::memory_order_release);
    }
  });

  cudaError_t e = cudaSuccess;
  int64_t issued_hi = U;                          // packed units [issued_hi, U) are on their way
  int inflight = 0, ev_head = 0;                  // ASCII copies in flight: events ev_head .. ev_head + inflight - 1 (mod DEPTH)
  bool all_claimed = false;
  int64_t boundary = -1;                          // first packer-claimed unit, known once everything is claimed
  while (e == cudaSuccess) {
    bool progress = false;
    // packed route: the run of finished units just below issued_hi
    int64_t lo = issued_hi;
    while (lo > 0 && done[lo - 1].load(std::memory_order_acquire)) --lo;
    if (lo < issued_hi && (issued_hi - lo >= 4 || all_claimed)) {
      const int64_t b0 = h_blk[unit[lo]], b1 = h_blk[unit[issued_hi]];
      if (b1 > b0) {
        e = cudaMemcpyAsync((uint8_t*)L.codes.p + 16 * b0, pkc + 16 * b0, (size_t)(b1 - b0) * 16, cudaMemcpyHostToDevice, L.st2);
        if (e == cudaSuccess) e = cudaMemcpyAsync((uint8_t*)L.valid.p + 8 * b0, pkv + 8 * b0, (size_t)(b1 - b0) * 8, cudaMemcpyHostToDevice, L.st2);
      }
      issued_hi = lo;
      progress = true;
    }
    // ASCII route: retire finished copies, claim the next units while the engine has room
    while (inflight > 0 && cudaEventQuery(L.dma_ev[ev_head]) == cudaSuccess) { ev_head = (ev_head + 1) % DMA_DEPTH; --inflight; progress = true; }
    if (!all_claimed && !pack_all && inflight < DMA_DEPTH) {
      int64_t u0 = -1, u1 = -1;
      { std::lock_guard<std::mutex> lk(mu); if (front <= back) { u0 = front; u1 = std::min(front + 2, back + 1); front = u1; } }
      if (u0 >= 0) {
        const int64_t o0 = h_off[unit[u0]], o1 = h_off[unit[u1]];
        if (o1 > o0) e = cudaMemcpyAsync((uint8_t*)L.ascii.p + o0, base + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, L.st);
        const int slot = (ev_head + inflight) % DMA_DEPTH;
        if (e == cudaSuccess) e = cudaEventRecord(L.dma_ev[slot], L.st);
        ++inflight;
        progress = true;
      }
    }
    if (!all_claimed) {
      std::lock_guard<std::mutex> lk(mu);
      if (front > back) { all_claimed = true; boundary = front; }
    }
    if (all_claimed && issued_hi <= boundary) break;
    if (!progress) std::this_thread::yield();
  }
  if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(mu); front = back + 1; }      // stop the packers
  S.pool->wait();
  if (e != cudaSuccess) { dbb::set_error("hybrid upload failed: %s", cudaGetErrorString(e)); return DBB_ERR_CUDA; }
  *n_ascii = unit[boundary];
  S.ascii_bytes -= nbytes - h_off[*n_ascii];                 // (the caller charged the whole batch to the ASCII route)
  S.plane_bytes += (h_blk[ns] - h_blk[*n_ascii]) * 24;
  DBB_CUDA_TRY(cudaEventRecord(L.ev2, L.st2));
  DBB_CUDA_TRY(cudaStreamWaitEvent(L.st, L.ev2, 0));
  return DBB_OK;
}

// where the sequences of a call live: back to back in one buffer (ascii + seq_off) or one host pointer per scaffold
struct SeqSource {
  const uint8_t* ascii = nullptr;
  const int64_t* seq_off = nullptr;
  const uint8_t* const* ptrs = nullptr;
  const int64_t* lens = nullptr;
  int64_t len(int64_t s) const { return ptrs ? lens[s] : seq_off[s + 1] - seq_off[s]; }
};

int submit_batch(Lane& L, SigState& S, int device, const SeqSource& src, int64_t s0, int64_t s1, bool want_counts, bool want_freq) {
  const int64_t ns = s1 - s0;
  L.s0 = s0; L.s1 = s1;
  L.h_off.resize(ns + 1);
  L.h_blk.resize(ns + 1);
  int64_t acc = 0, bytes = 0;
  for (int64_t i = 0; i < ns; ++i) {
    const int64_t len = src.len(s0 + i);
    L.h_off[i] = bytes;
    L.h_blk[i] = acc;
    bytes += len;
    acc += dbb_packed_blocks(len);
  }
  L.h_off[ns] = bytes;
  L.h_blk[ns] = acc;
  const int64_t nblk = acc, nbytes = bytes;
  if (!src.ptrs) DBB_TRY(L.ascii.reserve((size_t)std::max<int64_t>(nbytes, 16)));
  DBB_TRY(L.seq_off.reserve((ns + 1) * 8));
  DBB_TRY(L.blk_off.reserve((ns + 1) * 8));
  DBB_TRY(L.codes.reserve((size_t)std::max<int64_t>(nblk, 1) * 16 + 16));
  DBB_TRY(L.valid.reserve((size_t)std::max<int64_t>(nblk, 1) * 8 + 16));
  if (want_counts) DBB_TRY(L.counts.reserve((size_t)ns * DBB_NUM_KMERS * 4));
  if (want_freq) DBB_TRY(L.freq.reserve((size_t)ns * DBB_NUM_KMERS * 4));
  if (!L.plan) DBB_TRY(dbb_count_plan_create(L.h_blk.data(), ns, &L.plan));
  else DBB_TRY(dbb::count_plan_rebuild(L.plan, L.h_blk.data(), ns, L.st));
  DBB_CUDA_TRY(cudaMemcpyAsync(L.seq_off.p, L.h_off.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, L.st));
  DBB_CUDA_TRY(cudaMemcpyAsync(L.blk_off.p, L.h_blk.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, L.st));
  int64_t n_ascii = ns;                 // scaffolds [0, n_ascii) arrive as ASCII, the others as host-packed planes
  S.ascii_bytes += nbytes;
  DBB_TRY(hybrid_upload(L, S, device, src.ptrs ? nullptr : src.ascii + src.seq_off[s0], src.ptrs ? src.ptrs + s0 : nullptr, ns, nbytes, &n_ascii));
  if (n_ascii > 0)
    DBB_TRY(dbb_pack_ascii_dev((const uint8_t*)L.ascii.p, (const int64_t*)L.seq_off.p, (const int64_t*)L.blk_off.p, n_ascii,
                               L.h_blk[n_ascii], L.codes.p, L.valid.p, L.st));
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

namespace {

// batches of ~batch_bytes of sequence, two in flight (upload of batch b+1 overlaps compute of batch b).  The lanes
// (streams, device buffers, count plans) belong to the default context of the current device and only grow: a
// per-sequence caller such as GenomicSignatures.seqSignature pays no allocation after its first call.
int signatures_host_impl(const SeqSource& src, int64_t n_seq, uint32_t* counts, float* freq) {
  dbb_ctx* ctx = nullptr;
  int rc = dbb::ctx_resolve(nullptr, &ctx);
  if (rc != DBB_OK) return rc;
  dbb::CtxLock lock(ctx);
  if (!ctx->sig) ctx->sig = new dbb::SigState();
  Lane* lanes = ctx->sig->lanes;
  ctx->sig->ascii_bytes = ctx->sig->plane_bytes = 0;
  // batch size: pinned input streams best in small batches (the last batch's count + copy-back is the only part the
  // link does not overlap: C2 53.5 / 52.5 / 51.6 / 48.9 Gbp/s for 32 / 128 / 256 / 512 MB), pageable input in large ones
  // (every batch ends with the packer threads meeting: 23.6 / 37.9 / 44.4 / 47.7 Gbp/s)
  const char* env = getenv("DBB_HOST_BATCH_MB");
  bool big_batches = src.ptrs != nullptr;
  if (!src.ptrs && src.ascii) {
    cudaPointerAttributes pattr;
    big_batches = cudaPointerGetAttributes(&pattr, src.ascii) != cudaSuccess || pattr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
  }
  const int64_t batch_bytes = (env && *env ? atoll(env) : (big_batches && pack_threads() > 0) ? 512 : 32) << 20;
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
    int64_t e = s, bytes = 0;
    while (e < n_seq && (e == s || bytes + src.len(e) <= batch_bytes)) { bytes += src.len(e); ++e; }
    Lane& L = lanes[cur];
    if (pending[cur]) { rc = collect_batch(L, counts, freq); pending[cur] = false; if (rc != DBB_OK) break; }
    rc = submit_batch(L, *ctx->sig, ctx->device, src, s, e, counts != nullptr, freq != nullptr);
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

}  // namespace

DBB_EXPORT int dbb_tetra_signatures_host(const uint8_t* ascii, const int64_t* seq_off, int64_t n_seq,
                                         uint32_t* counts, float* freq) {
  DBB_REQUIRE(n_seq >= 0, "negative n_seq");
  if (n_seq == 0) return DBB_OK;
  DBB_REQUIRE(seq_off != nullptr && (counts || freq), "NULL argument");
  DBB_REQUIRE(ascii != nullptr || seq_off[n_seq] == seq_off[0], "ascii is NULL");
  for (int64_t s = 0; s < n_seq; ++s) DBB_REQUIRE(seq_off[s + 1] >= seq_off[s], "seq_off must be non-decreasing");
  SeqSource src;
  src.ascii = ascii; src.seq_off = seq_off;
  return signatures_host_impl(src, n_seq, counts, freq);
}

DBB_EXPORT int dbb_tetra_signatures_host_ptrs(const uint8_t* const* seqs, const int64_t* seq_len, int64_t n_seq,
                                              uint32_t* counts, float* freq) {
  DBB_REQUIRE(n_seq >= 0, "negative n_seq");
  if (n_seq == 0) return DBB_OK;
  DBB_REQUIRE(seqs != nullptr && seq_len != nullptr && (counts || freq), "NULL argument");
  for (int64_t s = 0; s < n_seq; ++s) {
    DBB_REQUIRE(seq_len[s] >= 0, "negative sequence length");
    DBB_REQUIRE(seq_len[s] == 0 || seqs[s] != nullptr, "NULL sequence pointer");
  }
  SeqSource src;
  src.ptrs = seqs; src.lens = seq_len;
  return signatures_host_impl(src, n_seq, counts, freq);
}

DBB_EXPORT int dbb_host_upload_stats(int64_t* ascii_bytes, int64_t* plane_bytes, int32_t* pack_threads_out) {
  dbb_ctx* ctx = nullptr;
  int rc = dbb::ctx_resolve(nullptr, &ctx);
  if (rc != DBB_OK) return rc;
  dbb::CtxLock lock(ctx);
  if (ascii_bytes) *ascii_bytes = ctx->sig ? ctx->sig->ascii_bytes : 0;
  if (plane_bytes) *plane_bytes = ctx->sig ? ctx->sig->plane_bytes : 0;
  if (pack_threads_out) *pack_threads_out = pack_threads();
  return DBB_OK;
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
