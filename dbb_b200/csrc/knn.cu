// dbb_knn_dev / dbb_knn_host: the default stage 2 behind ONE C call (no torch, no Python orchestration).
//
// What the call replaces in the reference: TdNeighbours.run's N x N loop (tdNeighbours.py:41-51) with the GC / coverage
// predicates of gcNeighbours.py:43-52 / coverageNeighbours.py:41-50, as described at dbb_pairs_input, plus the new fused
// top-k.  The orchestration here is the exact tile-skipping sweep:
//   class quantiles of td_bound (radix sort + scatter) -> dbb_prune_plan_dev (projection, sort, gathers, tile lists)
//   -> sparse dbb_pairs_dev -> [top-k not TD-filtered: dbb_prune_second_dev -> second sparse sweep -> list merge]
//   -> rows back in the caller's order (single GPU) or this rank's rows compacted with their ids (multi-GPU).
// Everything is enqueued on the caller's stream; the host reads 64 bytes once (list sizes, rows that do not sum to 1).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <vector>

namespace dbb {

struct KnnState {
  DevBuf sig_sorted, length, td_bound, gc, gc_mbin, gc_lbin, cov, cov_lbin;     // the sorted order
  DevBuf col_id, p, order, off, tl, cnt, mine, info, scratch;
  DevBuf cls, key_in, key_out, val_in, val_out, cub, weights;
  DevBuf s_idx, s_dist, s_cnt, s_bp, s2_idx, s2_dist, tile_pos;
  // host entry point: device copies of the caller's arrays and outputs
  DevBuf h_dense, h_sig, h_len, h_tdb, h_gc, h_gcm, h_gcl, h_gclo, h_gchi, h_cov, h_covl, h_covlo, h_covhi;
  DevBuf o_ids, o_idx, o_dist, o_cnt, o_bp;
  int64_t* h_info = nullptr;            // pinned, 8 x int64
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool weights_ready = false;
};

void free_knn_state(KnnState* s) {
  if (!s) return;
  for (DevBuf* b : {&s->sig_sorted, &s->length, &s->td_bound, &s->gc, &s->gc_mbin, &s->gc_lbin, &s->cov, &s->cov_lbin, &s->col_id, &s->p,
                    &s->order, &s->off, &s->tl, &s->cnt, &s->mine, &s->info, &s->scratch, &s->cls, &s->key_in, &s->key_out, &s->val_in,
                    &s->val_out, &s->cub, &s->weights, &s->s_idx, &s->s_dist, &s->s_cnt, &s->s_bp, &s->s2_idx, &s->s2_dist, &s->tile_pos,
                    &s->h_dense, &s->h_sig, &s->h_len, &s->h_tdb, &s->h_gc, &s->h_gcm, &s->h_gcl, &s->h_gclo, &s->h_gchi, &s->h_cov,
                    &s->h_covl, &s->h_covlo, &s->h_covhi, &s->o_ids, &s->o_idx, &s->o_dist, &s->o_cnt, &s->o_bp})
    b->release();
  if (s->h_info) cudaFreeHost(s->h_info);
  for (auto& e : s->ev) if (e) cudaEventDestroy(e);
  delete s;
}

}  // namespace dbb

namespace {

using dbb::DevBuf;
using dbb::KnnState;
constexpr int T = 128;

#define DBB_TRY(expr) do { int _rc = (expr); if (_rc != DBB_OK) return _rc; } while (0)

// keys of the class sort: td_bound >= 0, so the float bit pattern is monotone (NaN / negative bounds sort last / first,
// any arrangement is exact)
__global__ void class_keys_kernel(const float* __restrict__ td_bound, int64_t n, uint32_t* __restrict__ key, int32_t* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = __float_as_uint(fmaxf(td_bound[i], 0.f)); val[i] = (int32_t)i; }
}
// equal-count classes of the per-scaffold TD bound: class of the scaffold at sorted position i = floor(i * nc / n)
__global__ void class_rank_kernel(const int32_t* __restrict__ perm, int64_t n, int nc, float* __restrict__ cls) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cls[perm[i]] = (float)((i * nc) / n);
}

// single GPU: results of sorted position r go to row col_id[r] of the caller's arrays (one warp per row)
__global__ void __launch_bounds__(256) unsort_kernel(const int32_t* __restrict__ col_id, int64_t n, int k,
                                                     const int32_t* __restrict__ s_idx, const float* __restrict__ s_dist,
                                                     const int32_t* __restrict__ s_cnt, const int64_t* __restrict__ s_bp,
                                                     int32_t* __restrict__ o_idx, float* __restrict__ o_dist,
                                                     int32_t* __restrict__ o_cnt, int64_t* __restrict__ o_bp, int64_t* __restrict__ o_ids) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t dst = col_id[r];
  if (lane < k) { o_idx[dst * k + lane] = s_idx[r * k + lane]; o_dist[dst * k + lane] = s_dist[r * k + lane]; }
  if (lane == 0) {
    if (o_cnt) o_cnt[dst] = s_cnt[r];
    if (o_bp) o_bp[dst] = s_bp[r];
    if (o_ids) o_ids[dst] = dst;
  }
}

// multi-GPU: tile_pos[t] = number of this rank's row tiles before tile t (ascending sorted position); info[5] = rows of
// this rank.  One CTA.
__global__ void __launch_bounds__(1024) mine_scan_kernel(const uint8_t* __restrict__ mine, int64_t n_rt, int64_t n,
                                                         int32_t* __restrict__ tile_pos, int64_t* __restrict__ info) {
  __shared__ int s_sum[1024];
  const int t = threadIdx.x;
  const int64_t per = (n_rt + 1023) / 1024, b = t * per, e = min(b + per, n_rt);
  int s = 0;
  for (int64_t i = b; i < e; ++i) s += mine[i] ? 1 : 0;
  s_sum[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    int v = 0;
    if (t >= d) v = s_sum[t - d];
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  int run = s_sum[t] - s;
  for (int64_t i = b; i < e; ++i) { tile_pos[i] = run; run += mine[i] ? 1 : 0; }
  if (t == 1023) {
    const int64_t tiles = s_sum[1023];
    const int64_t last_rows = n - (n_rt - 1) * T;                       // rows of the last (maybe partial) tile
    info[5] = tiles * T - ((n_rt > 0 && mine[n_rt - 1]) ? (T - last_rows) : 0);
  }
}

// multi-GPU: this rank's rows, compacted in ascending sorted position, with their original ids (one warp per row)
__global__ void __launch_bounds__(256) gather_mine_kernel(const uint8_t* __restrict__ mine, const int32_t* __restrict__ tile_pos,
                                                          const int32_t* __restrict__ col_id, int64_t n, int k,
                                                          const int32_t* __restrict__ s_idx, const float* __restrict__ s_dist,
                                                          const int32_t* __restrict__ s_cnt, const int64_t* __restrict__ s_bp,
                                                          int32_t* __restrict__ o_idx, float* __restrict__ o_dist,
                                                          int32_t* __restrict__ o_cnt, int64_t* __restrict__ o_bp,
                                                          int64_t* __restrict__ o_ids, int64_t cap_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t rt = r / T;
  if (!mine[rt]) return;
  const int64_t dst = (int64_t)tile_pos[rt] * T + (r - rt * T);
  if (dst >= cap_rows) return;
  if (lane < k) { o_idx[dst * k + lane] = s_idx[r * k + lane]; o_dist[dst * k + lane] = s_dist[r * k + lane]; }
  if (lane == 0) {
    if (o_cnt) o_cnt[dst] = s_cnt[r];
    if (o_bp) o_bp[dst] = s_bp[r];
    if (o_ids) o_ids[dst] = col_id[r];
  }
}

__global__ void iota_kernel(int64_t* __restrict__ ids, int64_t r0, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) ids[i] = r0 + i;
}

// projection weights: per canonical tetramer 0 / 0.5 / 1 for <= 1 / 2 / >= 3 G or C, zero for the padding columns.  Any
// weights in [0, 1] give the bound; these spread metagenomes by their GC content, and the saturated form keeps more of
// the distance than the plain GC fraction (on the synthetic C2 / C4 inputs the pair-level bound leaves 29 % / 33 % of
// the pairs instead of 40 % / 44 %).
int upload_weights(KnnState& K, cudaStream_t st) {
  if (K.weights_ready) return DBB_OK;
  float w[DBB_SIG_STRIDE];
  const dbb::KmerTable& t = dbb::kmer_table();
  for (int c = 0; c < DBB_SIG_STRIDE; ++c) {
    w[c] = 0.f;
    if (c < DBB_NUM_KMERS) {
      int gc = 0;
      for (int p = 0; p < 4; ++p) gc += (t.names[c][p] == 'G' || t.names[c][p] == 'C') ? 1 : 0;
      w[c] = std::min(1.f, std::max(0.f, (gc - 1) / 2.f));
    }
  }
  DBB_TRY(K.weights.reserve(sizeof(w)));
  DBB_CUDA_TRY(cudaMemcpyAsync(K.weights.p, w, sizeof(w), cudaMemcpyHostToDevice, st));
  DBB_CUDA_TRY(cudaStreamSynchronize(st));                    // `w` is a stack array
  K.weights_ready = true;
  return DBB_OK;
}

int split_for(int64_t n_act, int64_t target) {
  // ~20 work items per SM (at most 8 segments per row tile): the lists differ in length by an order of magnitude, whole
  // row tiles do not balance.  Every segment fills its own top-k lists, but the segments of a row tile share the k-th best
  // distance they have proven so far (Params::gthr in pairs.cu), so a later segment starts with a tight threshold
  return (int)std::min<int64_t>(8, std::max<int64_t>(1, (target + std::max<int64_t>(n_act, 1) - 1) / std::max<int64_t>(n_act, 1)));
}

int full_sweep(const dbb_pairs_input* in, const dbb_knn_params* par, dbb_knn_output* out, cudaStream_t st) {
  const int64_t n = in->n;
  const int world = std::max(par->world, 1), rank = par->world > 1 ? par->rank : 0;
  const int64_t r0 = (n * rank) / world, r1 = (n * (rank + 1)) / world;
  DBB_REQUIRE(out->cap_rows >= r1 - r0, "cap_rows too small");
  dbb_pairs_input d = *in;
  d.col_id = nullptr; d.item_order = nullptr; d.tile_off = nullptr; d.tile_list = nullptr; d.split_all = 0; d.n_active_row_tiles = 0; d.split_tiles = 0;
  dbb_pairs_output o = {};
  o.k = par->k; o.topk_filter = par->topk_filter;
  o.knn_idx = out->knn_idx; o.knn_dist = out->knn_dist; o.count = out->count; o.neighbour_bp = out->neighbour_bp;
  if (r1 > r0) DBB_TRY(dbb_pairs_dev(&d, r0, r1, &o, st));
  if (out->row_ids && r1 > r0) {
    iota_kernel<<<(unsigned)((r1 - r0 + 255) / 256), 256, 0, st>>>(out->row_ids, r0, r1 - r0);
    DBB_CUDA_TRY(cudaGetLastError());
  }
  out->n_rows = r1 - r0;
  out->tiles_total = ((r1 - r0 + T - 1) / T) * ((n + T - 1) / T);
  out->tiles_visited = out->tiles_total;
  out->tiles_second = out->row_tiles_second = 0;
  out->pruned = 0;
  out->launches = r1 > r0 ? dbb_pairs_launches(n, r1 - r0) + (out->row_ids ? 1 : 0) : 0;
  out->ms_plan = out->ms_sweep = out->ms_second = 0.f;
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_knn_dev(dbb_ctx* ctx_in, const dbb_pairs_input* in, const dbb_knn_params* par, dbb_knn_output* out, void* stream) {
  DBB_REQUIRE(in && par && out, "NULL argument struct");
  DBB_REQUIRE(in->n >= 0 && in->n < (1ll << 31) - T, "n out of range");
  DBB_REQUIRE(par->k >= 1 && par->k <= DBB_MAX_K, "k out of range (1..32)");
  DBB_REQUIRE(par->world <= 1 || (par->rank >= 0 && par->rank < par->world), "bad rank / world");
  DBB_REQUIRE(out->knn_idx && out->knn_dist, "knn_idx / knn_dist are required");
  DBB_REQUIRE(par->world <= 1 || out->row_ids, "row_ids is required with world > 1");
  out->n_rows = 0; out->tiles_visited = out->tiles_total = 0; out->tiles_second = out->row_tiles_second = 0; out->pruned = 0; out->launches = 0;
  out->ms_plan = out->ms_sweep = out->ms_second = 0.f;
  if (in->n == 0) return DBB_OK;
  DBB_REQUIRE(in->sig && in->length, "sig/length are required");
  dbb_ctx* ctx = nullptr;
  DBB_TRY(dbb::ctx_resolve(ctx_in, &ctx));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = in->n;
  const int world = std::max(par->world, 1), rank = par->world > 1 ? par->rank : 0;
  const int k = par->k;
  // without a TD bound every pair counts (nothing can be skipped for count / neighbour_bp); a plain k-NN still prunes
  const bool want_count = out->count || out->neighbour_bp;
  const bool can_prune = (in->td_bound != nullptr || !want_count) && in->metric == DBB_METRIC_L1 && n > 2 * T;
  DBB_REQUIRE(par->prune != 1 || can_prune,
              "prune = 1 needs the Manhattan metric, more than two tiles of scaffolds and (for count / neighbour_bp) td_bound");
  if (par->prune == 0 || !can_prune) return full_sweep(in, par, out, st);

  if (!ctx->knn) ctx->knn = new KnnState();
  KnnState& K = *ctx->knn;
  const int64_t n_rt = (n + T - 1) / T, n_ct = n_rt;
  const int64_t n_mine = world > 1 ? (n_rt - rank + world - 1) / world : n_rt;
  const int64_t n_slots = world > 1 ? (n_rt + world - 1) / world : n_rt;
  DBB_REQUIRE(out->cap_rows >= (world > 1 ? std::min<int64_t>(n_mine * T, n) : n), "cap_rows too small");
  const int n_classes = !in->td_bound ? 1 : par->n_classes > 0 ? par->n_classes : 8;
  const float margin = par->margin > 0.f ? par->margin : 1e-5f;
  int dev = 0, sms = 148;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  DBB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static const int64_t items_target_env = [] { const char* e = getenv("DBB_PRUNE_ITEMS"); return (e && *e) ? atoll(e) : 0ll; }();
  const int64_t items_target = items_target_env > 0 ? items_target_env : 20ll * sms;

  // ---- buffers (grow-only, kept in the context between calls)
  if (!K.h_info) DBB_CUDA_TRY(cudaMallocHost((void**)&K.h_info, 8 * sizeof(int64_t)));
  for (auto& e : K.ev) if (!e) DBB_CUDA_TRY(cudaEventCreate(&e));
  const size_t nn = (size_t)n;
  DBB_TRY(K.sig_sorted.reserve(nn * DBB_SIG_STRIDE * 4));
  DBB_TRY(K.length.reserve(nn * 4)); if (in->td_bound) DBB_TRY(K.td_bound.reserve(nn * 4));
  if (in->gc) { DBB_TRY(K.gc.reserve(nn * 8)); DBB_TRY(K.gc_mbin.reserve(nn * 4)); DBB_TRY(K.gc_lbin.reserve(nn * 4)); }
  if (in->cov) { DBB_TRY(K.cov.reserve(nn * 8)); DBB_TRY(K.cov_lbin.reserve(nn * 4)); }
  DBB_TRY(K.col_id.reserve(nn * 4)); DBB_TRY(K.p.reserve(nn * 4));
  DBB_TRY(K.order.reserve((size_t)n_rt * 4)); DBB_TRY(K.off.reserve((size_t)(n_rt + 1) * 8));
  DBB_TRY(K.tl.reserve((size_t)std::max<int64_t>(n_slots * n_ct, 1) * 4));
  DBB_TRY(K.cnt.reserve((size_t)n_rt * 4)); DBB_TRY(K.mine.reserve((size_t)n_rt)); DBB_TRY(K.info.reserve(8 * 8));
  const int64_t scratch_bytes = dbb_prune_scratch_bytes(n, n, world);
  if (scratch_bytes < 0) return DBB_ERR_UNSUPPORTED;
  DBB_TRY(K.scratch.reserve((size_t)scratch_bytes));
  DBB_TRY(K.cls.reserve(nn * 4)); DBB_TRY(K.key_in.reserve(nn * 4)); DBB_TRY(K.key_out.reserve(nn * 4));
  DBB_TRY(K.val_in.reserve(nn * 4)); DBB_TRY(K.val_out.reserve(nn * 4));
  size_t cub_bytes = 0;
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                               (int32_t*)nullptr, (int)n));
  DBB_TRY(K.cub.reserve(cub_bytes));
  DBB_TRY(K.s_idx.reserve(nn * k * 4)); DBB_TRY(K.s_dist.reserve(nn * k * 4));
  DBB_TRY(K.s_cnt.reserve(nn * 4)); DBB_TRY(K.s_bp.reserve(nn * 8));
  DBB_TRY(upload_weights(K, st));
  DBB_CUDA_TRY(cudaMemsetAsync(K.info.p, 0, 8 * 8, st));

  DBB_CUDA_TRY(cudaEventRecord(K.ev[0], st));
  // ---- sort classes: equal-count quantiles of td_bound
  const float* cls = nullptr;
  int launches = 0;
  if (n_classes > 1) {
    const unsigned g = (unsigned)((n + 255) / 256);
    class_keys_kernel<<<g, 256, 0, st>>>(in->td_bound, n, (uint32_t*)K.key_in.p, (int32_t*)K.val_in.p);
    DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(K.cub.p, cub_bytes, (const uint32_t*)K.key_in.p, (uint32_t*)K.key_out.p,
                                                 (const int32_t*)K.val_in.p, (int32_t*)K.val_out.p, (int)n, 0, 32, st));
    class_rank_kernel<<<g, 256, 0, st>>>((const int32_t*)K.val_out.p, n, n_classes, (float*)K.cls.p);
    DBB_CUDA_TRY(cudaGetLastError());
    cls = (const float*)K.cls.p;
    launches += 4;
  }
  // ---- first plan
  dbb_prune_input pin = {};
  pin.n = n; pin.sig = in->sig; pin.weights = (const float*)K.weights.p; pin.cls = cls;
  pin.length = in->length; pin.td_bound = in->td_bound; pin.gc = in->gc; pin.gc_mbin = in->gc_mbin; pin.gc_lbin = in->gc_lbin;
  pin.cov = in->cov; pin.cov_lbin = in->cov_lbin;
  pin.row0 = 0; pin.row1 = n; pin.rank = rank; pin.world = world; pin.margin = margin;
  dbb_prune_buffers pb = {};
  pb.sig_sorted = (float*)K.sig_sorted.p; pb.length = (int32_t*)K.length.p; pb.td_bound = in->td_bound ? (float*)K.td_bound.p : nullptr;
  pb.gc = in->gc ? (double*)K.gc.p : nullptr; pb.gc_mbin = in->gc ? (int32_t*)K.gc_mbin.p : nullptr; pb.gc_lbin = in->gc ? (int32_t*)K.gc_lbin.p : nullptr;
  pb.cov = in->cov ? (double*)K.cov.p : nullptr; pb.cov_lbin = in->cov ? (int32_t*)K.cov_lbin.p : nullptr;
  pb.col_id = (int32_t*)K.col_id.p; pb.p_sorted = (float*)K.p.p; pb.item_order = (int32_t*)K.order.p; pb.tile_off = (int64_t*)K.off.p;
  pb.tile_list = (int32_t*)K.tl.p; pb.tile_count = (int32_t*)K.cnt.p; pb.mine = (uint8_t*)K.mine.p; pb.info = (int64_t*)K.info.p;
  pb.scratch = K.scratch.p; pb.scratch_bytes = scratch_bytes;
  DBB_TRY(dbb_prune_plan_dev(&pin, &pb, st));
  launches += world > 1 ? 12 : 11;
  DBB_CUDA_TRY(cudaEventRecord(K.ev[1], st));

  // ---- first sweep (sorted order; a rank's rows are the rows of its own tiles)
  dbb_pairs_input d = *in;
  d.sig = pb.sig_sorted; d.length = pb.length; d.td_bound = pb.td_bound; d.gc = pb.gc; d.gc_mbin = pb.gc_mbin; d.gc_lbin = pb.gc_lbin;
  d.cov = pb.cov; d.cov_lbin = pb.cov_lbin;
  d.col_id = pb.col_id; d.item_order = pb.item_order; d.tile_off = pb.tile_off; d.tile_list = pb.tile_list;
  // Work items: every row tile in split_for(...) strided segments.  The sweep ends with its longest item; since the
  // segments of a row tile share their proven k-th best distances, finer cuts are nearly free (C2, 391 row tiles: 2 / 4 /
  // 8 segments -> 11.0 / 10.8 / 10.7 ms, CTA finish times max / mean 1.05 / 1.03 / 1.01; without the shared bound 2
  // segments took 12.2 ms and finer cuts were slower, profiles/r02_k2_notes.md).  DBB_PRUNE_SEG_TILES=<tiles per segment>
  // selects the length-adaptive cut (dbb_pairs_input.split_tiles) instead.
  static const int64_t seg_env = [] { const char* e = getenv("DBB_PRUNE_SEG_TILES"); return (e && *e) ? atoll(e) : 0ll; }();
  if (seg_env <= 0) { d.split_all = split_for(n_mine, items_target); d.split_tiles = 0; }
  else { d.split_all = 8; d.split_tiles = (int32_t)std::min<int64_t>(seg_env, 1 << 30); }
  d.n_active_row_tiles = world > 1 ? (int32_t)n_mine : 0;
  dbb_pairs_output o = {};
  o.k = k; o.topk_filter = par->topk_filter;
  o.knn_idx = (int32_t*)K.s_idx.p; o.knn_dist = (float*)K.s_dist.p; o.count = (int32_t*)K.s_cnt.p; o.neighbour_bp = (int64_t*)K.s_bp.p;
  if (!want_count) { o.count = nullptr; o.neighbour_bp = nullptr; }
  if (n_mine > 0) { DBB_TRY(dbb_pairs_dev(&d, 0, n, &o, st)); launches += 2; }
  DBB_CUDA_TRY(cudaEventRecord(K.ev[2], st));

  // ---- second plan (top-k candidates beyond the TD window), then the one host read
  const bool second = !(par->topk_filter & 1) && n_mine > 0;
  if (second) {
    DBB_TRY(dbb_prune_second_dev(&pin, &pb, (const float*)K.s_dist.p + (k - 1), k, st));
    launches += 8;
  }
  if (world > 1) {
    DBB_TRY(K.tile_pos.reserve((size_t)n_rt * 4));
    mine_scan_kernel<<<1, 1024, 0, st>>>((const uint8_t*)K.mine.p, n_rt, n, (int32_t*)K.tile_pos.p, (int64_t*)K.info.p);
    DBB_CUDA_TRY(cudaGetLastError());
    launches += 1;
  }
  DBB_CUDA_TRY(cudaMemcpyAsync(K.h_info, K.info.p, 8 * 8, cudaMemcpyDeviceToHost, st));
  DBB_CUDA_TRY(cudaStreamSynchronize(st));
  if (K.h_info[4] != 0) {
    if (par->prune == 1) {
      dbb::set_error("invalid argument: %lld signature rows do not sum to 1 within %g: the tile-skipping bound L1 >= 2|p_a - p_b| does "
                     "not hold for them (pass frequencies, or prune = 0 / -1 for the full sweep)", (long long)K.h_info[4], 0.3 * margin);
      return DBB_ERR_INVALID;
    }
    return full_sweep(in, par, out, st);                      // auto: the exact full sweep on the GPU
  }
  int64_t visited = K.h_info[0];
  if (second) {
    const int64_t n_aff = K.h_info[3];
    visited += K.h_info[2];
    out->tiles_second = K.h_info[2]; out->row_tiles_second = n_aff;
    if (n_aff > 0) {
      DBB_TRY(K.s2_idx.reserve(nn * k * 4)); DBB_TRY(K.s2_dist.reserve(nn * k * 4));
      dbb_pairs_input d2 = d;
      // few row tiles, and a row whose filtered list is not full visits EVERY remaining column tile: the sweep takes as
      // long as its longest item, so lists are cut into pieces of <= ~32 tiles (C3: 11.3 -> 2 ms, at any GPU count)
      d2.split_all = (int)std::min<int64_t>(64, std::max<int64_t>(split_for(n_aff, 2 * items_target), (n_ct + 31) / 32));
      d2.n_active_row_tiles = (int32_t)n_aff;
      d2.split_tiles = 0;
      dbb_pairs_output o2 = {};
      o2.k = k; o2.topk_filter = par->topk_filter;
      o2.knn_idx = (int32_t*)K.s2_idx.p; o2.knn_dist = (float*)K.s2_dist.p;
      DBB_TRY(dbb_pairs_dev(&d2, 0, n, &o2, st));
      DBB_TRY(dbb_prune_merge_dev(&pb, n, k, (int32_t*)K.s_idx.p, (float*)K.s_dist.p, (const int32_t*)K.s2_idx.p, (const float*)K.s2_dist.p, st));
      launches += 3;
    }
  }
  DBB_CUDA_TRY(cudaEventRecord(K.ev[3], st));

  // ---- rows back to the caller
  const unsigned gw = (unsigned)((n + 7) / 8);
  if (world <= 1) {
    unsort_kernel<<<gw, 256, 0, st>>>((const int32_t*)K.col_id.p, n, k, (const int32_t*)K.s_idx.p, (const float*)K.s_dist.p,
                                      (const int32_t*)K.s_cnt.p, (const int64_t*)K.s_bp.p, out->knn_idx, out->knn_dist,
                                      want_count ? out->count : nullptr, want_count ? out->neighbour_bp : nullptr, out->row_ids);
    out->n_rows = n;
  } else {
    gather_mine_kernel<<<gw, 256, 0, st>>>((const uint8_t*)K.mine.p, (const int32_t*)K.tile_pos.p, (const int32_t*)K.col_id.p, n, k,
                                           (const int32_t*)K.s_idx.p, (const float*)K.s_dist.p, (const int32_t*)K.s_cnt.p,
                                           (const int64_t*)K.s_bp.p, out->knn_idx, out->knn_dist, want_count ? out->count : nullptr,
                                           want_count ? out->neighbour_bp : nullptr, out->row_ids, out->cap_rows);
    out->n_rows = K.h_info[5];
  }
  DBB_CUDA_TRY(cudaGetLastError());
  launches += 1;
  out->tiles_visited = visited;
  out->tiles_total = n_mine * n_ct;
  out->pruned = 1;
  out->launches = launches;
  // phase times of THIS call (the events up to ev[2] completed before the host read; ev[3] may still be in flight)
  cudaEventElapsedTime(&out->ms_plan, K.ev[0], K.ev[1]);
  cudaEventElapsedTime(&out->ms_sweep, K.ev[1], K.ev[2]);
  if (cudaEventSynchronize(K.ev[3]) == cudaSuccess) cudaEventElapsedTime(&out->ms_second, K.ev[2], K.ev[3]);
  return DBB_OK;
}

namespace {

template <typename Tp>
int to_dev(DevBuf& b, const Tp* h, size_t count, cudaStream_t st) {
  DBB_TRY(b.reserve(std::max<size_t>(count * sizeof(Tp), 16)));
  if (count) DBB_CUDA_TRY(cudaMemcpyAsync(b.p, h, count * sizeof(Tp), cudaMemcpyHostToDevice, st));
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_knn_host(dbb_ctx* ctx_in, const dbb_pairs_input* in, const dbb_knn_params* par, dbb_knn_output* out) {
  DBB_REQUIRE(in && par && out, "NULL argument struct");
  DBB_REQUIRE(in->n >= 0, "negative n");
  DBB_REQUIRE(par->k >= 1 && par->k <= DBB_MAX_K, "k out of range (1..32)");
  DBB_REQUIRE(out->knn_idx && out->knn_dist, "knn_idx / knn_dist are required");
  out->n_rows = 0;
  if (in->n == 0) return DBB_OK;
  DBB_REQUIRE(in->sig && in->length, "sig/length are required");
  dbb_ctx* ctx = nullptr;
  DBB_TRY(dbb::ctx_resolve(ctx_in, &ctx));
  dbb::CtxLock lock(ctx);
  if (!ctx->knn) ctx->knn = new KnnState();
  KnnState& K = *ctx->knn;
  cudaStream_t st = ctx->st;
  const size_t n = (size_t)in->n;
  const size_t k = (size_t)par->k;
  const size_t cap = (size_t)out->cap_rows;
  dbb_pairs_input d = *in;
  d.col_id = nullptr; d.item_order = nullptr; d.tile_off = nullptr; d.tile_list = nullptr; d.split_all = 0; d.n_active_row_tiles = 0; d.split_tiles = 0;
  // on the host side `sig` is dense [n][136]
  DBB_TRY(to_dev(K.h_dense, in->sig, n * DBB_NUM_KMERS, st));
  DBB_TRY(K.h_sig.reserve(n * DBB_SIG_STRIDE * 4));
  DBB_TRY(dbb_pad_signatures_dev((const float*)K.h_dense.p, in->n, (float*)K.h_sig.p, st));
  d.sig = (const float*)K.h_sig.p;
  DBB_TRY(to_dev(K.h_len, in->length, n, st)); d.length = (const int32_t*)K.h_len.p;
  if (in->td_bound) { DBB_TRY(to_dev(K.h_tdb, in->td_bound, n, st)); d.td_bound = (const float*)K.h_tdb.p; }
  if (in->gc) {
    DBB_REQUIRE(in->gc_mbin && in->gc_lbin && in->gc_lo && in->gc_hi && in->gc_nm > 0 && in->gc_nl > 0, "incomplete GC inputs");
    DBB_TRY(to_dev(K.h_gc, in->gc, n, st)); DBB_TRY(to_dev(K.h_gcm, in->gc_mbin, n, st)); DBB_TRY(to_dev(K.h_gcl, in->gc_lbin, n, st));
    DBB_TRY(to_dev(K.h_gclo, in->gc_lo, (size_t)in->gc_nm * in->gc_nl, st)); DBB_TRY(to_dev(K.h_gchi, in->gc_hi, (size_t)in->gc_nm * in->gc_nl, st));
    d.gc = (const double*)K.h_gc.p; d.gc_mbin = (const int32_t*)K.h_gcm.p; d.gc_lbin = (const int32_t*)K.h_gcl.p;
    d.gc_lo = (const double*)K.h_gclo.p; d.gc_hi = (const double*)K.h_gchi.p;
  }
  if (in->cov) {
    DBB_REQUIRE(in->cov_lbin && in->cov_lo && in->cov_hi && in->cov_nl > 0, "incomplete coverage inputs");
    DBB_TRY(to_dev(K.h_cov, in->cov, n, st)); DBB_TRY(to_dev(K.h_covl, in->cov_lbin, n, st));
    DBB_TRY(to_dev(K.h_covlo, in->cov_lo, (size_t)in->cov_nl, st)); DBB_TRY(to_dev(K.h_covhi, in->cov_hi, (size_t)in->cov_nl, st));
    d.cov = (const double*)K.h_cov.p; d.cov_lbin = (const int32_t*)K.h_covl.p;
    d.cov_lo = (const double*)K.h_covlo.p; d.cov_hi = (const double*)K.h_covhi.p;
  }
  dbb_knn_output o = *out;
  DBB_TRY(K.o_idx.reserve(std::max<size_t>(cap * k * 4, 16))); DBB_TRY(K.o_dist.reserve(std::max<size_t>(cap * k * 4, 16)));
  o.knn_idx = (int32_t*)K.o_idx.p; o.knn_dist = (float*)K.o_dist.p;
  o.row_ids = nullptr; o.count = nullptr; o.neighbour_bp = nullptr;
  if (out->row_ids || par->world > 1) { DBB_TRY(K.o_ids.reserve(std::max<size_t>(cap * 8, 16))); o.row_ids = (int64_t*)K.o_ids.p; }
  if (out->count) { DBB_TRY(K.o_cnt.reserve(std::max<size_t>(cap * 4, 16))); o.count = (int32_t*)K.o_cnt.p; }
  if (out->neighbour_bp) { DBB_TRY(K.o_bp.reserve(std::max<size_t>(cap * 8, 16))); o.neighbour_bp = (int64_t*)K.o_bp.p; }
  DBB_REQUIRE(par->world <= 1 || out->row_ids, "row_ids is required with world > 1");
  DBB_TRY(dbb_knn_dev(ctx, &d, par, &o, st));
  const size_t rows = (size_t)o.n_rows;
  cudaError_t e = cudaSuccess;
  auto back = [&](void* h, const DevBuf& b, size_t bytes) {
    if (h && bytes && e == cudaSuccess) e = cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st);
  };
  back(out->knn_idx, K.o_idx, rows * k * 4);
  back(out->knn_dist, K.o_dist, rows * k * 4);
  back(out->row_ids, K.o_ids, rows * 8);
  back(out->count, K.o_cnt, rows * 4);
  back(out->neighbour_bp, K.o_bp, rows * 8);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { dbb::set_error("knn_host failed: %s", cudaGetErrorString(e)); return DBB_ERR_CUDA; }
  out->n_rows = o.n_rows; out->tiles_visited = o.tiles_visited; out->tiles_total = o.tiles_total; out->pruned = o.pruned;
  out->tiles_second = o.tiles_second; out->row_tiles_second = o.row_tiles_second;
  out->launches = o.launches; out->ms_plan = o.ms_plan; out->ms_sweep = o.ms_sweep; out->ms_second = o.ms_second;
  return DBB_OK;
}
