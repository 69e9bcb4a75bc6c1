// Planner of the exact pruned sweep of stage 2 (dbb_b200/device.py::pairs_pruned): everything between the signature
// matrix and the tile lists dbb_pairs_dev consumes, on the device, with no host round trip.
//
// Bound (include/dbb_cuda.h, dbb_pairs_input): signatures sum to 1, so for weights w_k in [0, 1] and
// p = sum_k w_k sig_k,  L1(a, b) >= 2 |p_a - p_b|.  Scaffolds are sorted by (TD-bound class, p); every 128-row tile
// then has a p interval and a largest TD bound, and a (row tile, column tile) pair whose intervals are farther apart
// than half the larger of the two bounds cannot hold a TD neighbour (tdNeighbours.py:43-51: the bound of a pair is
// the bound of its shorter scaffold, i.e. one of the two).  For an unfiltered top-k a second plan takes the tiles
// whose lower bound does not exceed the k-th best distance the first sweep left in some row of the row tile.
//
// Steps (one stream, ~12 small launches): projection + sort key -> radix sort (CUB, plumbing) -> gather of the
// signature rows and the per-scaffold tables into the sorted order + per-tile statistics -> per row tile the number
// of tiles to visit -> (multi-GPU: row tiles dealt to the ranks by decreasing work) -> 64-bit keys (slot | quantised
// lower bound | column tile) -> radix sort -> offsets -> tile lists (nearest tiles first: the top-k thresholds tighten
// at once) -> row tiles by decreasing work.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>

namespace {

constexpr int T = 128;                 // tile edge of pairs_kernel
constexpr int QBITS = 20;              // bits of the quantised lower bound inside a sort key
constexpr uint32_t QNONE = (1u << QBITS) - 1u;   // "not visited" sorts behind every visited tile of its row tile

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int bits_for(int64_t v) { int b = 1; while ((1ll << b) < v) ++b; return b; }

// byte offsets into the caller's scratch; the first part (up to `persistent`) must survive from dbb_prune_plan_dev to
// dbb_prune_second_dev
struct Layout {
  size_t pmin_c, pmax_c, tmax_c, pmin_r, pmax_r, tmax_r, thr_r, cnt_all, slot, persistent;
  size_t p, key_in, key_out, val_in, val_out, okey_in, okey_out, oval_in, oval_out, tkey_in, tkey_out, cub_temp, total;
  size_t cub_bytes;
  int64_t n_ct, n_rt, n_slots;
};

int make_layout(int64_t n, int64_t rows, int32_t world, Layout* L) {
  L->n_ct = (n + T - 1) / T;
  L->n_rt = (rows + T - 1) / T;
  L->n_slots = world > 1 ? (L->n_rt + world - 1) / world : L->n_rt;
  const size_t nct = (size_t)std::max<int64_t>(L->n_ct, 1), nrt = (size_t)std::max<int64_t>(L->n_rt, 1), nn = (size_t)std::max<int64_t>(n, 1);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes); return r; };
  L->pmin_c = take(nct * 4); L->pmax_c = take(nct * 4); L->tmax_c = take(nct * 4);
  L->pmin_r = take(nrt * 4); L->pmax_r = take(nrt * 4); L->tmax_r = take(nrt * 4); L->thr_r = take(nrt * 4);
  L->cnt_all = take(nrt * 4); L->slot = take(nrt * 4);
  L->persistent = o;
  L->p = take(nn * 4); L->key_in = take(nn * 4); L->key_out = take(nn * 4); L->val_in = take(nn * 4); L->val_out = take(nn * 4);
  L->okey_in = take(nrt * 4); L->okey_out = take(nrt * 4); L->oval_in = take(nrt * 4); L->oval_out = take(nrt * 4);
  const size_t n_keys = (size_t)std::max<int64_t>(L->n_slots, 1) * nct;
  if (n_keys >= (1ull << 31)) { dbb::set_error("pruned sweep: %zu tile pairs per rank exceed one radix sort", n_keys); return DBB_ERR_UNSUPPORTED; }
  L->tkey_in = take(n_keys * 8); L->tkey_out = take(n_keys * 8);
  size_t b1 = 0, b2 = 0, b3 = 0;
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)nn));
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)nrt));
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, b3, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n_keys));
  L->cub_bytes = std::max(b1, std::max(b2, b3));
  L->cub_temp = take(L->cub_bytes);
  L->total = o;
  return DBB_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// p = <sig, w> per scaffold (one warp per row, 35 x 128-bit loads) and its sort key (class * 4 + clamp(p, 0, 1): p is in
// [0, 1], so classes do not interleave; a NaN signature sorts as p = 0 and is left out of the tile intervals).
// The bound L1(a, b) >= 2 |p_a - p_b| needs sum(a) = sum(b): in general L1 >= 2 |p_a - p_b| - |sum(a) - sum(b)|.  Both
// sums are accumulated in FP64 here; a finite row whose sum is farther than sum_tol from 1 is counted in *bad_rows and
// the caller refuses the plan (margin covers 2 * sum_tol plus the FP32 rounding of p).
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ sig, const float* __restrict__ w,
                                                      const float* __restrict__ cls, int64_t n, float sum_tol, float* __restrict__ p,
                                                      uint32_t* __restrict__ key, int32_t* __restrict__ val,
                                                      unsigned long long* __restrict__ bad_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float4* s = reinterpret_cast<const float4*>(sig + row * DBB_SIG_STRIDE);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  double acc = 0.0, tot = 0.0;
  for (int q = lane; q < DBB_SIG_STRIDE / 4; q += 32) {
    const float4 v = __ldg(s + q), x = __ldg(w4 + q);
    acc += ((double)v.x * x.x + (double)v.y * x.y) + ((double)v.z * x.z + (double)v.w * x.w);
    tot += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
    tot += __shfl_xor_sync(0xffffffffu, tot, d);
  }
  if (lane == 0) {
    const float pf = (float)acc;
    p[row] = pf;
    const float c = fminf(fmaxf(pf, 0.f), 1.f);                        // fmaxf(NaN, 0) = 0
    key[row] = __float_as_uint((cls ? cls[row] * 4.f : 0.f) + c);      // >= +0: the bit pattern is monotone
    val[row] = (int32_t)row;
    if (tot == tot && !(fabs(tot - 1.0) <= (double)sum_tol)) atomicAdd(bad_rows, 1ull);
  }
}

struct Tables {
  const int32_t* length; const float* td_bound; const double* gc; const int32_t* gc_mbin; const int32_t* gc_lbin;
  const double* cov; const int32_t* cov_lbin;
};
struct TablesOut {
  int32_t* length; float* td_bound; double* gc; int32_t* gc_mbin; int32_t* gc_lbin; double* cov; int32_t* cov_lbin;
};

// one CTA per column tile of the sorted order: rows and tables gathered, p interval and largest TD bound of the tile
__global__ void __launch_bounds__(T) gather_kernel(const float* __restrict__ sig, const float* __restrict__ p,
                                                   const int32_t* __restrict__ perm, int64_t n, Tables in, TablesOut out,
                                                   float* __restrict__ sig_s, float* __restrict__ p_s, int32_t* __restrict__ col_id,
                                                   float* __restrict__ pmin_c, float* __restrict__ pmax_c, float* __restrict__ tmax_c) {
  __shared__ float s_red[3][T / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t tile = blockIdx.x, r = tile * T + t;
  float pmn = INFINITY, pmx = -INFINITY, tmx = -INFINITY;
  if (r < n) {
    const int32_t src = perm[r];
    const float pv = p[src];
    p_s[r] = pv;
    col_id[r] = src;
    if (pv == pv) { pmn = pv; pmx = pv; }
    if (in.length) out.length[r] = in.length[src];
    if (in.td_bound) { const float b = in.td_bound[src]; out.td_bound[r] = b; tmx = b; }
    if (in.gc) out.gc[r] = in.gc[src];
    if (in.gc_mbin) out.gc_mbin[r] = in.gc_mbin[src];
    if (in.gc_lbin) out.gc_lbin[r] = in.gc_lbin[src];
    if (in.cov) out.cov[r] = in.cov[src];
    if (in.cov_lbin) out.cov_lbin[r] = in.cov_lbin[src];
  }
  pmn = warp_min(pmn); pmx = warp_max(pmx); tmx = warp_max(tmx);
  if (lane == 0) { s_red[0][warp] = pmn; s_red[1][warp] = pmx; s_red[2][warp] = tmx; }
  __syncthreads();
  if (t == 0) {
    float a = s_red[0][0], b = s_red[1][0], c = s_red[2][0];
    #pragma unroll
    for (int i = 1; i < T / 32; ++i) { a = fminf(a, s_red[0][i]); b = fmaxf(b, s_red[1][i]); c = fmaxf(c, s_red[2][i]); }
    pmin_c[tile] = a; pmax_c[tile] = b; tmax_c[tile] = c;
  }
  for (int rr = warp; rr < T; rr += T / 32) {
    const int64_t row = tile * T + rr;
    if (row >= n) break;
    const float4* s = reinterpret_cast<const float4*>(sig + (int64_t)perm[row] * DBB_SIG_STRIDE);
    float4* d = reinterpret_cast<float4*>(sig_s + row * DBB_SIG_STRIDE);
    for (int q = lane; q < DBB_SIG_STRIDE / 4; q += 32) d[q] = __ldg(s + q);
  }
}

// statistics of the row tiles of [row0, row1) (they coincide with column tiles only when row0 is a multiple of 128);
// kth != nullptr: the largest k-th best distance of the tile's rows instead (second plan)
__global__ void __launch_bounds__(T) row_tile_kernel(const float* __restrict__ p_s, const float* __restrict__ td_s,
                                                     int64_t row0, int64_t row1, const float* __restrict__ kth, int64_t kth_stride,
                                                     float* __restrict__ pmin_r, float* __restrict__ pmax_r,
                                                     float* __restrict__ tmax_r, float* __restrict__ thr_r) {
  __shared__ float s_red[3][T / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t r = row0 + (int64_t)blockIdx.x * T + t;
  float a = INFINITY, b = -INFINITY, c = -INFINITY;
  if (r < row1) {
    if (kth) {
      c = kth[(r - row0) * kth_stride];
    } else {
      const float pv = p_s[r];
      if (pv == pv) { a = pv; b = pv; }
      if (td_s) c = td_s[r];
    }
  }
  a = warp_min(a); b = warp_max(b); c = warp_max(c);
  if (lane == 0) { s_red[0][warp] = a; s_red[1][warp] = b; s_red[2][warp] = c; }
  __syncthreads();
  if (t == 0) {
    #pragma unroll
    for (int i = 1; i < T / 32; ++i) { a = fminf(a, s_red[0][i]); b = fmaxf(b, s_red[1][i]); c = fmaxf(c, s_red[2][i]); }
    if (kth) thr_r[blockIdx.x] = c;
    else { pmin_r[blockIdx.x] = a; pmax_r[blockIdx.x] = b; tmax_r[blockIdx.x] = c; }
  }
}

struct Stats {
  const float* pmin_c; const float* pmax_c; const float* tmax_c;
  const float* pmin_r; const float* pmax_r; const float* tmax_r; const float* thr_r;
  int64_t n_ct, row0;
  float margin;
  int has_td, second;
};

// lower bound of every distance between row tile rt and column tile ct, and whether this plan visits the pair
__device__ __forceinline__ bool visit(const Stats& S, int rt, int ct, float pmin_r, float pmax_r, float bound_r, float thr, float* lb_out) {
  const float cmin = __ldg(S.pmin_c + ct), cmax = __ldg(S.pmax_c + ct);
  const float gap = fmaxf(fmaxf(cmin - pmax_r, pmin_r - cmax), 0.f);
  const float lb = 2.f * gap - S.margin;
  *lb_out = lb;
  // a tile whose rows are all NaN signatures has an empty interval: no finite distance in the pair
  const bool empty = !(pmin_r <= pmax_r) || !(cmin <= cmax);
  // the tile that holds the row tile's own first row is always taken: every processed row tile then writes its
  // outputs, also one whose rows are all NaN signatures
  const bool diag = ct == (int)((S.row0 + (int64_t)rt * T) / T);
  const bool first = diag || (!empty && (S.has_td ? lb <= fmaxf(bound_r, __ldg(S.tmax_c + ct)) : lb <= 0.f));
  return S.second ? (!empty && lb <= thr && !first) : first;
}

__global__ void __launch_bounds__(256) tile_count_kernel(Stats S, const uint8_t* __restrict__ mine, int32_t* __restrict__ cnt) {
  __shared__ int s_cnt[8];
  const int rt = blockIdx.x, t = threadIdx.x;
  int c = 0;
  if (!mine || mine[rt]) {
    const float a = S.pmin_r[rt], b = S.pmax_r[rt], tb = S.tmax_r[rt], thr = S.second ? S.thr_r[rt] : 0.f;
    float lb;
    for (int ct = t; ct < S.n_ct; ct += 256) c += visit(S, rt, ct, a, b, tb, thr, &lb) ? 1 : 0;
  }
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((t & 31) == 0) s_cnt[t >> 5] = c;
  __syncthreads();
  if (t == 0) { int s = 0; for (int i = 0; i < 8; ++i) s += s_cnt[i]; cnt[rt] = s; }
}

// single GPU: every row tile is mine, slot = row tile
__global__ void all_mine_kernel(int64_t n_rt, uint8_t* __restrict__ mine, int32_t* __restrict__ slot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rt) { mine[i] = 1; slot[i] = (int32_t)i; }
}

// multi-GPU: by_work = the row tiles by decreasing number of tiles; rank r takes by_work[r], by_work[r + world], ...
__global__ void deal_kernel(const int32_t* __restrict__ by_work, int64_t n_rt, int rank, int world,
                            uint8_t* __restrict__ mine, int32_t* __restrict__ slot, int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rt) return;
  const int32_t rt = by_work[i];
  const bool m = (i % world) == rank;
  mine[rt] = m ? 1 : 0;
  slot[rt] = m ? (int32_t)(i / world) : -1;
  if (!m) cnt[rt] = 0;
}

__global__ void order_keys_kernel(const int32_t* __restrict__ cnt, int64_t n_rt, uint32_t* __restrict__ key, int32_t* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rt) { key[i] = 0xffffffffu - (uint32_t)cnt[i]; val[i] = (int32_t)i; }   // descending, stable
}

// keys of one row tile's segment: slot | quantised lower bound (QNONE = not visited) | column tile
__global__ void __launch_bounds__(256) tile_keys_kernel(Stats S, const uint8_t* __restrict__ mine, const int32_t* __restrict__ slot,
                                                        int bits_ct, uint64_t* __restrict__ keys) {
  const int rt = blockIdx.x;
  if (!mine[rt]) return;
  const uint64_t sl = (uint64_t)slot[rt];
  const float a = S.pmin_r[rt], b = S.pmax_r[rt], tb = S.tmax_r[rt], thr = S.second ? S.thr_r[rt] : 0.f;
  uint64_t* seg = keys + sl * S.n_ct;
  for (int ct = threadIdx.x; ct < S.n_ct; ct += 256) {
    float lb;
    const bool v = visit(S, rt, ct, a, b, tb, thr, &lb);
    // distances lie in [0, 2]; 2^18 steps per unit keep "nearest first" (the order is scheduling only)
    const uint32_t q = v ? min((uint32_t)(fminf(fmaxf(lb, 0.f), 2.f) * 262144.f), QNONE - 1u) : QNONE;
    seg[ct] = (sl << (QBITS + bits_ct)) | ((uint64_t)q << bits_ct) | (uint64_t)ct;
  }
}

// tile_off = exclusive sum of the counts (indexed by row tile); info[0] = tiles in all lists, info[1] = row tiles with a
// non-empty list.  One CTA.
__global__ void __launch_bounds__(1024) offsets_kernel(const int32_t* __restrict__ cnt, int64_t n_rt, int64_t* __restrict__ off,
                                                       int64_t* __restrict__ info) {
  __shared__ long long s_sum[1024];
  __shared__ int s_act[1024];
  const int t = threadIdx.x;
  const int64_t per = (n_rt + 1023) / 1024, b = t * per, e = min(b + per, n_rt);
  long long s = 0; int act = 0;
  for (int64_t i = b; i < e; ++i) { s += cnt[i]; act += cnt[i] > 0; }
  s_sum[t] = s; s_act[t] = act;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {                     // inclusive Hillis-Steele scan of the 1024 partial sums
    long long v = 0; int w = 0;
    if (t >= d) { v = s_sum[t - d]; w = s_act[t - d]; }
    __syncthreads();
    s_sum[t] += v; s_act[t] += w;
    __syncthreads();
  }
  long long run = s_sum[t] - s;
  for (int64_t i = b; i < e; ++i) { off[i] = run; run += cnt[i]; }
  if (t == 1023) { off[n_rt] = s_sum[1023]; info[0] = s_sum[1023]; info[1] = s_act[1023]; }
}

__global__ void __launch_bounds__(256) tile_list_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ cnt,
                                                        const int32_t* __restrict__ slot, const int64_t* __restrict__ off,
                                                        int64_t n_ct, int bits_ct, int32_t* __restrict__ tile_list) {
  const int rt = blockIdx.x;
  const int c = cnt[rt];
  if (c == 0) return;
  const uint64_t* seg = keys + (int64_t)slot[rt] * n_ct;
  int32_t* dst = tile_list + off[rt];
  const uint64_t mask = (1ull << bits_ct) - 1ull;
  for (int i = threadIdx.x; i < c; i += 256) dst[i] = (int32_t)(seg[i] & mask);
}

// Lists of the second sweep merged into the first sweep's: one warp per row of the affected row tiles (the first
// info[1] entries of item_order).  Both lists are sorted by (distance, id) and padded with (+inf, -1); the sweeps
// visit disjoint tiles, so no id occurs twice.  Rank merge: position = own index + number of smaller entries of the
// other list.
__global__ void __launch_bounds__(256) merge_lists_kernel(const int32_t* __restrict__ item_order, const int64_t* __restrict__ info,
                                                          int64_t rows, int k, int32_t* __restrict__ idx1, float* __restrict__ dist1,
                                                          const int32_t* __restrict__ idx2, const float* __restrict__ dist2) {
  const int pos = blockIdx.x / (T / 8);
  if (pos >= info[1]) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)item_order[pos] * T + (blockIdx.x % (T / 8)) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float da = INFINITY, db = INFINITY;
  int32_t ia = INT32_MAX, ib = INT32_MAX;
  if (lane < k) {
    da = dist1[r * k + lane]; ia = idx1[r * k + lane]; db = dist2[r * k + lane]; ib = idx2[r * k + lane];
    if (ia < 0) ia = INT32_MAX;
    if (ib < 0) ib = INT32_MAX;
  }
  int ca = 0, cb = 0;
  for (int j = 0; j < k; ++j) {
    const float dbj = __shfl_sync(0xffffffffu, db, j), daj = __shfl_sync(0xffffffffu, da, j);
    const int32_t ibj = __shfl_sync(0xffffffffu, ib, j), iaj = __shfl_sync(0xffffffffu, ia, j);
    ca += (dbj < da || (dbj == da && ibj < ia)) ? 1 : 0;        // entries of the second list strictly before a
    cb += (daj < db || (daj == db && iaj <= ib)) ? 1 : 0;       // entries of the first list before or equal to b
  }
  __syncwarp();
  if (lane < k) {
    const int pa = lane + ca, pb = lane + cb;
    if (pa < k) { dist1[r * k + pa] = da; idx1[r * k + pa] = ia == INT32_MAX ? -1 : ia; }
    if (pb < k) { dist1[r * k + pb] = db; idx1[r * k + pb] = ib == INT32_MAX ? -1 : ib; }
  }
}

int check(const dbb_prune_input* in, const dbb_prune_buffers* buf) {
  DBB_REQUIRE(in && buf, "NULL argument struct");
  DBB_REQUIRE(in->n > 0 && in->n < (1ll << 31), "n out of range");
  DBB_REQUIRE(in->row0 >= 0 && in->row0 < in->row1 && in->row1 <= in->n, "bad row range");
  DBB_REQUIRE(in->world <= 1 || (in->row0 == 0 && in->row1 == in->n && in->rank >= 0 && in->rank < in->world),
              "multi-GPU dealing needs the full row range and 0 <= rank < world");
  DBB_REQUIRE(buf->item_order && buf->tile_off && buf->tile_list && buf->tile_count && buf->mine && buf->info && buf->scratch,
              "incomplete list buffers");
  return DBB_OK;
}

// count -> (deal) -> keys -> sort -> offsets -> lists -> order, for the first (second = 0) or second plan
int build_lists(const dbb_prune_input* in, const dbb_prune_buffers* buf, const Layout& L, int second, int64_t* info, cudaStream_t st) {
  char* sc = (char*)buf->scratch;
  auto F = [&](size_t o) { return (float*)(sc + o); };
  Stats S;
  S.pmin_c = F(L.pmin_c); S.pmax_c = F(L.pmax_c); S.tmax_c = F(L.tmax_c);
  S.pmin_r = F(L.pmin_r); S.pmax_r = F(L.pmax_r); S.tmax_r = F(L.tmax_r); S.thr_r = F(L.thr_r);
  S.n_ct = L.n_ct; S.row0 = in->row0; S.margin = in->margin; S.has_td = in->td_bound != nullptr; S.second = second;
  int32_t* slot = (int32_t*)(sc + L.slot);
  int32_t* cnt = buf->tile_count;
  uint32_t *okey_in = (uint32_t*)(sc + L.okey_in), *okey_out = (uint32_t*)(sc + L.okey_out);
  int32_t *oval_in = (int32_t*)(sc + L.oval_in), *oval_out = (int32_t*)(sc + L.oval_out);
  const unsigned n_rt = (unsigned)L.n_rt, g256 = (n_rt + 255) / 256;
  size_t cub_bytes = L.cub_bytes;
  tile_count_kernel<<<n_rt, 256, 0, st>>>(S, second ? buf->mine : nullptr, cnt);
  if (!second) {
    if (in->world > 1) {
      order_keys_kernel<<<g256, 256, 0, st>>>(cnt, L.n_rt, okey_in, oval_in);
      DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc + L.cub_temp, cub_bytes, okey_in, okey_out, oval_in, oval_out, (int)L.n_rt, 0, 32, st));
      deal_kernel<<<g256, 256, 0, st>>>(oval_out, L.n_rt, in->rank, in->world, buf->mine, slot, cnt);
    } else {
      all_mine_kernel<<<g256, 256, 0, st>>>(L.n_rt, buf->mine, slot);
    }
  }
  const int bits_ct = bits_for(L.n_ct), bits_slot = bits_for(L.n_slots);
  uint64_t *tk_in = (uint64_t*)(sc + L.tkey_in), *tk_out = (uint64_t*)(sc + L.tkey_out);
  // exactly the segments that exist are sorted (the highest ranks own one row tile fewer than n_slots)
  const int64_t n_mine = in->world > 1 ? (L.n_rt - in->rank + in->world - 1) / in->world : L.n_rt;
  tile_keys_kernel<<<n_rt, 256, 0, st>>>(S, buf->mine, slot, bits_ct, tk_in);
  cub_bytes = L.cub_bytes;
  if (n_mine > 0)
    DBB_CUDA_TRY(cub::DeviceRadixSort::SortKeys(sc + L.cub_temp, cub_bytes, tk_in, tk_out, (int)(n_mine * L.n_ct), 0,
                                                bits_ct + QBITS + bits_slot, st));
  offsets_kernel<<<1, 1024, 0, st>>>(cnt, L.n_rt, buf->tile_off, info);
  tile_list_kernel<<<n_rt, 256, 0, st>>>(tk_out, cnt, slot, buf->tile_off, L.n_ct, bits_ct, buf->tile_list);
  order_keys_kernel<<<g256, 256, 0, st>>>(cnt, L.n_rt, okey_in, oval_in);
  cub_bytes = L.cub_bytes;
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc + L.cub_temp, cub_bytes, okey_in, okey_out, oval_in, buf->item_order, (int)L.n_rt, 0, 32, st));
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int64_t dbb_prune_scratch_bytes(int64_t n, int64_t rows, int32_t world) {
  Layout L;
  if (n <= 0 || rows <= 0 || make_layout(n, rows, world, &L) != DBB_OK) return -1;
  return (int64_t)L.total;
}

DBB_EXPORT int dbb_prune_plan_dev(const dbb_prune_input* in, const dbb_prune_buffers* buf, void* stream) {
  int rc = check(in, buf);
  if (rc != DBB_OK) return rc;
  DBB_REQUIRE(in->sig && in->weights && buf->sig_sorted && buf->col_id && buf->p_sorted, "incomplete inputs / sorted buffers");
  DBB_REQUIRE(!in->td_bound || buf->td_bound, "td_bound needs its sorted buffer");
  Layout L;
  if ((rc = make_layout(in->n, in->row1 - in->row0, in->world, &L)) != DBB_OK) return rc;
  DBB_REQUIRE(buf->scratch_bytes >= (int64_t)L.total, "scratch too small (dbb_prune_scratch_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  char* sc = (char*)buf->scratch;
  float* p = (float*)(sc + L.p);
  uint32_t *key_in = (uint32_t*)(sc + L.key_in), *key_out = (uint32_t*)(sc + L.key_out);
  int32_t *val_in = (int32_t*)(sc + L.val_in), *perm = (int32_t*)(sc + L.val_out);
  DBB_REQUIRE(in->margin > 0.f, "margin must be positive");
  DBB_CUDA_TRY(cudaMemsetAsync(buf->info + 4, 0, sizeof(int64_t), st));
  project_kernel<<<(unsigned)((in->n + 7) / 8), 256, 0, st>>>(in->sig, in->weights, in->cls, in->n, 0.3f * in->margin, p, key_in, val_in,
                                                             (unsigned long long*)(buf->info + 4));
  size_t cub_bytes = L.cub_bytes;
  DBB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc + L.cub_temp, cub_bytes, key_in, key_out, val_in, perm, (int)in->n, 0, 32, st));
  Tables ti = {in->length, in->td_bound, in->gc, in->gc_mbin, in->gc_lbin, in->cov, in->cov_lbin};
  TablesOut to = {buf->length, buf->td_bound, buf->gc, buf->gc_mbin, buf->gc_lbin, buf->cov, buf->cov_lbin};
  DBB_REQUIRE((!ti.length || to.length) && (!ti.gc || to.gc) && (!ti.gc_mbin || to.gc_mbin) && (!ti.gc_lbin || to.gc_lbin) &&
              (!ti.cov || to.cov) && (!ti.cov_lbin || to.cov_lbin), "a table without its sorted buffer");
  gather_kernel<<<(unsigned)L.n_ct, T, 0, st>>>(in->sig, p, perm, in->n, ti, to, buf->sig_sorted, buf->p_sorted, buf->col_id,
                                               (float*)(sc + L.pmin_c), (float*)(sc + L.pmax_c), (float*)(sc + L.tmax_c));
  row_tile_kernel<<<(unsigned)L.n_rt, T, 0, st>>>(buf->p_sorted, in->td_bound ? buf->td_bound : nullptr, in->row0, in->row1, nullptr, 0,
                                                 (float*)(sc + L.pmin_r), (float*)(sc + L.pmax_r), (float*)(sc + L.tmax_r), (float*)(sc + L.thr_r));
  DBB_CUDA_TRY(cudaGetLastError());
  return build_lists(in, buf, L, 0, buf->info, st);
}

DBB_EXPORT int dbb_prune_second_dev(const dbb_prune_input* in, const dbb_prune_buffers* buf, const float* d_kth, int64_t kth_stride,
                                    void* stream) {
  int rc = check(in, buf);
  if (rc != DBB_OK) return rc;
  DBB_REQUIRE(d_kth && kth_stride > 0, "the k-th best distances of the first sweep are required");
  Layout L;
  if ((rc = make_layout(in->n, in->row1 - in->row0, in->world, &L)) != DBB_OK) return rc;
  DBB_REQUIRE(buf->scratch_bytes >= (int64_t)L.total, "scratch too small (dbb_prune_scratch_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  char* sc = (char*)buf->scratch;
  row_tile_kernel<<<(unsigned)L.n_rt, T, 0, st>>>(nullptr, nullptr, in->row0, in->row1, d_kth, kth_stride, nullptr, nullptr, nullptr,
                                                 (float*)(sc + L.thr_r));
  DBB_CUDA_TRY(cudaGetLastError());
  return build_lists(in, buf, L, 1, buf->info + 2, st);
}

DBB_EXPORT int dbb_prune_merge_dev(const dbb_prune_buffers* buf, int64_t rows, int32_t k, int32_t* d_idx1, float* d_dist1,
                                   const int32_t* d_idx2, const float* d_dist2, void* stream) {
  DBB_REQUIRE(buf && buf->item_order && buf->info, "NULL list buffers");
  DBB_REQUIRE(rows > 0 && k > 0 && k <= DBB_MAX_K && d_idx1 && d_dist1 && d_idx2 && d_dist2, "bad lists");
  const int64_t n_rt = (rows + T - 1) / T;
  merge_lists_kernel<<<(unsigned)(n_rt * (T / 8)), 256, 0, (cudaStream_t)stream>>>(buf->item_order, buf->info + 2, rows, k, d_idx1, d_dist1,
                                                                                   d_idx2, d_dist2);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}
