// Stage 2 (K2): all-pairs tetranucleotide (L1) distance with fused neighbour predicates, per-row
// neighbourhood counts, optional bit masks and a fused per-row top-k.  The N x N matrix is never
// written.
//
// Reference semantics (tdNeighbours.py:41-51, gcNeighbours.py:43-52, coverageNeighbours.py:41-50,
// distributions.py:75-127): for every ordered pair (i, j), i == j included,
//   td(i,j)  = L1(sig_i, sig_j) <  tdBound[shorter of i, j]            (strict; ties in length -> i)
//   gc(i,j)  = lo <= GC_unbinned - GC_core <= hi, core = longer (ties -> j), bounds by (core mean-GC
//              bin, unbinned length bin)
//   cov(i,j) = lo <= (cov_unb - cov_core) * 100 / cov_core <= hi (cov_core <= 0 -> 0.1)
// GC / coverage arithmetic is FP64 exactly as in the reference; the distance is FP32 (tolerance
// 2e-6 absolute, tests/).
//
// Kernel shape (B200): persistent CTAs (thread 0 also issues the TMA loads), instances by warp count (Cfg<WARPS> below):
// one 8-warp CTA per SM on 128-row tiles, or -- the default -- two 4-warp CTAs per SM on 64-row halves of them.
//   * CTA tile 128 (64) rows x 128 columns; warp w owns rows 16w..16w+15 against all 128 columns, lane
//     (ty = lane/16, tx = lane%16) owns rows 16w+2i+ty and columns tx+16j (i, j < 8): an 8x8 register
//     tile, accumulated as float2 so the inner loop is two packed FADD2 per two dimensions
//     (d = a - b ; acc += |d|) -- one FP32 issue slot per pair-dimension instead of two.
//   * the signature matrix is [n][140] float32 (136 + 4 zero pad = 5 K-chunks of 28 floats); TMA
//     (cp.async.bulk.tensor.2d) stages [128 rows][28 floats] boxes: the A row tile is resident for a
//     whole row sweep (5 boxes), B boxes stream through a 5-stage mbarrier ring.  Row pitch 112 B =
//     7 x 16 B makes every LDS.128 of the tile walk distinct bank groups.
//   * epilogue per column tile: predicates + counts from registers, 16-bit mask pieces by warp ballot,
//     top-k candidates (d <= current k-th best of the row, kept in a register) are rare after the first
//     tiles and are inserted warp-cooperatively into a sorted per-row list in shared memory.
//     Ordering is the total order (distance, j), so results do not depend on tile schedule or on how
//     rows are sharded over GPUs.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdlib>

namespace {

constexpr int TILE = 128;
constexpr int KCH = 28;                      // floats per K chunk
constexpr int NKC = DBB_SIG_STRIDE / KCH;    // 5
#ifndef DBB_K2_NSTAGE
#define DBB_K2_NSTAGE 5
#endif
#ifndef DBB_K2_QTOTAL
#define DBB_K2_QTOTAL 4864
#endif
constexpr int NSTAGE = DBB_K2_NSTAGE;        // <= NKC (the column-scalar double buffer relies on it)
constexpr int CHUNK_BYTES = TILE * KCH * 4;  // 14336
constexpr int MAX_WARPS = 16;                // the kernel is instantiated for 8 warps (8x8 register tiles, <= 255
                                             // registers) and 16 warps (4x8 tiles, <= 128 registers, 4 warps per sub-partition)
#ifndef DBB_K2_PREFETCH
#define DBB_K2_PREFETCH 2
#endif
constexpr int PREFETCH = DBB_K2_PREFETCH;                  // B chunks in flight ahead of the consumers
constexpr int LIST_K = DBB_MAX_K;            // list slots per row
constexpr int QTOTAL = DBB_K2_QTOTAL;                 // queue entries per CTA; per warp QTOTAL / WARPS (>= 256 + one row pair's pushes)

struct RowScalars {       // staged in shared memory per row tile
  int32_t len;
  float tdb;
  int32_t gcm, gcl, covl;
  int32_t pad;
  double gc, cov;
};

// Per-column scalars of one column tile, staged once per CTA by bulk copies that complete on the same
// mbarrier as the tile's first B chunk (double-buffered by tile parity, see issue_chunk()).
struct ColScalars {
  alignas(16) float tdb[TILE];
  alignas(16) int32_t len[TILE];
  alignas(16) int32_t gcm[TILE];
  alignas(16) int32_t gcl[TILE];
  alignas(16) int32_t covl[TILE];
  alignas(16) double gc[TILE];
  alignas(16) double cov[TILE];
};

struct Params {
  int64_t n, row0, row1;
  const int32_t* length;
  const float* td_bound;
  const double* gc; const int32_t* gc_mbin; const int32_t* gc_lbin; const double* gc_lo; const double* gc_hi; int32_t gc_nl;
  const double* cov; const int32_t* cov_lbin; const double* cov_lo; const double* cov_hi;
  int32_t k, topk_filter;
  int32_t* knn_idx; float* knn_dist; int32_t* count; int64_t* neighbour_bp;
  uint32_t* td_mask; uint32_t* gc_mask; uint32_t* cov_mask; int64_t mask_words;
  // work items: row tiles [0, n_full_rt) are swept over all columns by one CTA; each remaining row tile is
  // split into n_seg column segments (separate items) whose partial results are merged afterwards
  int32_t n_full_rt, n_seg, n_items, debug_skip_epilogue;
  int32_t seg_tiles;        // > 0: a tile list of L entries uses min(n_seg, ceil(L / seg_tiles)) of its n_seg segment slots
  float* gthr;              // [split row tiles][TILE] k-th best distance of a row SHARED by the segments of its row tile
                            // (an upper bound of the final k-th best: any full segment list proves one), or nullptr
  int32_t opaque_zero;         // always 0 (see the k4 loop)
  uint32_t* item_counter;
  unsigned long long* stats;   // profiling builds only (-DDBB_K2_STATS, run with DBB_K2_STATS=1): [0] queued pairs, [1] drains, [2] top-k candidates
  int32_t* part_idx; float* part_dist; int32_t* part_cnt; int64_t* part_bp;
  // optional sparse sweep (all NULL = every row tile visits column tiles 0..n_ct-1 in order):
  const int32_t* item_order;   // [n_rt] row tile handled by work position p (whole sweeps first, then the split ones)
  const int64_t* tile_off;     // [n_rt + 1] offsets into tile_list, indexed by row tile
  const int32_t* tile_list;    // column tiles each row tile visits, ascending
  const int32_t* col_id;       // [n] the caller's id of column j (tie-breaks and knn_idx use it instead of j)
};

// Shared memory of one CTA.  TROWS = rows of the CTA's row tile (128, or 64 for the two-CTAs-per-SM instance), NST = stages of
// the B ring, QT = queue entries of the CTA.
template <int TROWS, int NST, int QT>
struct SmemLayoutT {
  alignas(128) float a[NKC][TROWS][KCH];
  alignas(128) float b[NST][TILE][KCH];
  float list_d[TROWS][LIST_K];
  int32_t list_j[TROWS][LIST_K];
  float2 q[QT];                          // per-warp queue of pairs that may be a neighbour or a top-k candidate:
                                         //   .x = distance, .y = bits of (row within the warp's rows) << 7 | (column
                                         //   within the tile); the every-pair path uses the first 256 .x as scratch
  RowScalars rows[TROWS];
  uint32_t bp_lo[TROWS], bp_hi[TROWS];   // neighbourhood base pairs per row (64-bit as two native 32-bit atomics)
  int32_t cnt[TROWS];                    // neighbourhood size per row
  float thr[TROWS];                      // current k-th best distance per row (+inf until k found)
  int32_t qn[MAX_WARPS];                 // queue fill per warp
  ColScalars cols[2];                    // column scalars of the t-th visited tile live in cols[t & 1]
  alignas(8) uint64_t full[NST];
  uint64_t empty[NST];
  uint64_t a_full;
  int32_t item;
};

// Instances of the kernel by warp count:
//    8 warps: one CTA per SM, 128-row tiles, 8x8 register tiles, 5-stage B ring (a whole column tile in flight)
//   16 warps: as above with 4x8 register tiles (-DDBB_K2_WARPS16)
//    4 warps: TWO CTAs per SM, each with a 64-row tile (one half of a 128-row work item), 8x8 register tiles, 2-stage
//             B ring.  The eight warps of the big CTA reach their epilogues together (the ring couples them), and the
//             FP32 pipe idles meanwhile (30 % of the cycles of the pruned sweep on C2); two independent CTAs drift
//             apart, so one warp of every sub-partition is mostly in the FADD2 loop.
template <int WARPS> struct Cfg {
  static constexpr int TROWS = WARPS == 4 ? 64 : TILE;
  static constexpr int NST = WARPS == 4 ? 2 : NSTAGE;
  static constexpr int PRE = WARPS == 4 ? 1 : PREFETCH;
  static constexpr int QT = WARPS == 4 ? QTOTAL / 2 : QTOTAL;
  static constexpr int MINB = WARPS == 4 ? 2 : 1;
  using Smem = SmemLayoutT<TROWS, NST, QT>;
};

static_assert(NSTAGE <= NKC, "column-scalar double buffer needs NSTAGE <= NKC");
static_assert(sizeof(Cfg<8>::Smem) <= 232448, "SmemLayout exceeds the 227 KB per-CTA shared memory limit");
static_assert(2 * (sizeof(Cfg<4>::Smem) + 1024) <= 233472, "two 4-warp CTAs do not fit the 228 KB of an SM");

// ---- mbarrier / TMA primitives ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int32_t x, int32_t y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// 1-D bulk copy global -> shared (bytes and both addresses multiples of 16), completing on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Stages `v` elements of a per-scaffold array into shared memory: the 16-byte-granular part as one bulk copy
// (returned byte count joins the mbarrier's expected transaction bytes), the <16-byte tail of the last,
// partial column tile by the calling thread (ordered before its arrive on the barrier).
template <typename T>
__device__ __forceinline__ uint32_t stage_cols(T* dst, const T* base, int64_t c0, int v, uint64_t* bar, bool issue) {
  if (base == nullptr) return 0u;
  const T* src = base + c0;
  const uint32_t bytes = ((uint32_t)v * (uint32_t)sizeof(T)) & ~15u;
  if (!issue) {
    for (int e = (int)(bytes / sizeof(T)); e < v; ++e) dst[e] = src[e];
  } else if (bytes) {
    bulk_load_1d(dst, src, bytes, bar);
  }
  return bytes;
}

__device__ __forceinline__ float2 absdiff_acc(float2 acc, float2 a, float2 b) {
  float2 d = __fadd2_rn(a, make_float2(-b.x, -b.y));
  return __fadd2_rn(acc, make_float2(fabsf(d.x), fabsf(d.y)));
}

// The dynamic shared memory of the CTA *is* the SmemLayout.  Every function re-derives it from the extern array
// (instead of taking a reference parameter) so that the compiler keeps the shared address space: a generic
// reference costs LD + QSPC-guarded atomics in the out-of-line slow paths.
extern __shared__ __align__(128) uint8_t smem_raw[];
template <int WARPS>
__device__ __forceinline__ typename Cfg<WARPS>::Smem& smem() { return *reinterpret_cast<typename Cfg<WARPS>::Smem*>(smem_raw); }

template <typename SM>
__device__ __forceinline__ void add_bp(SM& sm, int lr, uint32_t v) {
  const uint32_t old = atomicAdd(&sm.bp_lo[lr], v);
  if (old + v < old) atomicAdd(&sm.bp_hi[lr], 1u);
}

__device__ __forceinline__ bool cov_predicate(double cu, double cc, double lo, double hi);

// the caller's id of column position j (the lists hold positions; ids only decide exact distance ties and are
// substituted when the lists are written out)
__device__ __forceinline__ int32_t id_of(const int32_t* col_id, int32_t j) {
  return (col_id && j != INT_MAX) ? __ldg(col_id + j) : j;
}

// warp-cooperative insertion of (d, c) into the sorted list of `row`; returns the new k-th best
template <typename SM>
__device__ __forceinline__ float list_insert(SM& sm, int row, float d, int32_t c, int k, int lane, const int32_t* col_id) {
  float ed = sm.list_d[row][lane];
  int32_t ej = sm.list_j[row][lane];
  bool before = ed < d;
  if (ed == d) before = id_of(col_id, ej) < id_of(col_id, c);      // rare: an exact tie
  const int pos = __popc(__ballot_sync(0xffffffffu, before));
  float pd = __shfl_up_sync(0xffffffffu, ed, 1);
  int32_t pj = __shfl_up_sync(0xffffffffu, ej, 1);
  if (pos < k) {
    if (lane == pos) { ed = d; ej = c; }
    else if (lane > pos) { ed = pd; ej = pj; }
    sm.list_d[row][lane] = ed;
    sm.list_j[row][lane] = ej;
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, ed, k - 1);
}

// Slow path of the epilogue (one call per row pair that may hold a neighbour or a top-k candidate):
// evaluates the predicates of the 2 x 128 pairs whose distances sit in sm.scratch[warp], updates the
// per-row counts / base pairs / masks / top-k lists.  Whole warp, lanes 0-15 -> row lr0, 16-31 -> lr0+1.
template <int FLAGS, int WARPS>
__device__ __noinline__ void process_row_pair(const Params& P, int par, int warp, int lane, int lr0,
                                              int64_t row_base, int64_t col_base) {
  auto& sm = smem<WARPS>();
  const ColScalars& cs = sm.cols[par];
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_MASK = (FLAGS & 4) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int QCAP = Cfg<WARPS>::QT / WARPS;
  const int ty = lane >> 4, tx = lane & 15;
  const int lr = lr0 + ty;
  const int64_t gi = row_base + lr;
  const int64_t n = P.n;
  const RowScalars rs = sm.rows[lr];
  const bool has_gc = F_GCCOV && P.gc != nullptr, has_cov = F_GCCOV && P.cov != nullptr;
  const bool want_cnt = P.count != nullptr || P.neighbour_bp != nullptr;
  const int k = P.k;
  #pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const float d = sm.q[warp * QCAP + (j * 32 + lane)].x;
    const int64_t cj = col_base + tx + 16 * j;
    const bool col_ok = cj < n;
    int32_t len_j = 0;
    bool p_td = col_ok, p_gc = col_ok, p_cov = col_ok;
    if (col_ok) {
      const int c = tx + 16 * j;
      len_j = cs.len[c];
      const bool i_core = rs.len > len_j;
      if (F_TD) p_td = d < (i_core ? cs.tdb[c] : rs.tdb);
      if (has_gc) {
        const int m = i_core ? rs.gcm : cs.gcm[c], l = i_core ? cs.gcl[c] : rs.gcl;
        const double lo = __ldg(P.gc_lo + m * P.gc_nl + l), hi = __ldg(P.gc_hi + m * P.gc_nl + l);
        const double gj = cs.gc[c];
        const double diff = i_core ? (gj - rs.gc) : (rs.gc - gj);
        p_gc = diff >= lo && diff <= hi;
      }
      if (has_cov) {
        const double vj = cs.cov[c];
        const int l = i_core ? cs.covl[c] : rs.covl;
        p_cov = cov_predicate(i_core ? vj : rs.cov, i_core ? rs.cov : vj, __ldg(P.cov_lo + l), __ldg(P.cov_hi + l));
      }
    }
    const bool all = p_td && p_gc && p_cov;
    if (want_cnt) {
      const uint32_t ball = __ballot_sync(0xffffffffu, all);
      if (tx == 0) sm.cnt[lr] += __popc((ball >> (16 * ty)) & 0xffffu);
      if (all) add_bp(sm, lr, (uint32_t)len_j);
    }
    if (F_MASK) {
      // one ballot = 16 consecutive columns of row lr0 (low half) and of row lr0+1 (high half)
      const uint32_t bt = __ballot_sync(0xffffffffu, p_td);
      const uint32_t bg = __ballot_sync(0xffffffffu, p_gc);
      const uint32_t bc = __ballot_sync(0xffffffffu, p_cov);
      const int64_t w = (col_base >> 5) + (j >> 1);
      if (tx == 0 && gi < P.row1 && w < P.mask_words) {
        const int sh = 16 * (j & 1);
        const size_t o = (size_t)(gi - P.row0) * P.mask_words + w;
        // the two 16-bit halves of a word are written by the same lane in consecutive j
        if (P.td_mask && F_TD) { const uint32_t v = ((bt >> (16 * ty)) & 0xffffu) << sh; if (sh) P.td_mask[o] |= v; else P.td_mask[o] = v; }
        if (P.gc_mask && has_gc) { const uint32_t v = ((bg >> (16 * ty)) & 0xffffu) << sh; if (sh) P.gc_mask[o] |= v; else P.gc_mask[o] = v; }
        if (P.cov_mask && has_cov) { const uint32_t v = ((bc >> (16 * ty)) & 0xffffu) << sh; if (sh) P.cov_mask[o] |= v; else P.cov_mask[o] = v; }
      }
    }
    if (F_TOPK) {
      bool cand = col_ok && (cj != gi) && (d <= sm.thr[lr]);
      if (P.topk_filter & 1) cand = cand && p_td;
      if (P.topk_filter & 2) cand = cand && p_gc;
      if (P.topk_filter & 4) cand = cand && p_cov;
      uint32_t m = __ballot_sync(0xffffffffu, cand);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int crow = lr0 + (src >> 4);
        const float cd = __shfl_sync(0xffffffffu, d, src);
        if (cd <= sm.thr[crow]) {            // the list may have tightened since the ballot
          const int32_t cc = (int32_t)(col_base + (src & 15) + 16 * j);
          const float nt = list_insert(sm, crow, cd, cc, k, lane, P.col_id);
          if (lane == 0) sm.thr[crow] = fminf(nt, sm.thr[crow]);     // (a shared bound may already be tighter than this list)
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

// Coverage predicate of one pair (coverageNeighbours.py:42-50, distributions.py:106-127), FP64-exact.  Most pairs
// sit far from both bounds: an FP32 evaluation whose error margin clears them decides those without the FP64
// division; anything closer (or non-finite) takes the reference's arithmetic.
__device__ __forceinline__ bool cov_predicate(double cu, double cc, double lo, double hi) {
  const double eff = cc > 0 ? cc : 0.1;
  const double num = cu - eff;
  const float num32 = (float)num, eff32 = (float)eff;
  const float d32 = __fdividef(num32 * 100.f, eff32);
  const float lo32 = (float)lo, hi32 = (float)hi;
  const float margin = 4e-6f * fmaxf(fabsf(d32), fmaxf(fabsf(lo32), fabsf(hi32))) + 1e-30f;
  // FP32 error here is < 3e-7 relative; operands outside the comfortable float range, NaN and inf are "uncertain"
  const bool certain = fabsf(d32 - lo32) > margin && fabsf(d32 - hi32) > margin &&
                       eff32 > 1e-30f && eff32 < 1e30f && fabsf(num32) < 1e30f;
  bool p = d32 >= lo32 && d32 <= hi32;
  if (!certain) {
    const double diff = num * 100.0 / eff;
    p = diff >= lo && diff <= hi;
  }
  return p;
}

// Steady-state slow path: the warp's queue holds (distance, row, column) of the pairs of this column tile that
// passed the conservative in-register test (a few per cent: every TD neighbour is one).  Lane-parallel
// predicates and counts, then the (rare) top-k candidates are inserted one at a time, warp-cooperatively.
// With GC / coverage predicates the drain runs in two passes: pass A evaluates the shared-memory-only
// predicates (TD, coverage) and compacts the survivors in place; pass B looks up the GC bounds (a
// [mean-GC][length] table in global memory, L2 latency) for those few only.
template <int FLAGS, int WARPS>
__device__ __noinline__ void drain_queue(const Params& P, int par, int warp, int lane, int qn,
                                         int64_t row_base, int64_t col_base) {
  auto& sm = smem<WARPS>();
  const ColScalars& cs = sm.cols[par];
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int QCAP = Cfg<WARPS>::QT / WARPS, RW = Cfg<WARPS>::TROWS / WARPS;
  constexpr uint32_t ALL_BIT = 1u << 31, CAND_BIT = 1u << 30, CODE_MASK = (1u << 11) - 1u;
  const int64_t n = P.n;
  const bool has_gc = F_GCCOV && P.gc != nullptr, has_cov = F_GCCOV && P.cov != nullptr;
  const bool want_cnt = P.count != nullptr || P.neighbour_bp != nullptr;
  const int k = P.k;
  const uint32_t lt_mask = (1u << lane) - 1u;
  float2* q = &sm.q[warp * QCAP];
#ifdef DBB_K2_STATS
  if (P.stats && lane == 0) { atomicAdd(P.stats, (unsigned long long)qn); atomicAdd(P.stats + 1, 1ull); }
#endif

  if (F_GCCOV) {
    int w = 0;                             // survivors written so far (w <= base: in-place compaction is safe)
    #pragma unroll 1
    for (int base = 0; base < qn; base += 32) {
      const int e = base + lane;
      const bool valid = e < qn;
      const float2 qe = valid ? q[e] : make_float2(INFINITY, 0.f);
      const float d = qe.x;
      const uint32_t code = __float_as_uint(qe.y);
      const int c = (int)(code & 127u);
      const int lr = warp * RW + (int)(code >> 7);
      const int64_t cj = col_base + c;
      const bool col_ok = valid && cj < n;
      bool p_td = col_ok, p_cov = col_ok;
      if (col_ok) {
        const RowScalars rs = sm.rows[lr];
        const bool i_core = rs.len > cs.len[c];
        if (F_TD) p_td = d < (i_core ? cs.tdb[c] : rs.tdb);
        if (has_cov) {
          const double vj = cs.cov[c];
          const int l = i_core ? cs.covl[c] : rs.covl;
          p_cov = cov_predicate(i_core ? vj : rs.cov, i_core ? rs.cov : vj, __ldg(P.cov_lo + l), __ldg(P.cov_hi + l));
        }
      }
      const bool all = want_cnt && p_td && p_cov;
      bool cand = F_TOPK && col_ok && (cj != row_base + lr) && (d <= sm.thr[lr]);
      if (P.topk_filter & 1) cand = cand && p_td;
      if (P.topk_filter & 4) cand = cand && p_cov;
      const bool keep = all || cand;
      const uint32_t km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int o = w + __popc(km & lt_mask);
        q[o] = make_float2(d, __uint_as_float(code | (all ? ALL_BIT : 0u) | (cand ? CAND_BIT : 0u)));
      }
      w += __popc(km);
      __syncwarp();
    }
    qn = w;
  }

  #pragma unroll 1
  for (int base = 0; base < qn; base += 32) {
    const int e = base + lane;
    const bool valid = e < qn;
    const float2 qe = valid ? q[e] : make_float2(INFINITY, 0.f);
    const float d = qe.x;
    const uint32_t code = __float_as_uint(qe.y);
    const int c = (int)(code & 127u);
    const int lr = warp * RW + (int)((code & CODE_MASK) >> 7);
    const int64_t cj = col_base + c;
    const int64_t gi = row_base + lr;
    bool all, cand;
    int32_t len_j = 0;
    if (F_GCCOV) {
      // pass B: the GC predicate decides what pass A let through
      bool p_gc = valid;
      if (valid) len_j = cs.len[c];
      if (valid && has_gc) {
        const RowScalars rs = sm.rows[lr];
        const bool i_core = rs.len > len_j;
        const int m = i_core ? rs.gcm : cs.gcm[c], l = i_core ? cs.gcl[c] : rs.gcl;
        const double lo = __ldg(P.gc_lo + m * P.gc_nl + l), hi = __ldg(P.gc_hi + m * P.gc_nl + l);
        const double gj = cs.gc[c];
        const double diff = i_core ? (gj - rs.gc) : (rs.gc - gj);
        p_gc = diff >= lo && diff <= hi;
      }
      all = (code & ALL_BIT) && p_gc;
      cand = (code & CAND_BIT) != 0;
      if (P.topk_filter & 2) cand = cand && p_gc;
    } else {
      const bool col_ok = valid && cj < n;
      bool p_td = col_ok;
      if (col_ok) {
        const RowScalars rs = sm.rows[lr];
        len_j = cs.len[c];
        if (F_TD) p_td = d < ((rs.len > len_j) ? cs.tdb[c] : rs.tdb);
      }
      all = want_cnt && p_td;
      cand = F_TOPK && col_ok && (cj != gi) && (d <= sm.thr[lr]);
      if (P.topk_filter & 1) cand = cand && p_td;
    }
    if (all) {
      atomicAdd(&sm.cnt[lr], 1);
      add_bp(sm, lr, (uint32_t)len_j);
    }
    if (F_TOPK) {
      uint32_t m = __ballot_sync(0xffffffffu, cand);
      if (P.debug_skip_epilogue == 3) m = 0;
#ifdef DBB_K2_STATS
      if (P.stats && lane == 0 && m) atomicAdd(P.stats + 2, (unsigned long long)__popc(m));
#endif
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int crow = __shfl_sync(0xffffffffu, lr, src);
        const float cd = __shfl_sync(0xffffffffu, d, src);
        const int32_t cc = (int32_t)__shfl_sync(0xffffffffu, (int)cj, src);
        if (cd <= sm.thr[crow]) {          // the list may have tightened since the candidate was flagged
          const float nt = list_insert(sm, crow, cd, cc, k, lane, P.col_id);
          if (lane == 0) sm.thr[crow] = fminf(nt, sm.thr[crow]);     // (a shared bound may already be tighter than this list)
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

// Steady-state test of one row pair (lanes 0-15: row 2i, lanes 16-31: row 2i+1 of the warp's rows; d0..d7 = the
// lane's distances to columns tx + 16 j).  A pair is queued only if it can be a TD neighbour (d below the row's or
// the column's bound -- the exact rule is re-applied in the drain) or can enter the row's top-k (d <= k-th best).
// On the first tile of a sweep the lists are empty; instead of flooding them with all 128 columns the threshold is
// the largest of the 16 lanes' m-th smallest distances, m = ceil((k+1)/16): at least k+1 columns pass.
// Out of line on purpose (one copy of the code instead of eight per column tile); everything it needs arrives in
// registers -- reading Params here means generic loads behind the call (profiles/r01_pairs_notes.md).
// flags: bit 0 first tile of the sweep, bit 1 TD neighbours are wanted (counts requested).
template <int FLAGS, int WARPS>
__device__ __noinline__ void test_row_pair(int par, int warp, int lane, int i, int flags, int ncol, int k, float d0,
                                           float d1, float d2, float d3, float d4, float d5, float d6, float d7) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int QCAP = Cfg<WARPS>::QT / WARPS, RW = Cfg<WARPS>::TROWS / WARPS;
  auto& sm = smem<WARPS>();
  const ColScalars& cs = sm.cols[par];
  const int ty = lane >> 4, tx = lane & 15;
  const int lr = warp * RW + 2 * i + ty;
  const bool td_hot = F_TD && (flags & 2);
  float d[8] = {d0, d1, d2, d3, d4, d5, d6, d7};

  float thr_i = F_TOPK ? sm.thr[lr] : -INFINITY;
  if (flags & 1) {
    const int m = (k + 16) / 16;                         // 1, 2 or 3
    float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;   // the lane's three smallest valid distances
    #pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = (tx + 16 * j < ncol && d[j] == d[j]) ? d[j] : INFINITY;
      const float t1 = fminf(m1, v), c1 = fmaxf(m1, v);
      const float t2 = fminf(m2, c1), c2 = fmaxf(m2, c1);
      m1 = t1; m2 = t2; m3 = fminf(m3, c2);
    }
    float tau = m == 1 ? m1 : (m == 2 ? m2 : m3);
    #pragma unroll
    for (int o = 1; o < 16; o <<= 1) tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, o));
    thr_i = fminf(tau, thr_i);               // (the row's shared bound, if another segment already has one)
  }
  // row-side bound: the row's TD bound or its current k-th best, whichever is larger; column side: the column's
  // TD bound (-inf beyond the last column, see issue_chunk)
  const float u = fmaxf(td_hot ? sm.rows[lr].tdb : -INFINITY, thr_i);
  uint32_t bits = 0;
  #pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = td_hot ? cs.tdb[tx + 16 * j] : -INFINITY;
    bits |= (d[j] <= u || d[j] <= t) ? (1u << j) : 0u;
  }
  if (bits) {
    int slot = warp * QCAP + atomicAdd(&sm.qn[warp], __popc(bits));    // reserve queue slots
    const uint32_t code0 = (uint32_t)((2 * i + ty) << 7) | (uint32_t)tx;
    #pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (bits & (1u << j)) {
        sm.q[slot] = make_float2(d[j], __uint_as_float(code0 | (uint32_t)(16 * j)));
        ++slot;
      }
    }
  }
}

// Producer step (one thread): B chunk (column tile `ctile`, K-chunk `kc`) goes to ring stage qg % NSTAGE once the
// consumers have released it.  The column scalars of the tile travel with its first chunk into cols[ctile & 1]:
// that buffer was last read by the epilogues of tile ctile-2, and the empty-barrier wait (chunk qg - NSTAGE, which
// belongs to tile ctile-1 because NSTAGE <= NKC, has been consumed by every warp) implies those are over.
// (Inlined: an out-of-line call was 1.5 % slower, profiles/r01_pairs_notes.md.)
template <int FLAGS, int WARPS>
__device__ __forceinline__ void issue_chunk(const CUtensorMap* tmap, const Params& P, uint32_t qg, int ctile, int kc, int par) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0;
  constexpr int NSTAGE = Cfg<WARPS>::NST;
  auto& sm = smem<WARPS>();
  const bool has_gc = F_GCCOV && P.gc != nullptr, has_cov = F_GCCOV && P.cov != nullptr;
  const int64_t n = P.n;
  const uint32_t st = qg % NSTAGE;
  uint64_t* bar = &sm.full[st];
  if (qg >= NSTAGE) mbar_wait(&sm.empty[st], ((qg / NSTAGE) - 1) & 1);
  if (kc != 0) {
    mbar_arrive_expect_tx(bar, CHUNK_BYTES);
    tma_load_2d(&sm.b[st][0][0], tmap, kc * KCH, ctile * TILE, bar);
    return;
  }
  ColScalars& cs = sm.cols[par];
  const int64_t c0 = (int64_t)ctile * TILE;
  if (c0 + TILE <= n) {
    // full column tile: fixed sizes, no tails
    const uint32_t bytes = CHUNK_BYTES + TILE * 4 + (F_TD ? TILE * 4 : 0) + (has_gc ? TILE * 16 : 0) + (has_cov ? TILE * 12 : 0);
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_2d(&sm.b[st][0][0], tmap, 0, ctile * TILE, bar);
    bulk_load_1d(cs.len, P.length + c0, TILE * 4, bar);
    if (F_TD) bulk_load_1d(cs.tdb, P.td_bound + c0, TILE * 4, bar);
    if (has_gc) {
      bulk_load_1d(cs.gc, P.gc + c0, TILE * 8, bar);
      bulk_load_1d(cs.gcm, P.gc_mbin + c0, TILE * 4, bar);
      bulk_load_1d(cs.gcl, P.gc_lbin + c0, TILE * 4, bar);
    }
    if (has_cov) {
      bulk_load_1d(cs.cov, P.cov + c0, TILE * 8, bar);
      bulk_load_1d(cs.covl, P.cov_lbin + c0, TILE * 4, bar);
    }
    return;
  }
  // last, partial tile: 16-byte-granular bulk copies + element tails written by this thread before its arrive;
  // TD bounds beyond the last column read as -inf (test_row_pair does not check column validity)
  const int v = (int)(n - c0);
  if (F_TD) for (int e = v; e < TILE; ++e) cs.tdb[e] = -INFINITY;
  uint32_t bytes = CHUNK_BYTES;
  #pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const bool issue = pass == 1;
    if (issue) {
      mbar_arrive_expect_tx(bar, bytes);
      tma_load_2d(&sm.b[st][0][0], tmap, 0, ctile * TILE, bar);
    }
    uint32_t b = stage_cols(cs.len, P.length, c0, v, bar, issue);
    if (F_TD) b += stage_cols(cs.tdb, P.td_bound, c0, v, bar, issue);
    if (has_gc) {
      b += stage_cols(cs.gc, P.gc, c0, v, bar, issue);
      b += stage_cols(cs.gcm, P.gc_mbin, c0, v, bar, issue);
      b += stage_cols(cs.gcl, P.gc_lbin, c0, v, bar, issue);
    }
    if (has_cov) {
      b += stage_cols(cs.cov, P.cov, c0, v, bar, issue);
      b += stage_cols(cs.covl, P.cov_lbin, c0, v, bar, issue);
    }
    bytes += b;
  }
}

// FLAGS: bit0 TD predicate, bit1 GC and/or coverage predicates, bit2 masks, bit3 top-k, bit4 Euclidean distance
// (sqrt of the sum of squared differences, FFMA2 instead of the second FADD2) instead of the reference's Manhattan
template <int FLAGS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, Cfg<WARPS>::MINB) pairs_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap,
                                                                           const __grid_constant__ Params P) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_MASK = (FLAGS & 4) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr bool F_L2 = (FLAGS & 16) != 0;
  auto fin = [](float2 a) -> float { return F_L2 ? __fsqrt_rn(a.x + a.y) : a.x + a.y; };   // accumulator -> distance
  constexpr int TROWS = Cfg<WARPS>::TROWS;  // rows of this CTA's row tile: 128, or 64 (one half of a 128-row work item)
  constexpr int HALVES = TILE / TROWS;      // CTAs that share a work item
  constexpr int NSTAGE = Cfg<WARPS>::NST, PREFETCH = Cfg<WARPS>::PRE;
  constexpr int RW = TROWS / WARPS;         // rows per warp: 16 or 8
  constexpr int RI = RW / 2;                // rows per lane: 8 or 4 (lane owns rows warp*RW + 2i + ty)
  constexpr int QCAP = Cfg<WARPS>::QT / WARPS;
  auto& sm = smem<WARPS>();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n = P.n;
  const int64_t n_rows = P.row1 - P.row0;
  const int n_rt = (int)((n_rows + TILE - 1) / TILE);
  const int n_ct = (int)((n + TILE - 1) / TILE);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], WARPS); }
    mbar_init(&sm.a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int ty = lane >> 4, tx = lane & 15;
  const bool has_gc = F_GCCOV && P.gc != nullptr, has_cov = F_GCCOV && P.cov != nullptr;
  const bool want_cnt = P.count != nullptr || P.neighbour_bp != nullptr;
  // without a TD bound the distance cannot veto a neighbour: every pair takes the full path
  const bool always_slow = F_MASK || (want_cnt && !F_TD);
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t q = 0;          // B chunks consumed so far by this CTA (ring position / mbarrier phase)
  uint32_t items_done = 0;
  if (P.stats && tid == 0 && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.stats[8] = t;
  }

  while (true) {
    __syncthreads();                       // everyone is done with the previous item's A tile and lists
    if (tid == 0) sm.item = (int32_t)atomicAdd(P.item_counter, 1u);
    __syncthreads();
    // a work unit is an item (HALVES = 1) or one 64-row half of it; the halves of an item are consecutive units
    const int item = sm.item / HALVES;
    const int lro = (sm.item % HALVES) * TROWS;        // first row of this CTA's part within the 128-row tile
    if (item >= P.n_items) {
      if (P.stats && tid == 0 && blockIdx.x < 1000) {      // DBB_K2_STATS: when did this CTA run out of work
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.stats[16 + blockIdx.x] = t;
      }
      break;
    }
    // work position -> row tile, segment of its tile sequence [t0, t1)
    int pos = item, seg = -1, srt = 0;
    if (item >= P.n_full_rt) {
      // segment-major: segment 0 of every split row tile, then segment 1 of every one, ... -- when a later segment of a
      // row tile starts, its earlier ones have (mostly) finished and left their k-th best distances in gthr
      const int t = item - P.n_full_rt;
      const int n_split = (P.n_items - P.n_full_rt) / P.n_seg;
      seg = t / n_split;
      srt = t % n_split;                   // index among the split row tiles (slot of the partial results)
      pos = P.n_full_rt + srt;
    }
    const int rt = P.item_order ? __ldg(P.item_order + pos) : pos;
    const int64_t list_base = P.tile_off ? __ldg(P.tile_off + rt) : 0;
    const int n_tiles = P.tile_off ? (int)(__ldg(P.tile_off + rt + 1) - list_base) : n_ct;
    // a segment of a full sweep is a contiguous range of column tiles; a segment of a tile list takes every
    // n_seg-th entry (the lists come nearest-first: every segment then sees near tiles early and its top-k
    // threshold tightens at once)
    const bool strided = P.tile_list != nullptr && seg >= 0;
    // segments in use: all n_seg, or (seg_tiles > 0) as many as keep a segment near seg_tiles entries -- the sweep ends
    // with its longest item, so long lists are cut finely, while short ones stay whole (every segment refills its own
    // top-k lists); the unused slots of a row tile write empty partial results
    int n_use = P.n_seg;
    if (strided && P.seg_tiles > 0) n_use = max(1, min(P.n_seg, (n_tiles + P.seg_tiles - 1) / P.seg_tiles));
    const int64_t row_base = P.row0 + (int64_t)rt * TILE + lro;
    if (row_base >= P.row1) continue;        // (the second half of a last, short row tile)
    if (strided && seg >= n_use) {
      for (int lr = tid; lr < TROWS; lr += WARPS * 32) {
        if (row_base + lr < P.row1) {
          const int64_t o = ((int64_t)srt * TILE + lro + lr) * P.n_seg + seg;
          P.part_cnt[o] = 0;
          P.part_bp[o] = 0;
          if (F_TOPK) for (int e = 0; e < P.k; ++e) { P.part_idx[o * P.k + e] = INT_MAX; P.part_dist[o * P.k + e] = INFINITY; }
        }
      }
      continue;
    }
    const int t0 = (seg < 0 || strided) ? 0 : (int)(((int64_t)n_tiles * seg) / P.n_seg);
    const int t1 = seg < 0 ? n_tiles : (strided ? (n_tiles - seg + n_use - 1) / n_use : (int)(((int64_t)n_tiles * (seg + 1)) / P.n_seg));
    auto tile_at = [&](int t) -> int {
      return P.tile_list ? __ldg(P.tile_list + list_base + (strided ? seg + t * n_use : t)) : t;
    };
    const uint32_t item_chunks = (uint32_t)(t1 - t0) * NKC;
    const uint32_t qbase = q;

    // thread 0 doubles as the TMA producer: A tile of this item + its first PREFETCH B chunks
    if (tid == 0) {
      mbar_arrive_expect_tx(&sm.a_full, NKC * TROWS * KCH * 4);
      for (int kc = 0; kc < NKC; ++kc) tma_load_2d(&sm.a[kc][0][0], &tmap_a, kc * KCH, (int32_t)row_base, &sm.a_full);
      for (uint32_t l = 0; l < (uint32_t)PREFETCH && l < item_chunks; ++l) {
        const int ts = t0 + (int)(l / NKC);
        issue_chunk<FLAGS, WARPS>(&tmap, P, qbase + l, tile_at(ts), (int)(l % NKC), ts & 1);
      }
    }
    // stage row scalars and reset the per-row state of this warp's 16 rows
    if (lane < RW) {
      const int lr = warp * RW + lane;
      const int64_t gr = row_base + lr;
      RowScalars rs;
      rs.len = 0; rs.tdb = 0.f; rs.gcm = 0; rs.gcl = 0; rs.covl = 0; rs.pad = 0; rs.gc = 0.0; rs.cov = 0.0;
      if (gr < n) {
        rs.len = P.length[gr];
        if (F_TD) rs.tdb = P.td_bound[gr];
        if (has_gc) { rs.gc = P.gc[gr]; rs.gcm = P.gc_mbin[gr]; rs.gcl = P.gc_lbin[gr]; }
        if (has_cov) { rs.cov = P.cov[gr]; rs.covl = P.cov_lbin[gr]; }
      }
      sm.rows[lr] = rs;
      sm.cnt[lr] = 0;
      sm.bp_lo[lr] = 0u; sm.bp_hi[lr] = 0u;
      sm.thr[lr] = F_TOPK ? ((P.gthr && seg >= 0) ? __ldcg(P.gthr + (int64_t)srt * TILE + lro + lr) : INFINITY) : -INFINITY;
    }
    for (int r = 0; r < RW; ++r) {
      sm.list_d[warp * RW + r][lane] = INFINITY;
      sm.list_j[warp * RW + r][lane] = INT_MAX;
    }
    __syncwarp();
    mbar_wait(&sm.a_full, items_done & 1);

    const float* a_base = &sm.a[0][warp * RW + ty][0];
    float4 av[RI];                         // A values of the upcoming k4 step (software pipeline, see below)
    #pragma unroll
    for (int i = 0; i < RI; ++i) av[i] = *reinterpret_cast<const float4*>(a_base + (2 * i) * KCH);

    uint32_t l = 0;                        // chunk index within the item
#ifdef DBB_K2_STATS
    long long t_main = 0, t_test = 0, t_drain = 0, t_mark = clock64();
#define DBB_LAP(var) { const long long t_now = clock64(); var += t_now - t_mark; t_mark = t_now; }
#else
#define DBB_LAP(var)
#endif
    float thr_pub = INFINITY;              // what this lane last published for its row (lanes < RW)
    for (int ts = t0; ts < t1; ++ts) {
      const int ct = tile_at(ts);
      // the bound the other segments of this row tile have proven by now: requested here, used after the K loop
      float thr_seen = INFINITY;
      if (F_TOPK && P.gthr && seg >= 0 && lane < RW) thr_seen = __ldcg(P.gthr + (int64_t)srt * TILE + lro + warp * RW + lane);
      float2 acc[RI][8];
      #pragma unroll
      for (int i = 0; i < RI; ++i)
        #pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);

      #pragma unroll 1
      for (int kc = 0; kc < NKC; ++kc, ++q, ++l) {
        const uint32_t st = q % NSTAGE;
        if (tid == 0 && l + PREFETCH < item_chunks) {
          const uint32_t ln = l + PREFETCH;
          const int tn = t0 + (int)(ln / NKC);
          issue_chunk<FLAGS, WARPS>(&tmap, P, qbase + ln, tile_at(tn), (int)(ln % NKC), tn & 1);
        }
        mbar_wait(&sm.full[st], (q / NSTAGE) & 1);
        const float* bq = &sm.b[st][tx][0];
        // The A values of the NEXT k4 step are requested as soon as the current ones have seen their last use
        // (column j = 7): the A tile is resident for the whole sweep, so the prefetch simply runs on across K
        // chunks and column tiles and the step never starts by waiting for eight LDS.128.
        // Not unrolled: one k4 step is ~280 instructions (4.4 KB); unrolling all 7 steps of a chunk (31 KB) cost
        // ~25 % in instruction-fetch stalls, unrolling by two was 3 % slower (profiles/r01_pairs_notes.md).
        // The loop start is a run-time zero on purpose: with a compile-time trip count ptxas carries the
        // prefetched A registers through ~25 MOVs per step (measured -6 %, profiles/r01_pairs_notes.md).
        #pragma unroll 1
        for (int k4 = P.opaque_zero; k4 < KCH / 4; ++k4) {
          const int nkc = (k4 + 1 < KCH / 4) ? kc : (kc + 1 < NKC ? kc + 1 : 0);
          const int nk4 = (k4 + 1 < KCH / 4) ? k4 + 1 : 0;
          const float* an = a_base + nkc * (TROWS * KCH) + nk4 * 4;
          #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(bq + (16 * j) * KCH + k4 * 4);
            // eight independent subtractions, then the eight dependent accumulations: keeps every dependent
            // FADD2 pair >= 8 issue slots apart (a lone warp otherwise stalls on the 4-cycle FADD2 latency)
            float2 dd[RI];
            #pragma unroll
            for (int i = 0; i < RI; ++i) dd[i] = __fadd2_rn(make_float2(av[i].x, av[i].y), make_float2(-bv.x, -bv.y));
            #pragma unroll
            for (int i = 0; i < RI; ++i)
              acc[i][j] = F_L2 ? __ffma2_rn(dd[i], dd[i], acc[i][j]) : __fadd2_rn(acc[i][j], make_float2(fabsf(dd[i].x), fabsf(dd[i].y)));
            #pragma unroll
            for (int i = 0; i < RI; ++i) dd[i] = __fadd2_rn(make_float2(av[i].z, av[i].w), make_float2(-bv.z, -bv.w));
            #pragma unroll
            for (int i = 0; i < RI; ++i) {
              acc[i][j] = F_L2 ? __ffma2_rn(dd[i], dd[i], acc[i][j]) : __fadd2_rn(acc[i][j], make_float2(fabsf(dd[i].x), fabsf(dd[i].y)));
              if (j == 7) av[i] = *reinterpret_cast<const float4*>(an + (2 * i) * KCH);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
      }

      DBB_LAP(t_main)
      // ---------------- epilogue of this column tile ----------------
      const int64_t col_base = (int64_t)ct * TILE;
      const ColScalars& cs = sm.cols[ts & 1];
      if (P.debug_skip_epilogue == 1) {
        float keep = 0.f;
        #pragma unroll
        for (int i = 0; i < RI; ++i)
          #pragma unroll
          for (int j = 0; j < 8; ++j) keep += acc[i][j].x + acc[i][j].y;
        if (keep == -1.f) sm.thr[0] = keep;
      } else if (always_slow) {
        // every pair matters (masks, or counts without a TD bound): evaluate all of them
        #pragma unroll
        for (int i = 0; i < RI; ++i) {
          #pragma unroll
          for (int j = 0; j < 8; ++j) sm.q[warp * QCAP + (j * 32 + lane)].x = fin(acc[i][j]);
          __syncwarp();
          process_row_pair<FLAGS, WARPS>(P, ts & 1, warp, lane, warp * RW + 2 * i, row_base, col_base);
        }
      } else {
        // steady state: one out-of-line call per row pair (see test_row_pair); a row pair pushes at most 256
        // entries, so checking the fill after every second one needs 512 free slots
        if (lane == 0) sm.qn[warp] = 0;
        __syncwarp();
        // (the first-tile threshold assumes that any column may enter the list: not with a filtered top-k, where the
        // columns below it may all fail the filter -- then the first tile is taken whole, like any tile before the
        // list is full)
        const int tflags = ((F_TOPK && ts == t0 && P.topk_filter == 0) ? 1 : 0) | (want_cnt ? 2 : 0);
        const int ncol = (int)((n - col_base) < (int64_t)TILE ? (n - col_base) : (int64_t)TILE);
        #pragma unroll
        for (int i = 0; i < RI; ++i) {
          test_row_pair<FLAGS, WARPS>(ts & 1, warp, lane, i, tflags, ncol, P.k,
                                      fin(acc[i][0]), fin(acc[i][1]), fin(acc[i][2]), fin(acc[i][3]), fin(acc[i][4]),
                                      fin(acc[i][5]), fin(acc[i][6]), fin(acc[i][7]));
          if ((i & 1) && i + 1 < RI) {
            __syncwarp();
            const int qn = sm.qn[warp];
            if (qn > QCAP - 512) {
              drain_queue<FLAGS, WARPS>(P, ts & 1, warp, lane, qn, row_base, col_base);
              if (lane == 0) sm.qn[warp] = 0;
              __syncwarp();
            }
          }
        }
        __syncwarp();
        const int qn = sm.qn[warp];
        DBB_LAP(t_test)
        if (qn && P.debug_skip_epilogue != 2) drain_queue<FLAGS, WARPS>(P, ts & 1, warp, lane, qn, row_base, col_base);
      }
      DBB_LAP(t_drain)
      if (F_TOPK && P.gthr && seg >= 0) {
        // exchange with the other segments of this row tile: publish this list's k-th best (once it has one) and pick
        // up the smallest one published so far -- every full list proves an upper bound of the row's final k-th best,
        // so candidates above it cannot be in the merged top-k (ties pass: the tests are <=)
        // (the load was issued before the K loop and the publication returns nothing, so neither waits for L2)
        __syncwarp();
        if (lane < RW) {
          const int lr = warp * RW + lane;
          const float loc = sm.thr[lr];
          if (loc < thr_pub) {
            atomicMin(reinterpret_cast<unsigned int*>(P.gthr + (int64_t)srt * TILE + lro + lr), __float_as_uint(loc));
            thr_pub = loc;
          }
          sm.thr[lr] = fminf(loc, thr_seen);
        }
        __syncwarp();
      }
    }  // column tiles
#ifdef DBB_K2_STATS
    if (P.stats && lane == 0) {
      atomicAdd(P.stats + 3, (unsigned long long)t_main);
      atomicAdd(P.stats + 4, (unsigned long long)t_test);
      atomicAdd(P.stats + 5, (unsigned long long)t_drain);
    }
#endif

    // ---------------- per-row outputs (final for whole sweeps, partial for column segments) -----------
    __syncwarp();
    const int k = P.k;
    if (lane < RW) {
      const int lr = warp * RW + lane;
      const int64_t gi = row_base + lr;
      if (gi < P.row1) {
        const int64_t bp = (int64_t)(((unsigned long long)sm.bp_hi[lr] << 32) | sm.bp_lo[lr]);
        if (seg < 0) {
          if (P.count) P.count[gi - P.row0] = sm.cnt[lr];
          if (P.neighbour_bp) P.neighbour_bp[gi - P.row0] = bp;
        } else {
          const int64_t o = ((int64_t)srt * TILE + lro + lr) * P.n_seg + seg;
          P.part_cnt[o] = sm.cnt[lr];
          P.part_bp[o] = bp;
        }
      }
    }
    if (F_TOPK) {
      for (int r = 0; r < RW; ++r) {
        const int lr = warp * RW + r;
        const int64_t gi = row_base + lr;
        if (gi < P.row1 && lane < k) {
          const int32_t ej = sm.list_j[lr][lane];
          const float ed = sm.list_d[lr][lane];
          if (seg < 0) {
            if (P.knn_idx) P.knn_idx[(gi - P.row0) * k + lane] = (ej == INT_MAX) ? -1 : id_of(P.col_id, ej);
            if (P.knn_dist) P.knn_dist[(gi - P.row0) * k + lane] = ed;
          } else {
            const int64_t o = (((int64_t)srt * TILE + lro + lr) * P.n_seg + seg) * k + lane;
            P.part_idx[o] = ej;
            P.part_dist[o] = ed;
          }
        }
      }
    }
    ++items_done;
  }  // items
}

// Merge of the partial results of split row tiles: one warp per row, lane l holds entry l of the sorted
// list; the (distance, index) order is total, so the merged list is the same as an unsplit sweep's.
__global__ void __launch_bounds__(128) merge_kernel(const Params P, int64_t n_split_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);   // split row tile r / TILE, row r % TILE of it
  if (r >= n_split_rows) return;
  const int pos = P.n_full_rt + (int)(r / TILE);
  const int rt = P.item_order ? P.item_order[pos] : pos;
  const int64_t gi = (int64_t)rt * TILE + (r % TILE);               // row index relative to row0
  if (gi >= P.row1 - P.row0) return;
  const int k = P.k;
  int32_t cnt = 0;
  int64_t bp = 0;
  for (int sgi = 0; sgi < P.n_seg; ++sgi) { cnt += P.part_cnt[r * P.n_seg + sgi]; bp += P.part_bp[r * P.n_seg + sgi]; }
  if (lane == 0) {
    if (P.count) P.count[gi] = cnt;
    if (P.neighbour_bp) P.neighbour_bp[gi] = bp;
  }
  if (k > 0) {
    float ed = INFINITY;
    int32_t ej = INT_MAX;
    for (int sgi = 0; sgi < P.n_seg; ++sgi) {
      // the whole partial list in one coalesced load (lane e holds entry e), then walked through shuffles: no memory
      // latency inside the insertion loop
      const int64_t o = (r * P.n_seg + sgi) * k;
      const float sd = lane < k ? P.part_dist[o + lane] : INFINITY;
      const int32_t sj = lane < k ? P.part_idx[o + lane] : INT_MAX;
      for (int e = 0; e < k; ++e) {
        const float d = __shfl_sync(0xffffffffu, sd, e);
        const int32_t c = __shfl_sync(0xffffffffu, sj, e);
        if (c == INT_MAX) break;              // lists are sorted: the rest is padding
        // ... and so is everything behind an entry that is farther than the current k-th best (ties still compare ids)
        if (d > __shfl_sync(0xffffffffu, ed, k - 1)) break;
        bool before = ed < d;
        if (ed == d) before = id_of(P.col_id, ej) < id_of(P.col_id, c);
        const int pos = __popc(__ballot_sync(0xffffffffu, before));
        const float pd = __shfl_up_sync(0xffffffffu, ed, 1);
        const int32_t pj = __shfl_up_sync(0xffffffffu, ej, 1);
        if (pos < k) {
          if (lane == pos) { ed = d; ej = c; }
          else if (lane > pos) { ed = pd; ej = pj; }
        }
      }
    }
    if (lane < k) {
      if (P.knn_idx) P.knn_idx[gi * k + lane] = (ej == INT_MAX) ? -1 : id_of(P.col_id, ej);
      if (P.knn_dist) P.knn_dist[gi * k + lane] = ed;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DBB_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) {
      dbb::set_error("cuTensorMapEncodeTiled is not available in this driver");
      return DBB_ERR_CUDA;
    }
    cached = (EncodeTiledFn)p;
  }
  *fn = cached;
  return DBB_OK;
}

// Warp count of the instance to launch: 4 (two 64-row CTAs per SM) or 8 (one 128-row CTA per SM).  The two small CTAs
// fetch every column tile twice, which is free while the signature matrix sits in L2.  Measured 4 vs 8 warps: C2 (28 MB of
// signatures) 10.49 vs 10.77 ms; C3 (140 MB) 241.6 vs 241.5 ms; C4 (560 MB) 3 456 vs 3 460 ms -- a gain where the matrix
// is L2-resident and the items are short, none beyond.  So: 4 warps up to 64 MB of signatures, 8 (the instance with half
// the L2 traffic) beyond; DBB_K2_WARPS=4|8 forces one for A/B runs.  -DDBB_K2_WARPS16 builds the 16-warp instance instead.
int pairs_warps(int64_t n) {
#ifdef DBB_K2_WARPS16
  return 16;
#else
  static const int forced = [] { const char* e = getenv("DBB_K2_WARPS"); const int w = (e && *e) ? atoi(e) : 0; return (w == 4 || w == 8) ? w : 0; }();
  if (forced) return forced;
  return n * DBB_SIG_STRIDE * 4 <= (64ll << 20) ? 4 : 8;
#endif
}

template <int FLAGS, int WARPS>
int launch_inst(int sms, int64_t n_items, cudaStream_t st, const CUtensorMap& tmap_a, const CUtensorMap& tmap, const Params& P) {
  constexpr size_t smem = sizeof(typename Cfg<WARPS>::Smem);
  constexpr int per_item = TILE / Cfg<WARPS>::TROWS;
  DBB_CUDA_TRY(cudaFuncSetAttribute(pairs_kernel<FLAGS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)std::min<int64_t>(n_items * per_item, (int64_t)sms * Cfg<WARPS>::MINB);
  pairs_kernel<FLAGS, WARPS><<<grid, WARPS * 32, smem, st>>>(tmap_a, tmap, P);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

template <int FLAGS>
int launch_one(int warps, int sms, int64_t n_items, cudaStream_t st, const CUtensorMap& tmap_a, const CUtensorMap& tmap, const Params& P) {
  // 16 warps (4x8 register tiles, 128 registers, four warps per sub-partition) hide the epilogue better than 8 (2.1 ms
  // instead of 3.6 ms on C2's full sweep) but pay 50 % more LDS per FADD2 in the distance loop: a tie
  // (profiles/r01_pairs_notes.md).
#ifdef DBB_K2_WARPS16
  return launch_inst<FLAGS, 16>(sms, n_items, st, tmap_a, tmap, P);
#else
  if (warps == 8) return launch_inst<FLAGS, 8>(sms, n_items, st, tmap_a, tmap, P);
  return launch_inst<FLAGS, 4>(sms, n_items, st, tmap_a, tmap, P);
#endif
}

int launch_pairs(int flags, int warps, int sms, int64_t n_items, cudaStream_t st, const CUtensorMap& tmap_a, const CUtensorMap& tmap, const Params& P) {
  switch (flags) {
#define DBB_CASE(F) case F: return launch_one<F>(warps, sms, n_items, st, tmap_a, tmap, P);
    DBB_CASE(0) DBB_CASE(1) DBB_CASE(2) DBB_CASE(3) DBB_CASE(4) DBB_CASE(5) DBB_CASE(6) DBB_CASE(7)
    DBB_CASE(8) DBB_CASE(9) DBB_CASE(10) DBB_CASE(11) DBB_CASE(12) DBB_CASE(13) DBB_CASE(14) DBB_CASE(15)
    DBB_CASE(24) DBB_CASE(25)     // Euclidean: k-NN, optionally with the per-scaffold distance bound and counts
#undef DBB_CASE
  }
  dbb::set_error("bad pairs flags");
  return DBB_ERR_INVALID;
}

// Wave quantisation: with one CTA per SM, n_rt row sweeps take ceil(n_rt / SMs) rounds.  The row tiles of the
// last, partial round are split into n_seg column segments so that the leftover work spreads over all SMs;
// items are handed out dynamically, whole sweeps first (longest-processing-time order).
void plan_items(int64_t n_rt, int64_t n_ct, int sms, int64_t* n_full_out, int64_t* rem_out, int* n_seg_out) {
  int64_t n_full = (n_rt / sms) * sms, rem = n_rt - n_full;
  int n_seg = 1;
  if (rem > 0) {
    double best = 1e30;
    const int max_seg = (int)std::min<int64_t>(8, std::max<int64_t>(1, n_ct / 8));   // keep segments >= 8 column tiles
    for (int sg = 1; sg <= max_seg; ++sg) {
      const double rounds = (double)((rem * sg + sms - 1) / sms) / sg + 0.02 * (sg - 1);   // + flood/merge overhead per split
      if (rounds < best - 1e-9) { best = rounds; n_seg = sg; }
    }
    const char* env_seg = getenv("DBB_K2_SEG");      // tests force a split on small inputs
    if (env_seg && *env_seg) n_seg = std::max(1, std::min(atoi(env_seg), (int)std::max<int64_t>(1, n_ct)));
  }
  if (n_seg == 1) { n_full = n_rt; rem = 0; }
  *n_full_out = n_full; *rem_out = rem; *n_seg_out = n_seg;
}

}  // namespace

DBB_EXPORT int64_t dbb_pairs_scratch_bytes(int64_t n, int64_t rows, int32_t k) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t n_full, rem; int n_seg;
  plan_items((rows + TILE - 1) / TILE, (n + TILE - 1) / TILE, sms, &n_full, &rem, &n_seg);
  return 8192 + rem * TILE * n_seg * ((int64_t)std::max(k, 1) * 8 + 16) + rem * TILE * 4;
}

DBB_EXPORT int dbb_pairs_launches(int64_t n, int64_t rows) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t n_full, rem; int n_seg;
  plan_items((rows + TILE - 1) / TILE, (n + TILE - 1) / TILE, sms, &n_full, &rem, &n_seg);
  return rem > 0 ? 2 : 1;
}

DBB_EXPORT int dbb_pairs_dev(const dbb_pairs_input* in, int64_t row0, int64_t row1, const dbb_pairs_output* out,
                             void* stream) {
  DBB_REQUIRE(in && out, "NULL argument struct");
  DBB_REQUIRE(in->n >= 0 && row0 >= 0 && row1 >= row0 && row1 <= in->n, "bad row range");
  if (row1 == row0 || in->n == 0) return DBB_OK;
  DBB_REQUIRE(in->n < (1ll << 31) - TILE, "n too large for int32 indices");
  DBB_REQUIRE(in->sig && in->length, "sig/length are required");
  DBB_REQUIRE(((uintptr_t)in->sig & 15) == 0, "sig must be 16-byte aligned");
  // the per-scaffold arrays are staged tile by tile with 16-byte-granular bulk copies
  DBB_REQUIRE((((uintptr_t)in->length | (uintptr_t)in->td_bound | (uintptr_t)in->gc | (uintptr_t)in->gc_mbin |
                (uintptr_t)in->gc_lbin | (uintptr_t)in->cov | (uintptr_t)in->cov_lbin) & 15) == 0,
              "per-scaffold arrays must be 16-byte aligned");
  DBB_REQUIRE(out->k >= 0 && out->k <= DBB_MAX_K, "k out of range (0..32)");
  DBB_REQUIRE(in->metric == DBB_METRIC_L1 || in->metric == DBB_METRIC_L2, "metric must be DBB_METRIC_L1 or DBB_METRIC_L2");
  DBB_REQUIRE(in->metric == DBB_METRIC_L1 || (out->k > 0 && !in->gc && !in->cov && !out->td_mask && !out->gc_mask && !out->cov_mask),
              "the Euclidean metric supports k-NN (k > 0) with the optional distance bound and counts only");
  DBB_REQUIRE(!in->gc || (in->gc_mbin && in->gc_lbin && in->gc_lo && in->gc_hi && in->gc_nl > 0), "incomplete GC inputs");
  DBB_REQUIRE(!in->cov || (in->cov_lbin && in->cov_lo && in->cov_hi), "incomplete coverage inputs");
  const bool masks = out->td_mask || out->gc_mask || out->cov_mask;
  DBB_REQUIRE(!masks || out->mask_words * 32 >= in->n, "mask_words too small");
  DBB_REQUIRE(!out->td_mask || in->td_bound, "td_mask requested without td_bound");
  DBB_REQUIRE((in->tile_off == nullptr) == (in->tile_list == nullptr), "tile_off and tile_list go together");
  DBB_REQUIRE(!in->tile_off || !masks, "masks need every column tile: not available with a sparse sweep (tile_off / tile_list)");
  DBB_REQUIRE(in->split_all >= 0 && in->split_all <= 64 && in->n_active_row_tiles >= 0 && in->split_tiles >= 0, "bad split_all / n_active_row_tiles / split_tiles");
  DBB_REQUIRE(in->n_active_row_tiles == 0 || in->item_order, "n_active_row_tiles needs item_order");

  EncodeTiledFn encode;
  int rc = get_encode_fn(&encode);
  if (rc != DBB_OK) return rc;
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)DBB_SIG_STRIDE, (cuuint64_t)in->n};
  cuuint64_t gstr[1] = {(cuuint64_t)DBB_SIG_STRIDE * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)in->sig, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { dbb::set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return DBB_ERR_CUDA; }
  // the row (A) tile of a CTA: the same matrix with a box of the instance's tile rows
  CUtensorMap tmap_a = tmap;
  const int warps = pairs_warps(in->n);
  if (warps == 4) {
    cuuint32_t box_a[2] = {(cuuint32_t)KCH, (cuuint32_t)Cfg<4>::TROWS};
    cr = encode(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)in->sig, gdim, gstr, box_a, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { dbb::set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return DBB_ERR_CUDA; }
  }

  Params P;
  P.n = in->n; P.row0 = row0; P.row1 = row1;
  P.length = in->length; P.td_bound = in->td_bound;
  P.gc = in->gc; P.gc_mbin = in->gc_mbin; P.gc_lbin = in->gc_lbin; P.gc_lo = in->gc_lo; P.gc_hi = in->gc_hi; P.gc_nl = in->gc_nl;
  P.cov = in->cov; P.cov_lbin = in->cov_lbin; P.cov_lo = in->cov_lo; P.cov_hi = in->cov_hi;
  P.k = out->k; P.topk_filter = out->topk_filter;
  P.col_id = in->col_id; P.item_order = in->item_order; P.tile_off = in->tile_off; P.tile_list = in->tile_list;
  P.knn_idx = out->knn_idx; P.knn_dist = out->knn_dist; P.count = out->count; P.neighbour_bp = out->neighbour_bp;
  P.td_mask = out->td_mask; P.gc_mask = out->gc_mask; P.cov_mask = out->cov_mask; P.mask_words = out->mask_words;

  int dev = 0, sms = 0;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  DBB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = row1 - row0;
  const int64_t n_rt = (n_rows + TILE - 1) / TILE;
  const int64_t n_ct = (in->n + TILE - 1) / TILE;

  int64_t n_full, rem;
  int n_seg;
  const int64_t n_rt_active = in->n_active_row_tiles > 0 ? std::min<int64_t>(in->n_active_row_tiles, n_rt) : n_rt;
  if (in->split_all > 0) { n_full = 0; rem = n_rt_active; n_seg = in->split_all; }     // caller-planned sparse sweep
  else plan_items(n_rt_active, n_ct, sms, &n_full, &rem, &n_seg);
  P.n_full_rt = (int32_t)n_full;
  P.n_seg = n_seg;
  P.seg_tiles = (in->split_all > 0 && in->tile_list) ? in->split_tiles : 0;
  P.n_items = (int32_t)(n_full + rem * n_seg);
  P.opaque_zero = 0;
  { const char* e2 = getenv("DBB_K2_DEBUG_SKIP_EPILOGUE"); P.debug_skip_epilogue = (e2 && *e2) ? atoi(e2) : 0; }   // profiling only
  const int64_t split_rows = rem * TILE;          // merge_kernel skips the rows of a last, partial row tile itself
  const size_t k_ = (size_t)std::max(out->k, 1);
  size_t off_idx = 8192, off_dist, off_cnt, off_bp, total;          // header: item counter, statistics (DBB_K2_STATS)
  off_dist = off_idx + (size_t)rem * TILE * n_seg * k_ * 4;
  off_cnt = off_dist + (size_t)rem * TILE * n_seg * k_ * 4;
  off_bp = (off_cnt + (size_t)rem * TILE * n_seg * 4 + 7) & ~(size_t)7;
  const size_t off_gthr = off_bp + (size_t)rem * TILE * n_seg * 8;
  const bool share_thr = rem > 0 && n_seg > 1 && out->k > 0;
  total = off_gthr + (share_thr ? (size_t)rem * TILE * 4 : 0);
  // stream-ordered scratch; keep freed blocks cached in the device pool instead of returning them to the
  // driver at every synchronisation (the default release threshold is 0)
  static bool pool_configured[64] = {false};       // per device ordinal
  if (dev >= 0 && dev < 64 && !pool_configured[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_configured[dev] = true;
  }
  uint8_t* scratch = nullptr;
  DBB_CUDA_TRY(cudaMallocAsync((void**)&scratch, total, st));
  DBB_CUDA_TRY(cudaMemsetAsync(scratch, 0, 8192, st));
  P.item_counter = (uint32_t*)scratch;
  static const bool want_stats = [] { const char* e = getenv("DBB_K2_STATS"); return e && *e == '1'; }();
  P.stats = want_stats ? (unsigned long long*)(scratch + 64) : nullptr;
  P.part_idx = (int32_t*)(scratch + off_idx);
  P.part_dist = (float*)(scratch + off_dist);
  P.part_cnt = (int32_t*)(scratch + off_cnt);
  P.part_bp = (int64_t*)(scratch + off_bp);
  P.gthr = nullptr;
  if (share_thr) {
    P.gthr = (float*)(scratch + off_gthr);
    DBB_CUDA_TRY(cudaMemsetAsync(P.gthr, 0x7f, (size_t)rem * TILE * 4, st));      // 0x7f7f7f7f = 3.4e38: "no bound yet"
  }

  const unsigned grid = (unsigned)std::min<int64_t>((int64_t)P.n_items * (warps == 4 ? 2 : 1), (int64_t)sms * (warps == 4 ? 2 : 1));
  int flags = (in->td_bound ? 1 : 0) | ((in->gc || in->cov) ? 2 : 0) | (masks ? 4 : 0) | (out->k > 0 ? 8 : 0);
  if (in->metric == DBB_METRIC_L2) flags |= 16;
  rc = launch_pairs(flags, warps, sms, P.n_items, st, tmap_a, tmap, P);
  if (rc == DBB_OK && rem > 0 && split_rows > 0) {
    merge_kernel<<<(unsigned)((split_rows + 3) / 4), 128, 0, st>>>(P, split_rows);
    if (cudaGetLastError() != cudaSuccess) { dbb::set_error("merge kernel launch failed"); rc = DBB_ERR_CUDA; }
  }
  if (want_stats && rc == DBB_OK) {
    static unsigned long long h[1024];
    cudaMemcpyAsync(h, scratch + 64, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    const double pairs = (double)(row1 - row0) * (double)in->n;
    const double tt = (double)(h[3] + h[4] + h[5]) + 1e-9;
    fprintf(stderr, "[dbb K2 stats] pairs %.3e queued %llu (%.4f %%) drains %llu top-k candidates %llu | warp cycles: distance loop "
            "%.1f %% hot test %.1f %% final drain %.1f %%\n", pairs, h[0], 100.0 * (double)h[0] / pairs, h[1], h[2],
            100.0 * h[3] / tt, 100.0 * h[4] / tt, 100.0 * h[5] / tt);
    // when each CTA ran out of work, relative to the start of CTA 0 (us): the sweep ends with the last one
    double mn = 1e30, mx = 0, sum = 0;
    for (unsigned b = 0; b < grid && b < 1000; ++b) {
      const double e = (double)(h[16 + b] - h[8]) * 1e-3;
      mn = std::min(mn, e); mx = std::max(mx, e); sum += e;
    }
    fprintf(stderr, "[dbb K2 stats] items %d over %u CTAs: CTA finish times min %.0f mean %.0f max %.0f us (max / mean %.3f)\n",
            P.n_items, grid, mn, sum / grid, mx, mx / (sum / grid));
  }
  cudaFreeAsync(scratch, st);
  return rc;
}
