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
// Kernel shape (B200): one persistent CTA per SM, 8 warps (thread 0 also issues the TMA loads).
//   * CTA tile 128 rows x 128 columns; warp w owns rows 16w..16w+15 against all 128 columns, lane
//     (ty = lane/16, tx = lane%16) owns rows 16w+2i+ty and columns tx+16j (i, j < 8): an 8x8 register
//     tile, accumulated as float2 so the inner loop is two packed FADD2 per two dimensions
//     (d = a - b ; acc += |d|) -- one FP32 issue slot per pair-dimension instead of two.
//   * the signature matrix is [n][140] float32 (136 + 4 zero pad = 5 K-chunks of 28 floats); TMA
//     (cp.async.bulk.tensor.2d) stages [128 rows][28 floats] boxes: the A row tile is resident for a
//     whole row sweep (5 boxes), B boxes stream through a 4-stage mbarrier ring.  Row pitch 112 B =
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
constexpr int NSTAGE = 4;
constexpr int CHUNK_BYTES = TILE * KCH * 4;  // 14336
constexpr int MAX_WARPS = 16;                // the kernel is instantiated for 8 warps (8x8 register tiles, <= 255
                                             // registers) and 16 warps (4x8 tiles, <= 128 registers, 4 warps per sub-partition)
constexpr int PREFETCH = 2;                  // B chunks in flight ahead of the consumers
constexpr int LIST_K = DBB_MAX_K;            // list slots per row
constexpr int QTOTAL = 5120;                 // queue entries per CTA; per warp QTOTAL / WARPS (>= 256 + one row pair's pushes)

struct RowScalars {       // staged in shared memory per row tile
  int32_t len;
  float tdb;
  int32_t gcm, gcl, covl;
  int32_t pad;
  double gc, cov;
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
  uint32_t* item_counter;
  int32_t* part_idx; float* part_dist; int32_t* part_cnt; int64_t* part_bp;
};

struct SmemLayout {
  alignas(128) float a[NKC][TILE][KCH];
  alignas(128) float b[NSTAGE][TILE][KCH];
  float list_d[TILE][LIST_K];
  int32_t list_j[TILE][LIST_K];
  float qd[QTOTAL];                       // per-warp queue of pairs that may be a neighbour or a top-k candidate
                                         //   (the every-pair path reuses the first 256 entries as its scratch)
  uint32_t qcode[QTOTAL];                //   (row within the warp's rows) << 7 | (column within the tile)
  RowScalars rows[TILE];
  uint32_t bp_lo[TILE], bp_hi[TILE];     // neighbourhood base pairs per row (64-bit as two native 32-bit atomics)
  int32_t cnt[TILE];                     // neighbourhood size per row
  float thr[TILE];                       // current k-th best distance per row (+inf until k found)
  int32_t qn[MAX_WARPS];                 // queue fill per warp
  float ctdb[MAX_WARPS][TILE];           // per-warp copy of the current column tile's TD bounds and lengths:
  int32_t clen[MAX_WARPS][TILE];         //   requested before the FADD2 loop, so their latency hides behind it
  alignas(8) uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  uint64_t a_full;
  int32_t item;
};

static_assert(sizeof(SmemLayout) <= 232448, "SmemLayout exceeds the 227 KB per-CTA shared memory limit");

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

__device__ __forceinline__ float2 absdiff_acc(float2 acc, float2 a, float2 b) {
  float2 d = __fadd2_rn(a, make_float2(-b.x, -b.y));
  return __fadd2_rn(acc, make_float2(fabsf(d.x), fabsf(d.y)));
}

__device__ __forceinline__ void add_bp(SmemLayout& sm, int lr, uint32_t v) {
  const uint32_t old = atomicAdd(&sm.bp_lo[lr], v);
  if (old + v < old) atomicAdd(&sm.bp_hi[lr], 1u);
}

// warp-cooperative insertion of (d, c) into the sorted list of `row`; returns the new k-th best
__device__ __forceinline__ float list_insert(SmemLayout& sm, int row, float d, int32_t c, int k, int lane) {
  float ed = sm.list_d[row][lane];
  int32_t ej = sm.list_j[row][lane];
  const bool before = (ed < d) || (ed == d && ej < c);
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
__device__ __noinline__ void process_row_pair(SmemLayout& sm, const Params& P, int warp, int lane, int lr0,
                                              int64_t row_base, int64_t col_base) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_MASK = (FLAGS & 4) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int QCAP = QTOTAL / WARPS;
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
    const float d = sm.qd[warp * QCAP + (j * 32 + lane)];
    const int64_t cj = col_base + tx + 16 * j;
    const bool col_ok = cj < n;
    int32_t len_j = 0;
    bool p_td = col_ok, p_gc = col_ok, p_cov = col_ok;
    if (col_ok) {
      len_j = __ldg(P.length + cj);
      const bool i_core = rs.len > len_j;
      if (F_TD) p_td = d < (i_core ? __ldg(P.td_bound + cj) : rs.tdb);
      if (has_gc) {
        const int m = i_core ? rs.gcm : __ldg(P.gc_mbin + cj), l = i_core ? __ldg(P.gc_lbin + cj) : rs.gcl;
        const double lo = __ldg(P.gc_lo + m * P.gc_nl + l), hi = __ldg(P.gc_hi + m * P.gc_nl + l);
        const double gj = __ldg(P.gc + cj);
        const double diff = i_core ? (gj - rs.gc) : (rs.gc - gj);
        p_gc = diff >= lo && diff <= hi;
      }
      if (has_cov) {
        const double vj = __ldg(P.cov + cj);
        const double cc = i_core ? rs.cov : vj, cu = i_core ? vj : rs.cov;
        const int l = i_core ? __ldg(P.cov_lbin + cj) : rs.covl;
        const double eff = cc > 0 ? cc : 0.1;
        const double diff = (cu - eff) * 100.0 / eff;
        p_cov = diff >= __ldg(P.cov_lo + l) && diff <= __ldg(P.cov_hi + l);
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
          const float nt = list_insert(sm, crow, cd, (int32_t)(col_base + (src & 15) + 16 * j), k, lane);
          if (lane == 0) sm.thr[crow] = nt;
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

// Steady-state slow path: the warp's queue holds (distance, row, column) of the few pairs of this column
// tile that passed the conservative in-register test.  Lane-parallel predicates and counts, then the
// (rare) top-k candidates are inserted one at a time, warp-cooperatively.
template <int FLAGS, int WARPS>
__device__ __noinline__ void drain_queue(SmemLayout& sm, const Params& P, int warp, int lane, int qn,
                                         int64_t row_base, int64_t col_base) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int QCAP = QTOTAL / WARPS, RW = TILE / WARPS;
  const int64_t n = P.n;
  const bool has_gc = F_GCCOV && P.gc != nullptr, has_cov = F_GCCOV && P.cov != nullptr;
  const bool want_cnt = P.count != nullptr || P.neighbour_bp != nullptr;
  const int k = P.k;
  #pragma unroll 1
  for (int base = 0; base < qn; base += 32) {
    const int e = base + lane;
    const bool valid = e < qn;
    const float d = valid ? sm.qd[warp * QCAP + (e)] : INFINITY;
    const uint32_t code = valid ? sm.qcode[warp * QCAP + (e)] : 0u;
    const int lr = warp * RW + (int)(code >> 7);
    const int64_t cj = col_base + (code & 127u);
    const int64_t gi = row_base + lr;
    const bool col_ok = valid && cj < n;
    int32_t len_j = 0;
    bool p_td = col_ok, p_gc = col_ok, p_cov = col_ok;
    if (col_ok) {
      const RowScalars rs = sm.rows[lr];
      len_j = sm.clen[warp][code & 127u];
      const bool i_core = rs.len > len_j;
      if (F_TD) p_td = d < (i_core ? sm.ctdb[warp][code & 127u] : rs.tdb);
      if (has_gc) {
        const int m = i_core ? rs.gcm : __ldg(P.gc_mbin + cj), l = i_core ? __ldg(P.gc_lbin + cj) : rs.gcl;
        const double lo = __ldg(P.gc_lo + m * P.gc_nl + l), hi = __ldg(P.gc_hi + m * P.gc_nl + l);
        const double gj = __ldg(P.gc + cj);
        const double diff = i_core ? (gj - rs.gc) : (rs.gc - gj);
        p_gc = diff >= lo && diff <= hi;
      }
      if (has_cov) {
        const double vj = __ldg(P.cov + cj);
        const double cc = i_core ? rs.cov : vj, cu = i_core ? vj : rs.cov;
        const int l = i_core ? __ldg(P.cov_lbin + cj) : rs.covl;
        const double eff = cc > 0 ? cc : 0.1;
        const double diff = (cu - eff) * 100.0 / eff;
        p_cov = diff >= __ldg(P.cov_lo + l) && diff <= __ldg(P.cov_hi + l);
      }
    }
    if (want_cnt && p_td && p_gc && p_cov) {
      atomicAdd(&sm.cnt[lr], 1);
      add_bp(sm, lr, (uint32_t)len_j);
    }
    if (F_TOPK) {
      bool cand = col_ok && (cj != gi) && (d <= sm.thr[lr]);
      if (P.topk_filter & 1) cand = cand && p_td;
      if (P.topk_filter & 2) cand = cand && p_gc;
      if (P.topk_filter & 4) cand = cand && p_cov;
      uint32_t m = __ballot_sync(0xffffffffu, cand);
      if (P.debug_skip_epilogue == 3) m = 0;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int crow = __shfl_sync(0xffffffffu, lr, src);
        const float cd = __shfl_sync(0xffffffffu, d, src);
        const int32_t cc = (int32_t)__shfl_sync(0xffffffffu, (int)cj, src);
        if (cd <= sm.thr[crow]) {
          const float nt = list_insert(sm, crow, cd, cc, k, lane);
          if (lane == 0) sm.thr[crow] = nt;
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

// FLAGS: bit0 TD predicate, bit1 GC and/or coverage predicates, bit2 masks, bit3 top-k
template <int FLAGS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pairs_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params P) {
  constexpr bool F_TD = (FLAGS & 1) != 0, F_GCCOV = (FLAGS & 2) != 0, F_MASK = (FLAGS & 4) != 0, F_TOPK = (FLAGS & 8) != 0;
  constexpr int RW = TILE / WARPS;          // rows per warp: 16 or 8
  constexpr int RI = RW / 2;                // rows per lane: 8 or 4 (lane owns rows warp*RW + 2i + ty)
  constexpr int QCAP = QTOTAL / WARPS;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
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

  // B chunk `l` of the current item (K-chunk l%5 of column tile ct0 + l/5) goes to ring stage (qbase+l)%NSTAGE
  auto issue_b = [&](uint32_t qg, int ctile, int kc) {
    const uint32_t st = qg % NSTAGE;
    if (qg >= NSTAGE) mbar_wait(&sm.empty[st], ((qg / NSTAGE) - 1) & 1);
    mbar_arrive_expect_tx(&sm.full[st], CHUNK_BYTES);
    tma_load_2d(&sm.b[st][0][0], &tmap, kc * KCH, ctile * TILE, &sm.full[st]);
  };

  while (true) {
    __syncthreads();                       // everyone is done with the previous item's A tile and lists
    if (tid == 0) sm.item = (int32_t)atomicAdd(P.item_counter, 1u);
    __syncthreads();
    const int item = sm.item;
    if (item >= P.n_items) break;
    int rt, seg = -1, ct0 = 0, ct1 = n_ct;
    if (item < P.n_full_rt) {
      rt = item;
    } else {
      const int t = item - P.n_full_rt;
      rt = P.n_full_rt + t / P.n_seg;
      seg = t % P.n_seg;
      ct0 = (int)(((int64_t)n_ct * seg) / P.n_seg);
      ct1 = (int)(((int64_t)n_ct * (seg + 1)) / P.n_seg);
    }
    const int64_t row_base = P.row0 + (int64_t)rt * TILE;
    const uint32_t item_chunks = (uint32_t)(ct1 - ct0) * NKC;
    const uint32_t qbase = q;

    // thread 0 doubles as the TMA producer: A tile of this item + its first PREFETCH B chunks
    if (tid == 0) {
      mbar_arrive_expect_tx(&sm.a_full, NKC * CHUNK_BYTES);
      for (int kc = 0; kc < NKC; ++kc) tma_load_2d(&sm.a[kc][0][0], &tmap, kc * KCH, (int32_t)row_base, &sm.a_full);
      for (uint32_t l = 0; l < (uint32_t)PREFETCH && l < item_chunks; ++l) issue_b(qbase + l, ct0 + (int)(l / NKC), (int)(l % NKC));
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
      sm.thr[lr] = F_TOPK ? INFINITY : -INFINITY;
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
    for (int ct = ct0; ct < ct1; ++ct) {
      float2 acc[RI][8];
      #pragma unroll
      for (int i = 0; i < RI; ++i)
        #pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
      // column scalars of this tile (4 columns per lane), consumed by the epilogue
      float pre_tdb[4];
      int32_t pre_len[4];
      #pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int64_t cj = (int64_t)ct * TILE + lane + 32 * qq;
        pre_tdb[qq] = (F_TD && cj < n) ? __ldg(P.td_bound + cj) : -INFINITY;
        pre_len[qq] = cj < n ? __ldg(P.length + cj) : 0;
      }

      #pragma unroll 1
      for (int kc = 0; kc < NKC; ++kc, ++q, ++l) {
        const uint32_t st = q % NSTAGE;
        if (tid == 0 && l + PREFETCH < item_chunks) {
          const uint32_t ln = l + PREFETCH;
          issue_b(qbase + ln, ct0 + (int)(ln / NKC), (int)(ln % NKC));
        }
        mbar_wait(&sm.full[st], (q / NSTAGE) & 1);
        const float* bq = &sm.b[st][tx][0];
        // The A values of the NEXT k4 step are requested as soon as the current ones have seen their last use
        // (column j = 7): the A tile is resident for the whole sweep, so the prefetch simply runs on across K
        // chunks and column tiles and the step never starts by waiting for eight LDS.128.
        // Not unrolled: one k4 step is ~280 instructions (4.4 KB); unrolling all 7 steps of a chunk (31 KB) cost
        // ~25 % in instruction-fetch stalls, unrolling by two was 3 % slower (profiles/r01_pairs_notes.md).
        #pragma unroll 1
        for (int k4 = (P.debug_skip_epilogue == 4 && warp >= 4) ? KCH / 4 : 0; k4 < KCH / 4; ++k4) {
          const int nkc = (k4 + 1 < KCH / 4) ? kc : (kc + 1 < NKC ? kc + 1 : 0);
          const int nk4 = (k4 + 1 < KCH / 4) ? k4 + 1 : 0;
          const float* an = a_base + nkc * (TILE * KCH) + nk4 * 4;
          #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(bq + (16 * j) * KCH + k4 * 4);
            // eight independent subtractions, then the eight dependent accumulations: keeps every dependent
            // FADD2 pair >= 8 issue slots apart (a lone warp otherwise stalls on the 4-cycle FADD2 latency)
            float2 dd[RI];
            #pragma unroll
            for (int i = 0; i < RI; ++i) dd[i] = __fadd2_rn(make_float2(av[i].x, av[i].y), make_float2(-bv.x, -bv.y));
            #pragma unroll
            for (int i = 0; i < RI; ++i) acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fabsf(dd[i].x), fabsf(dd[i].y)));
            #pragma unroll
            for (int i = 0; i < RI; ++i) dd[i] = __fadd2_rn(make_float2(av[i].z, av[i].w), make_float2(-bv.z, -bv.w));
            #pragma unroll
            for (int i = 0; i < RI; ++i) {
              acc[i][j] = __fadd2_rn(acc[i][j], make_float2(fabsf(dd[i].x), fabsf(dd[i].y)));
              if (j == 7) av[i] = *reinterpret_cast<const float4*>(an + (2 * i) * KCH);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
      }

      // ---------------- epilogue of this column tile ----------------
      const int64_t col_base = (int64_t)ct * TILE;
      #pragma unroll
      for (int qq = 0; qq < 4; ++qq) { sm.ctdb[warp][lane + 32 * qq] = pre_tdb[qq]; sm.clen[warp][lane + 32 * qq] = pre_len[qq]; }
      __syncwarp();
      if (P.debug_skip_epilogue == 1 || P.debug_skip_epilogue == 4) {
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
          for (int j = 0; j < 8; ++j) sm.qd[warp * QCAP + (j * 32 + lane)] = acc[i][j].x + acc[i][j].y;
          __syncwarp();
          process_row_pair<FLAGS, WARPS>(sm, P, warp, lane, warp * RW + 2 * i, row_base, col_base);
        }
      } else {
        // A pair is queued only if it can be a TD neighbour (d below the larger of the two bounds -- the exact
        // rule is re-applied in the drain) or can enter the row's top-k (d <= k-th best).  On the first tile of
        // a sweep the lists are empty; instead of flooding them with all 128 columns the threshold is the
        // largest of the 16 lanes' m-th smallest distances, m = ceil((k+1)/16): at least k+1 columns pass.
        const bool first_tile = F_TOPK && ct == ct0;
        float tdb_c[8];
        bool col_ok[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t cj = col_base + tx + 16 * j;
          col_ok[j] = cj < n;
          tdb_c[j] = sm.ctdb[warp][tx + 16 * j];
        }
        float tdb_cmax = tdb_c[0];
        #pragma unroll
        for (int j = 1; j < 8; ++j) tdb_cmax = fmaxf(tdb_cmax, tdb_c[j]);
        if (lane == 0) sm.qn[warp] = 0;
        __syncwarp();
        #pragma unroll
        for (int i = 0; i < RI; ++i) {
          const int lr = warp * RW + 2 * i + ty;
          float thr_i = F_TOPK ? sm.thr[lr] : -INFINITY;
          if (first_tile) {
            const int m = (P.k + 16) / 16;                       // 1, 2 or 3
            float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;   // the lane's three smallest valid distances
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
              float d = acc[i][j].x + acc[i][j].y;
              d = (col_ok[j] && d == d) ? d : INFINITY;
              const float t1 = fminf(m1, d), c1 = fmaxf(m1, d);
              const float t2 = fminf(m2, c1), c2 = fmaxf(m2, c1);
              m1 = t1; m2 = t2; m3 = fminf(m3, c2);
            }
            float tau = m == 1 ? m1 : (m == 2 ? m2 : m3);
            #pragma unroll
            for (int o = 1; o < 16; o <<= 1) tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, o));
            thr_i = tau;
          }
          const float u = fmaxf(F_TD ? sm.rows[lr].tdb : -INFINITY, thr_i);
          // cheap row test first: the smallest of the lane's 8 distances against the largest of its 8 bounds
          float dmin = INFINITY;
          #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = acc[i][j].x + acc[i][j].y;
            acc[i][j].x = d;
            dmin = fminf(dmin, d);
          }
          if (dmin <= fmaxf(u, tdb_cmax)) {
            uint32_t bits = 0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) bits |= (acc[i][j].x <= fmaxf(u, tdb_c[j])) ? (1u << j) : 0u;
            if (bits) {
              int slot = atomicAdd(&sm.qn[warp], __popc(bits));    // reserve queue slots (hits are rare)
              #pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (bits & (1u << j)) {
                  sm.qd[warp * QCAP + (slot)] = acc[i][j].x;
                  sm.qcode[warp * QCAP + (slot)] = (uint32_t)((2 * i + ty) << 7) | (uint32_t)(tx + 16 * j);
                  ++slot;
                }
              }
            }
          }
          __syncwarp();
          const int qn = sm.qn[warp];
          if (qn > QCAP - 256) {                                 // one row pair pushes at most 256 entries
            drain_queue<FLAGS, WARPS>(sm, P, warp, lane, qn, row_base, col_base);
            if (lane == 0) sm.qn[warp] = 0;
            __syncwarp();
          }
        }
        const int qn = sm.qn[warp];
        if (qn && P.debug_skip_epilogue != 2) drain_queue<FLAGS, WARPS>(sm, P, warp, lane, qn, row_base, col_base);
      }
    }  // column tiles

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
          const int64_t o = ((int64_t)(rt - P.n_full_rt) * TILE + lr) * P.n_seg + seg;
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
            if (P.knn_idx) P.knn_idx[(gi - P.row0) * k + lane] = (ej == INT_MAX) ? -1 : ej;
            if (P.knn_dist) P.knn_dist[(gi - P.row0) * k + lane] = ed;
          } else {
            const int64_t o = (((int64_t)(rt - P.n_full_rt) * TILE + lr) * P.n_seg + seg) * k + lane;
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
__global__ void __launch_bounds__(128) merge_kernel(const Params P, int64_t first_row, int64_t n_split_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n_split_rows) return;
  const int64_t gi = first_row + r;             // row index relative to row0
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
      for (int e = 0; e < k; ++e) {
        const int64_t o = (r * P.n_seg + sgi) * k + e;
        const float d = P.part_dist[o];
        const int32_t c = P.part_idx[o];
        if (c == INT_MAX) break;              // lists are sorted: the rest is padding
        const bool before = (ed < d) || (ed == d && ej < c);
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
      if (P.knn_idx) P.knn_idx[gi * k + lane] = (ej == INT_MAX) ? -1 : ej;
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

template <int FLAGS>
int launch_one(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tmap, const Params& P) {
  // The kernel is templated on the warp count.  16 warps (4x8 register tiles, 128 registers, four warps per
  // sub-partition) hide the epilogue better (2.1 ms instead of 3.6 ms on C2) but pay 50 % more LDS per FADD2 in the
  // distance loop (108.8 vs 116.9 G pairs/s without epilogue): the two tie at ~100 G pairs/s, so only the 8-warp
  // instance is built by default (-DDBB_K2_WARPS16 builds and selects the other; profiles/r01_pairs_notes.md).
#ifdef DBB_K2_WARPS16
  DBB_CUDA_TRY(cudaFuncSetAttribute(pairs_kernel<FLAGS, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pairs_kernel<FLAGS, 16><<<grid, 512, smem, st>>>(tmap, P);
#else
  DBB_CUDA_TRY(cudaFuncSetAttribute(pairs_kernel<FLAGS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pairs_kernel<FLAGS, 8><<<grid, 256, smem, st>>>(tmap, P);
#endif
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

int launch_pairs(int flags, unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& tmap, const Params& P) {
  switch (flags) {
#define DBB_CASE(F) case F: return launch_one<F>(grid, smem, st, tmap, P);
    DBB_CASE(0) DBB_CASE(1) DBB_CASE(2) DBB_CASE(3) DBB_CASE(4) DBB_CASE(5) DBB_CASE(6) DBB_CASE(7)
    DBB_CASE(8) DBB_CASE(9) DBB_CASE(10) DBB_CASE(11) DBB_CASE(12) DBB_CASE(13) DBB_CASE(14) DBB_CASE(15)
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
  return 256 + rem * TILE * n_seg * ((int64_t)std::max(k, 1) * 8 + 16);
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
  DBB_REQUIRE(out->k >= 0 && out->k <= DBB_MAX_K, "k out of range (0..32)");
  DBB_REQUIRE(!in->gc || (in->gc_mbin && in->gc_lbin && in->gc_lo && in->gc_hi && in->gc_nl > 0), "incomplete GC inputs");
  DBB_REQUIRE(!in->cov || (in->cov_lbin && in->cov_lo && in->cov_hi), "incomplete coverage inputs");
  const bool masks = out->td_mask || out->gc_mask || out->cov_mask;
  DBB_REQUIRE(!masks || out->mask_words * 32 >= in->n, "mask_words too small");
  DBB_REQUIRE(!out->td_mask || in->td_bound, "td_mask requested without td_bound");

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

  Params P;
  P.n = in->n; P.row0 = row0; P.row1 = row1;
  P.length = in->length; P.td_bound = in->td_bound;
  P.gc = in->gc; P.gc_mbin = in->gc_mbin; P.gc_lbin = in->gc_lbin; P.gc_lo = in->gc_lo; P.gc_hi = in->gc_hi; P.gc_nl = in->gc_nl;
  P.cov = in->cov; P.cov_lbin = in->cov_lbin; P.cov_lo = in->cov_lo; P.cov_hi = in->cov_hi;
  P.k = out->k; P.topk_filter = out->topk_filter;
  P.knn_idx = out->knn_idx; P.knn_dist = out->knn_dist; P.count = out->count; P.neighbour_bp = out->neighbour_bp;
  P.td_mask = out->td_mask; P.gc_mask = out->gc_mask; P.cov_mask = out->cov_mask; P.mask_words = out->mask_words;

  int dev = 0, sms = 0;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  DBB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(SmemLayout);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = row1 - row0;
  const int64_t n_rt = (n_rows + TILE - 1) / TILE;
  const int64_t n_ct = (in->n + TILE - 1) / TILE;

  int64_t n_full, rem;
  int n_seg;
  plan_items(n_rt, n_ct, sms, &n_full, &rem, &n_seg);
  P.n_full_rt = (int32_t)n_full;
  P.n_seg = n_seg;
  P.n_items = (int32_t)(n_full + rem * n_seg);
  { const char* e2 = getenv("DBB_K2_DEBUG_SKIP_EPILOGUE"); P.debug_skip_epilogue = (e2 && *e2) ? atoi(e2) : 0; }   // profiling only
  const int64_t split_rows = std::min<int64_t>(rem * TILE, n_rows - n_full * TILE);
  const size_t k_ = (size_t)std::max(out->k, 1);
  size_t off_idx = 256, off_dist, off_cnt, off_bp, total;
  off_dist = off_idx + (size_t)rem * TILE * n_seg * k_ * 4;
  off_cnt = off_dist + (size_t)rem * TILE * n_seg * k_ * 4;
  off_bp = (off_cnt + (size_t)rem * TILE * n_seg * 4 + 7) & ~(size_t)7;
  total = off_bp + (size_t)rem * TILE * n_seg * 8;
  // stream-ordered scratch; keep freed blocks cached in the device pool instead of returning them to the
  // driver at every synchronisation (the default release threshold is 0)
  static bool pool_configured = false;
  if (!pool_configured) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_configured = true;
  }
  uint8_t* scratch = nullptr;
  DBB_CUDA_TRY(cudaMallocAsync((void**)&scratch, total, st));
  DBB_CUDA_TRY(cudaMemsetAsync(scratch, 0, 256, st));
  P.item_counter = (uint32_t*)scratch;
  P.part_idx = (int32_t*)(scratch + off_idx);
  P.part_dist = (float*)(scratch + off_dist);
  P.part_cnt = (int32_t*)(scratch + off_cnt);
  P.part_bp = (int64_t*)(scratch + off_bp);

  const unsigned grid = (unsigned)std::min<int64_t>(P.n_items, sms);
  const int flags = (in->td_bound ? 1 : 0) | ((in->gc || in->cov) ? 2 : 0) | (masks ? 4 : 0) | (out->k > 0 ? 8 : 0);
  rc = launch_pairs(flags, grid, smem, st, tmap, P);
  if (rc == DBB_OK && rem > 0 && split_rows > 0) {
    merge_kernel<<<(unsigned)((split_rows + 3) / 4), 128, 0, st>>>(P, n_full * TILE, split_rows);
    if (cudaGetLastError() != cudaSuccess) { dbb::set_error("merge kernel launch failed"); rc = DBB_ERR_CUDA; }
  }
  cudaFreeAsync(scratch, st);
  return rc;
}
