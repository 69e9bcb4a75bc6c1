// "Next" row N1: rectangular filtered nearest-core search.
//
// Reference semantics (refineBins.py:42-96, the same loop in unbinned.py:156-176): for every unbinned contig u,
// over all core bins c in iteration order,
//   valid(u,c) = withinDistGC(u,c) and withinDistCov(u,c) and L1(u.sig, c.sig) < boundDistTD(u)
//     (distributions.py:75-127 with fixedSeqLen=None: GC bounds by (core mean-GC key, length key of u), coverage
//      and TD bounds by the length key of u; core.coverage <= 0 -> 0.1)
//   among the valid cores: closest in GC (|u.GC - c.GC|), in TD (the L1 distance) and in coverage
//   (|u.cov - c.cov|) space, strict '<' so the first core in iteration order wins ties;
//   u is assigned iff the three closest cores are the same one.
// Everything is FP64, as in the reference (signatures are float64 there); the problem is U x B with B << N, so
// FP64 throughput is ample (one warp per unbinned contig, cores streamed from L2).
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace {

struct RefineParams {
  dbb_refine_input in;
  dbb_refine_output out;
  // N5 (merge.py:96-156): the queries are binned contigs; exclude[u] (the own bin) is skipped by the closest-bin search
  // and the per-pair flags / closest bins are tallied per (own bin, other bin).  N6 (unbinned.py:111-186): nothing is
  // excluded, the tallies go per (scaffold of the contig, bin) with a 7th counter (closest in TD and in coverage space)
  const int32_t* group;          // [U] tally row of the query, or nullptr
  const int32_t* exclude;        // [U] core skipped by the query, or nullptr
  int32_t* table;                // [n_groups][B][table_stride] or nullptr
  int64_t n_groups;
  int table_stride;              // 6 or 7
};

__device__ __forceinline__ double warp_sum(double v) {
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void __launch_bounds__(256) refine_kernel(const RefineParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const dbb_refine_input& I = P.in;
  if (u >= I.n_unbinned) return;
  const int64_t B = I.n_cores;
  // the contig's signature: lane holds dimensions lane, lane+32, ...
  double us[5];
  #pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int k = lane + 32 * q;
    us[q] = k < DBB_NUM_KMERS ? I.u_sig[u * DBB_NUM_KMERS + k] : 0.0;
  }
  const double ugc = I.u_gc[u], ucov = I.u_cov[u], tdb = I.u_td_bound[u];
  const int gcl = I.u_gc_lbin[u], covl = I.u_cov_lbin[u];
  const double cov_lo = I.cov_lo[covl], cov_hi = I.cov_hi[covl];
  const int64_t own = P.exclude ? P.exclude[u] : -1;
  const int64_t grp = P.group ? P.group[u] : 0;
  const bool all_td = P.out.within || P.out.dist || P.table;   // the TD distance to every core is wanted
  int best[3] = {-1, -1, -1};
  double bestd[3] = {1e9, 1e9, 1e9};                     // refineBins.py:56

  for (int64_t c0 = 0; c0 < B; c0 += 32) {
    // lanes evaluate the scalar predicates of 32 cores at once
    const int64_t c = c0 + lane;
    bool p_gc = false, p_cov = false;
    double cgc = 0.0, ccov = 0.0;
    if (c < B) {
      cgc = I.c_gc[c]; ccov = I.c_cov[c];
      const int m = I.c_gc_mbin[c];
      const double lo = I.gc_lo[m * I.gc_nl + gcl], hi = I.gc_hi[m * I.gc_nl + gcl];
      const double gdiff = ugc - cgc;
      p_gc = gdiff >= lo && gdiff <= hi;
      const double eff = ccov > 0 ? ccov : 0.1;
      const double cdiff = (ucov - eff) * 100.0 / eff;
      p_cov = cdiff >= cov_lo && cdiff <= cov_hi;
    }
    const uint32_t m_gc = __ballot_sync(0xffffffffu, p_gc), m_cov = __ballot_sync(0xffffffffu, p_cov);
    // the TD distance is needed for cores that passed both (or for all of them when the flag matrix is wanted)
    uint32_t todo = all_td ? __ballot_sync(0xffffffffu, c < B) : (m_gc & m_cov);
    if (own >= c0 && own < c0 + 32 && !all_td) todo &= ~(1u << (own - c0));
    uint32_t m_td = 0;
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const double* cs = I.c_sig + (c0 + j) * DBB_NUM_KMERS;
      double part = 0.0;
      #pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int k = lane + 32 * q;
        if (k < DBB_NUM_KMERS) part += fabs(us[q] - cs[k]);
      }
      const double d = warp_sum(part);
      if (P.out.dist && lane == 0) P.out.dist[u * B + c0 + j] = d;
      const bool p_td = d < tdb;                         // distributions.py:103-104 (strict)
      if (p_td) m_td |= 1u << j;
      if (p_td && ((m_gc & m_cov) >> j & 1u) && c0 + j != own) {
        const double dg = fabs(ugc - __shfl_sync(0xffffffffu, cgc, j));
        const double dc = fabs(ucov - __shfl_sync(0xffffffffu, ccov, j));
        const int ci = (int)(c0 + j);
        if (dg < bestd[0]) { bestd[0] = dg; best[0] = ci; }
        if (d < bestd[1]) { bestd[1] = d; best[1] = ci; }
        if (dc < bestd[2]) { bestd[2] = dc; best[2] = ci; }
      } else {
        __shfl_sync(0xffffffffu, cgc, j); __shfl_sync(0xffffffffu, ccov, j);   // keep the warp converged
      }
    }
    if (P.out.within && c < B)
      P.out.within[u * B + c] = (uint8_t)((p_gc ? 1 : 0) | (p_cov ? 2 : 0) | ((m_td >> lane & 1u) ? 4 : 0));
    if (P.table && c < B && c != own) {                  // merge.py:112-119 (bin2 != bin1), unbinned.py:137-152
      int32_t* t = P.table + (grp * B + c) * P.table_stride;
      if (p_gc) atomicAdd(t, 1);
      if (p_cov) atomicAdd(t + 1, 1);
      if (m_td >> lane & 1u) atomicAdd(t + 2, 1);
    }
  }
  if (lane == 0) {
    if (P.out.closest) { P.out.closest[u * 3] = best[0]; P.out.closest[u * 3 + 1] = best[1]; P.out.closest[u * 3 + 2] = best[2]; }
    if (P.out.min_dist) { P.out.min_dist[u * 3] = bestd[0]; P.out.min_dist[u * 3 + 1] = bestd[1]; P.out.min_dist[u * 3 + 2] = bestd[2]; }
    if (P.out.assigned) P.out.assigned[u] = (best[0] >= 0 && best[0] == best[1] && best[0] == best[2]) ? best[0] : -1;
    if (P.table) {                                       // merge.py:147-152, unbinned.py:174-182
      #pragma unroll
      for (int i = 0; i < 3; ++i)
        if (best[i] >= 0) atomicAdd(P.table + (grp * B + best[i]) * P.table_stride + 3 + i, 1);
      if (P.table_stride == 7 && best[1] >= 0 && best[1] == best[2]) atomicAdd(P.table + (grp * B + best[1]) * 7 + 6, 1);
    }
  }
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <typename T> int up(const T* h, size_t n, cudaStream_t st) {
    if (cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) { dbb::set_error("cudaMalloc failed in dbb_refine_host"); return DBB_ERR_NOMEM; }
    if (n && cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) { dbb::set_error("H2D copy failed in dbb_refine_host"); return DBB_ERR_CUDA; }
    return DBB_OK;
  }
  int alloc(size_t bytes) {
    if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) { dbb::set_error("cudaMalloc failed in dbb_refine_host"); return DBB_ERR_NOMEM; }
    return DBB_OK;
  }
};

int check_input(const dbb_refine_input* in, const dbb_refine_output* out) {
  DBB_REQUIRE(in && out, "NULL argument struct");
  DBB_REQUIRE(in->n_unbinned >= 0 && in->n_cores >= 0, "negative size");
  if (in->n_unbinned == 0) return DBB_OK;
  DBB_REQUIRE(in->u_sig && in->u_gc && in->u_cov && in->u_td_bound && in->u_gc_lbin && in->u_cov_lbin, "incomplete unbinned inputs");
  DBB_REQUIRE(in->n_cores == 0 || (in->c_sig && in->c_gc && in->c_cov && in->c_gc_mbin), "incomplete core inputs");
  DBB_REQUIRE(in->gc_lo && in->gc_hi && in->cov_lo && in->cov_hi && in->gc_nm > 0 && in->gc_nl > 0 && in->cov_nl > 0, "incomplete tables");
  DBB_REQUIRE(out->closest || out->assigned || out->within || out->min_dist || out->dist, "no output requested");
  return DBB_OK;
}

struct Tally {                    // device pointers (launch) or host pointers (run_host)
  const int32_t* group = nullptr; const int32_t* exclude = nullptr; int32_t* table = nullptr;
  int64_t n_groups = 0; int stride = 6;
  size_t table_bytes(int64_t B) const { return (size_t)n_groups * (size_t)B * stride * sizeof(int32_t); }
};

int launch(const dbb_refine_input* in, const dbb_refine_output* out, const Tally& tl, cudaStream_t st) {
  RefineParams P;
  P.in = *in; P.out = *out; P.group = tl.group; P.exclude = tl.exclude; P.table = tl.table; P.n_groups = tl.n_groups;
  P.table_stride = tl.stride;
  int32_t* d_table = tl.table;
  const int64_t grid = (in->n_unbinned + 7) / 8;
  DBB_REQUIRE(grid < (1ll << 31), "too many query contigs for one launch");
  if (d_table && tl.table_bytes(in->n_cores)) DBB_CUDA_TRY(cudaMemsetAsync(d_table, 0, tl.table_bytes(in->n_cores), st));
  refine_kernel<<<(unsigned)grid, 256, 0, st>>>(P);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

// host buffers in / out: uploads, one launch, downloads (tl: the tallies of N5 / N6, host pointers)
int run_host(const dbb_refine_input* in, const dbb_refine_output* out, const Tally& tl) {
  int32_t* table = tl.table;
  const size_t U = (size_t)in->n_unbinned, B = (size_t)in->n_cores;
  int rc = DBB_OK;
  cudaStream_t st;
  DBB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  Buf usig, csig, ugc, ucov, cgc, ccov, utd, ugl, ucl, cgm, glo, ghi, clo, chi, o_cl, o_md, o_as, o_wi, o_di, b_grp, b_exc, o_tb;
  Tally dtl = tl;
  dbb_refine_input d = *in;
  dbb_refine_output o = *out;
  do {
#define UP(buf, field, count) if ((rc = buf.up(in->field, count, st)) != DBB_OK) break; d.field = (decltype(d.field))buf.p;
    UP(usig, u_sig, U * DBB_NUM_KMERS) UP(csig, c_sig, B * DBB_NUM_KMERS) UP(ugc, u_gc, U) UP(ucov, u_cov, U)
    UP(cgc, c_gc, B) UP(ccov, c_cov, B) UP(utd, u_td_bound, U) UP(ugl, u_gc_lbin, U) UP(ucl, u_cov_lbin, U) UP(cgm, c_gc_mbin, B)
    UP(glo, gc_lo, (size_t)in->gc_nm * in->gc_nl) UP(ghi, gc_hi, (size_t)in->gc_nm * in->gc_nl)
    UP(clo, cov_lo, (size_t)in->cov_nl) UP(chi, cov_hi, (size_t)in->cov_nl)
#undef UP
    if (tl.group) { if ((rc = b_grp.up(tl.group, U, st)) != DBB_OK) break; dtl.group = (const int32_t*)b_grp.p; }
    if (tl.exclude) { if ((rc = b_exc.up(tl.exclude, U, st)) != DBB_OK) break; dtl.exclude = (const int32_t*)b_exc.p; }
    if (table) { if ((rc = o_tb.alloc(tl.table_bytes(B))) != DBB_OK) break; dtl.table = (int32_t*)o_tb.p; }
    if (out->closest) { if ((rc = o_cl.alloc(U * 12)) != DBB_OK) break; o.closest = (int32_t*)o_cl.p; }
    if (out->min_dist) { if ((rc = o_md.alloc(U * 24)) != DBB_OK) break; o.min_dist = (double*)o_md.p; }
    if (out->assigned) { if ((rc = o_as.alloc(U * 4)) != DBB_OK) break; o.assigned = (int32_t*)o_as.p; }
    if (out->within) { if ((rc = o_wi.alloc(U * std::max<size_t>(B, 1))) != DBB_OK) break; o.within = (uint8_t*)o_wi.p; }
    if (out->dist) { if ((rc = o_di.alloc(U * std::max<size_t>(B, 1) * 8)) != DBB_OK) break; o.dist = (double*)o_di.p; }
    if ((rc = launch(&d, &o, dtl, st)) != DBB_OK) break;
    cudaError_t e = cudaSuccess;
    if (out->closest) e = cudaMemcpyAsync(out->closest, o.closest, U * 12, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out->min_dist) e = cudaMemcpyAsync(out->min_dist, o.min_dist, U * 24, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out->assigned) e = cudaMemcpyAsync(out->assigned, o.assigned, U * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out->within && B) e = cudaMemcpyAsync(out->within, o.within, U * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out->dist && B) e = cudaMemcpyAsync(out->dist, o.dist, U * B * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && table && tl.table_bytes(B)) e = cudaMemcpyAsync(table, o_tb.p, tl.table_bytes(B), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { dbb::set_error("filtered nearest-core search failed: %s", cudaGetErrorString(e)); rc = DBB_ERR_CUDA; }
  } while (0);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return rc;
}

int check_merge(const dbb_merge_input* in, const int32_t* table) {
  DBB_REQUIRE(in && table, "NULL argument");
  DBB_REQUIRE(in->search.n_unbinned == 0 || in->own, "own bin of every contig is required");
  DBB_REQUIRE(in->search.n_cores < 18919, "too many bins for an int32 [B][B][6] table");
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_refine_dev(const dbb_refine_input* in, const dbb_refine_output* out, void* stream) {
  int rc = check_input(in, out);
  if (rc != DBB_OK || in->n_unbinned == 0) return rc;
  return launch(in, out, Tally(), (cudaStream_t)stream);
}

DBB_EXPORT int dbb_refine_host(const dbb_refine_input* in, const dbb_refine_output* out) {
  int rc = check_input(in, out);
  if (rc != DBB_OK || in->n_unbinned == 0) return rc;
  return run_host(in, out, Tally());
}

// N5: merge.py:96-156.  own[] values are the caller's responsibility on the device entry point; the host entry
// point checks them.
DBB_EXPORT int dbb_merge_table_dev(const dbb_merge_input* in, int32_t* d_table, int32_t* d_closest, void* stream) {
  int rc = check_merge(in, d_table);
  if (rc != DBB_OK) return rc;
  dbb_refine_output out = {d_closest, nullptr, nullptr, nullptr, nullptr};
  dbb_refine_output chk = out;
  chk.assigned = reinterpret_cast<int32_t*>(d_table);     // "some output requested" for check_input
  if ((rc = check_input(&in->search, &chk)) != DBB_OK) return rc;
  if (in->search.n_unbinned == 0) {
    DBB_CUDA_TRY(cudaMemsetAsync(d_table, 0, (size_t)in->search.n_cores * in->search.n_cores * 6 * sizeof(int32_t), (cudaStream_t)stream));
    return DBB_OK;
  }
  Tally tl;
  tl.group = in->own; tl.exclude = in->own; tl.table = d_table; tl.n_groups = in->search.n_cores; tl.stride = 6;
  return launch(&in->search, &out, tl, (cudaStream_t)stream);
}

DBB_EXPORT int dbb_merge_table_host(const dbb_merge_input* in, int32_t* table, int32_t* closest) {
  int rc = check_merge(in, table);
  if (rc != DBB_OK) return rc;
  dbb_refine_output out = {closest, nullptr, nullptr, nullptr, nullptr};
  dbb_refine_output chk = out;
  chk.assigned = table;
  if ((rc = check_input(&in->search, &chk)) != DBB_OK) return rc;
  const int64_t U = in->search.n_unbinned, B = in->search.n_cores;
  memset(table, 0, (size_t)B * B * 6 * sizeof(int32_t));
  if (U == 0) return DBB_OK;
  for (int64_t u = 0; u < U; ++u) DBB_REQUIRE(in->own[u] >= 0 && in->own[u] < B, "own[] out of range");
  Tally tl;
  tl.group = in->own; tl.exclude = in->own; tl.table = table; tl.n_groups = B; tl.stride = 6;
  return run_host(&in->search, &out, tl);
}

// N6: unbinned.py:111-186.  The scaffold's contigs against every bin; nothing excluded; 7 counters per (scaffold, bin).
namespace {
int check_unbinned(const dbb_unbinned_input* in, const int32_t* table) {
  DBB_REQUIRE(in && table, "NULL argument");
  DBB_REQUIRE(in->n_groups >= 0 && (in->search.n_unbinned == 0 || in->group), "the scaffold index of every contig is required");
  DBB_REQUIRE((double)in->n_groups * (double)in->search.n_cores * 7.0 < 2147483648.0, "table too large");
  return DBB_OK;
}
}  // namespace

DBB_EXPORT int dbb_unbinned_table_dev(const dbb_unbinned_input* in, int32_t* d_table, void* stream) {
  int rc = check_unbinned(in, d_table);
  if (rc != DBB_OK) return rc;
  dbb_refine_output out = {nullptr, nullptr, nullptr, nullptr, nullptr};
  dbb_refine_output chk = out;
  chk.assigned = d_table;
  if ((rc = check_input(&in->search, &chk)) != DBB_OK) return rc;
  Tally tl;
  tl.group = in->group; tl.table = d_table; tl.n_groups = in->n_groups; tl.stride = 7;
  if (in->search.n_unbinned == 0) {
    if (tl.table_bytes(in->search.n_cores)) DBB_CUDA_TRY(cudaMemsetAsync(d_table, 0, tl.table_bytes(in->search.n_cores), (cudaStream_t)stream));
    return DBB_OK;
  }
  return launch(&in->search, &out, tl, (cudaStream_t)stream);
}

DBB_EXPORT int dbb_unbinned_table_host(const dbb_unbinned_input* in, int32_t* table) {
  int rc = check_unbinned(in, table);
  if (rc != DBB_OK) return rc;
  dbb_refine_output out = {nullptr, nullptr, nullptr, nullptr, nullptr};
  dbb_refine_output chk = out;
  chk.assigned = table;
  if ((rc = check_input(&in->search, &chk)) != DBB_OK) return rc;
  const int64_t U = in->search.n_unbinned, B = in->search.n_cores;
  Tally tl;
  tl.group = in->group; tl.table = table; tl.n_groups = in->n_groups; tl.stride = 7;
  memset(table, 0, tl.table_bytes(B));
  if (U == 0) return DBB_OK;
  for (int64_t u = 0; u < U; ++u) DBB_REQUIRE(in->group[u] >= 0 && in->group[u] < in->n_groups, "group[] out of range");
  return run_host(&in->search, &out, tl);
}
