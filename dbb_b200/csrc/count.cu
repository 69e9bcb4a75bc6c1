// Stage 1 (K1): canonical tetranucleotide counting + normalisation on packed sequence.
//
// Reference semantics (genomicSignatures.py:134-150): slide a 4-base window with stride 1, skip any
// window that contains a byte outside {A,C,G,T}, increment the canonical column, then divide by the
// sum (0/0 -> NaN).  The integer counts must be bit-exact.
//
// Hardware facts this design follows (profiles/microbench/, measured on B200):
//   * shared-memory atomics (ATOMS.POPC.INC) retire ~11.7 random updates / clk / SM no matter how the
//     histogram is replicated (16-18 when made bank-conflict-free at 32x the flush cost), i.e. one
//     atomic per WINDOW caps the kernel near 30 % of HBM bandwidth;
//   * so one atomic covers S windows: a (3+S)-mer histogram (4^(3+S) bins) is updated once every S
//     bases ("super-k-mer"), and folded to the 256 tetramer codes when the scaffold ends.  S=2/3/4
//     measured 21.5 / 32 / 42 windows/clk/SM.  The fold is exact integer arithmetic:
//         count4[c] = sum over window offsets o < S of the marginal of the (3+S)-mer histogram.
//   * the fold + re-zero costs O(4^(3+S)) shared-memory cycles per scaffold, so S is chosen per
//     scaffold from its length (short contigs S=1 ... megabase scaffolds S=4).
//
// Work decomposition: a work ITEM is a run of 64-base blocks of one scaffold (whole scaffold, or a
// chunk of a very long one); one CTA owns one item at a time and one histogram.  Lane i of a warp
// loads block i with a single 128-bit load (512 contiguous bytes per warp), the 3..6-base halo of the
// next block arrives by warp shuffle.  Windows whose super-k-mer is only partly valid (N, lowercase,
// scaffold end) take a rare per-window path.  Whole-scaffold items write counts and frequencies
// directly; chunks of split scaffolds add into a scratch row that a finalize kernel normalises.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Item {
  uint32_t blk_begin;   // first block (global block index)
  uint32_t n_blocks;
  uint32_t scaffold;
  uint32_t flags;       // bit31: last chunk of its scaffold; bits 29..30: (first block within scaffold) % 3;
                        // bits 0..28: split slot + 1 (0 = the item is the whole scaffold)
};
constexpr uint32_t ITEM_LAST = 1u << 31;
constexpr uint32_t ITEM_SLOT_MASK = (1u << 29) - 1;

__constant__ int16_t c_col_code1[DBB_NUM_KMERS];
__constant__ int16_t c_col_code2[DBB_NUM_KMERS];
__constant__ int c_debug_skip_flush = 0;   // profiling switch (DBB_K1_DEBUG_SKIP_FLUSH=1): count only, no fold / output

template <int THREADS>
__device__ __forceinline__ void group_sync() {
  if (THREADS == 32) __syncwarp(); else __syncthreads();
}

// Rare path: blocks that are not fully valid (N / IUPAC / lowercase, the padded last block of a scaffold, a
// block whose halo is invalid).  The warp handles them one at a time and COOPERATIVELY: the owning lane
// publishes the block through a 10-word shared scratch, lane a then treats anchor a (a, a+32 for S = 1).  A
// fully valid anchor still updates the (3+S)-mer histogram; otherwise its valid windows go one by one to the
// tetramer histogram.  ~60 instructions per slow block instead of ~1000 for a per-lane loop over anchors.
template <int S>
__device__ __noinline__ void slow_blocks(uint32_t slowmask, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                         uint32_t w4, uint64_t V, uint32_t Vh, int n_anchor,
                                         volatile uint32_t* sc, uint32_t* hS, uint32_t* c4, int lane) {
  constexpr int R = S + 3;
  constexpr int KB = 2 * R;
  while (slowmask) {
    const int src = __ffs(slowmask) - 1;
    slowmask &= slowmask - 1;
    if (lane == src) {
      sc[0] = w0; sc[1] = w1; sc[2] = w2; sc[3] = w3; sc[4] = w4; sc[5] = 0u;
      sc[6] = (uint32_t)(V >> 32); sc[7] = (uint32_t)V; sc[8] = Vh; sc[9] = (uint32_t)n_anchor;
    }
    __syncwarp();
    const uint64_t V64 = ((uint64_t)sc[6] << 32) | sc[7];
    const uint32_t vh = sc[8];
    const int na = (int)sc[9];
    for (int a = lane; a < na; a += 32) {
      const int p = a * S;
      // validity of positions p, p+1, ... in the top bits of x
      const uint64_t x = (V64 << p) | (p ? (((uint64_t)vh << 32) >> (64 - p)) : 0ull);
      const int k = p >> 4, o = p & 15;
      if ((x >> (64 - R)) == ((1ull << R) - 1ull)) {
        const uint64_t pair = ((uint64_t)sc[k] << 32) | sc[k + 1];
        atomicAdd(&hS[(uint32_t)(pair >> (64 - 2 * o - KB)) & ((1u << KB) - 1u)], 1u);
      } else {
        #pragma unroll
        for (int o2 = 0; o2 < S; ++o2) {
          if (((x >> (60 - o2)) & 15ull) == 15ull) {
            const int q = p + o2, k2 = q >> 4, oq = q & 15;
            const uint64_t pair = ((uint64_t)sc[k2] << 32) | sc[k2 + 1];
            atomicAdd(&c4[(uint32_t)(pair >> (64 - 2 * oq - 8)) & 255u], 1u);
          }
        }
      }
    }
    __syncwarp();
  }
}

// Fast path of one 64-base block whose bases (and the R-1 halo bases) are all valid: one unpredicated
// shared-memory atomic per anchor -- funnel shift, mask, ATOMS: 3 instructions per S windows.
//   w0..w3: code words of the block, w4: first code word of the next block (halo)
//   n_anchor: super-k-mers anchored in this block (64/S; 21 or 22 for S = 3)
template <int S>
__device__ __forceinline__ void fast_block(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                                           int n_anchor, uint32_t* hS) {
  constexpr int KB = 2 * (S + 3);
  constexpr uint32_t BINMASK4 = ((1u << KB) - 1u) << 2;
  constexpr int NA = (64 + S - 1) / S;                 // 64, 32, 22, 16
  const uint32_t w[5] = {w0, w1, w2, w3, w4};
  #pragma unroll
  for (int p = 0; p < 64; p += S) {
    const int k = p >> 4, o = p & 15;
    const int sh = 64 - 2 * o - KB - 2;                // >= 18
    const uint32_t x = (sh >= 32) ? (w[k] >> (sh - 32)) : __funnelshift_r(w[k + 1], w[k], sh);
    if (S == 3 && p == 63 && n_anchor != NA) continue; // S = 3: the anchor at 63 may belong to the next block
    atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hS) + (x & BINMASK4)), 1u);
  }
}

template <int S>
__device__ __forceinline__ bool block_all_valid(uint64_t V, uint32_t Vh) {
  constexpr uint32_t HALO = ~0u << (32 - (S + 2));     // the R-1 halo bases an anchor of this block may touch
  return (V == ~0ull) && ((Vh & HALO) == HALO);
}

template <int S, int THREADS>
__global__ void __launch_bounds__(THREADS) count_kernel(const Item* __restrict__ items, uint32_t n_items,
                                                        uint32_t* __restrict__ next_item,
                                                        const uint4* __restrict__ codes,
                                                        const uint64_t* __restrict__ valid,
                                                        uint32_t* __restrict__ counts,
                                                        float* __restrict__ freq, int64_t freq_stride,
                                                        uint32_t* __restrict__ split_counts) {
  constexpr int BINS = 1 << (2 * (3 + S));
  extern __shared__ uint32_t smem[];
  uint32_t* c4 = smem;                                   // [256]
  uint32_t* hS = (S == 1) ? smem : smem + 256;           // [BINS] (aliases c4 when S == 1)
  __shared__ uint32_t s_item;
  __shared__ uint32_t s_total[THREADS / 32];
  __shared__ uint32_t s_slow[THREADS / 32][10];
  const int tid = threadIdx.x, lane = tid & 31;

  for (int i = tid; i < (S == 1 ? 256 : 256 + BINS); i += THREADS) smem[i] = 0;
  group_sync<THREADS>();

  // Static schedule: CTA b takes items b, b + grid, ... of a list sorted by descending size, so every load of
  // the pipeline below is known one step ahead: the first block of the NEXT item is requested before the
  // current item is flushed, the next iteration's blocks before the current ones are processed.
  struct Fetch { uint4 c; uint64_t V; uint32_t hw, hv; };
  auto fetch = [&](const Item& im, uint32_t b) {
    Fetch f;
    f.c = make_uint4(0, 0, 0, 0); f.V = 0; f.hw = 0; f.hv = 0;
    if (b < im.n_blocks) {
      f.c = __ldg(codes + im.blk_begin + b);
      f.V = __ldg(valid + im.blk_begin + b);
      // the halo normally comes from the next lane; the last lane / last block fetches its own
      if (lane == 31 || b + 1 == im.n_blocks) {
        const bool has = (b + 1 < im.n_blocks) || !(im.flags & ITEM_LAST);
        if (has) {
          f.hw = __ldg(reinterpret_cast<const uint32_t*>(codes + im.blk_begin + b + 1));
          f.hv = (uint32_t)(__ldg(valid + im.blk_begin + b + 1) >> 32);
        }
      }
    }
    return f;
  };
  (void)next_item; (void)s_item;
  constexpr int NCOLT = (DBB_SIG_STRIDE + THREADS - 1) / THREADS;   // output columns per thread (136 + 4 padding)
  int cc1[NCOLT], cc2[NCOLT];
  #pragma unroll
  for (int q = 0; q < NCOLT; ++q) {
    const int col = tid + q * THREADS;
    cc1[q] = col < DBB_NUM_KMERS ? c_col_code1[col] : -1;
    cc2[q] = col < DBB_NUM_KMERS ? c_col_code2[col] : -1;
  }
  uint32_t it = blockIdx.x;
  Item item = Item{0, 0, 0, 0};
  Fetch nxt;
  nxt.c = make_uint4(0, 0, 0, 0); nxt.V = 0; nxt.hw = 0; nxt.hv = 0;
  if (it < n_items) { item = items[it]; nxt = fetch(item, tid); }

  while (it < n_items) {
    const uint32_t it_n = it + gridDim.x;
    const bool has_n = it_n < n_items;
    Item item_n = Item{0, 0, 0, 0};
    if (has_n) item_n = items[it_n];
    const uint32_t nb = item.n_blocks;
    const uint32_t slot = item.flags & ITEM_SLOT_MASK;
    const uint32_t first_mod3 = (item.flags >> 29) & 3u;

    for (uint32_t base = 0; base < nb || base == 0; base += THREADS) {
      const uint32_t b = base + tid;
      const bool act = b < nb;
      const Fetch cur = nxt;
      if (base + THREADS < nb) nxt = fetch(item, base + THREADS + tid);
      else if (has_n) nxt = fetch(item_n, tid);
      const uint4 c = cur.c;
      const uint64_t V = cur.V;
      uint32_t hw = __shfl_down_sync(0xffffffffu, c.x, 1);
      uint32_t hv = __shfl_down_sync(0xffffffffu, (uint32_t)(V >> 32), 1);
      if (lane == 31 || b + 1 >= nb) { hw = cur.hw; hv = cur.hv; }
      uint32_t w0 = c.x, w1 = c.y, w2 = c.z, w3 = c.w, w4 = hw;
      uint64_t Vb = V;
      uint32_t hvb = hv;
      int n_anchor = (64 + S - 1) / S;
      if (S == 3) {
        // anchors are the positions P = 0 mod 3 counted from the scaffold start; rotate the block so that
        // its first anchor sits at bit position 0
        const uint32_t bs = (first_mod3 + b) % 3u;            // (block index within scaffold) mod 3
        const uint32_t p0 = (3u - bs) % 3u;                   // 64 = 1 mod 3
        const uint32_t r2 = 2 * p0;
        w0 = __funnelshift_l(c.y, c.x, r2); w1 = __funnelshift_l(c.z, c.y, r2);
        w2 = __funnelshift_l(c.w, c.z, r2); w3 = __funnelshift_l(hw, c.w, r2);
        w4 = hw << r2;
        Vb = (V << p0) | (uint64_t)(p0 ? (hv >> (32 - p0)) : 0u);
        hvb = hv << p0;
        n_anchor = p0 == 0 ? 22 : 21;
      }
      const bool ok = act && block_all_valid<S>(Vb, hvb);
      if (ok) fast_block<S>(w0, w1, w2, w3, w4, n_anchor, hS);
      const uint32_t slow = __ballot_sync(0xffffffffu, act && !ok);
      if (slow) slow_blocks<S>(slow, w0, w1, w2, w3, w4, Vb, hvb, n_anchor, s_slow[tid >> 5], hS, c4, lane);
    }
    group_sync<THREADS>();
    if (c_debug_skip_flush) { it = it_n; item = item_n; continue; }

    if (S > 1) {
      // Fold the (3+S)-mer histogram into the 256 tetramer codes: for window offset o the count of code c is
      // the marginal over the o bases before and the S-1-o bases after it.  Thread <-> code, so no atomics
      // are needed (two thread groups split the offsets when the CTA has >= 512 threads); runs of >= 4 bins
      // are read as 128-bit words, rotated by the code so that neighbouring lanes hit different banks.
      constexpr int NPART = THREADS >= 512 ? 2 : 1;
      const int part = tid >> 8;
      if (part < NPART) {
        for (int c = tid & 255; c < 256; c += (THREADS < 256 ? THREADS : 256)) {
          uint32_t sum = 0;
          #pragma unroll
          for (int o = 0; o < S; ++o) {
            if (NPART == 2 && (o & 1) != part) continue;
            constexpr int dummy = 0; (void)dummy;
            const int e = S - 1 - o;                  // bases after the window
            const int run = 1 << (2 * e);             // consecutive bins per prefix value
            const int np = 1 << (2 * o);              // prefix values
            #pragma unroll
            for (int pfx = 0; pfx < np; ++pfx) {
              const int base = (pfx << (8 + 2 * e)) | (c << (2 * e));
              if (run >= 4) {
                const int nq = run >> 2;
                #pragma unroll
                for (int r4 = 0; r4 < nq; ++r4) {
                  const uint4 v = *reinterpret_cast<const uint4*>(hS + base + 4 * ((r4 + c) & (nq - 1)));
                  sum += (v.x + v.y) + (v.z + v.w);
                }
              } else {
                sum += hS[base];
              }
            }
          }
          if (NPART == 2) atomicAdd(&c4[c], sum); else c4[c] += sum;
        }
      }
      group_sync<THREADS>();
      for (int x = tid * 4; x < BINS; x += THREADS * 4) *reinterpret_cast<uint4*>(hS + x) = make_uint4(0, 0, 0, 0);
      // (the next item's first update is behind the two group_syncs below)
    }

    // canonical fold 256 -> 136 (column tables in registers), total = sum of the folded counts, one
    // reciprocal per scaffold: freq = (float)(n * (1.0 / total))  (0 * inf = NaN when nothing was counted,
    // as genomicSignatures.py:147-148 yields)
    uint32_t nloc[NCOLT];
    uint32_t part = 0;
    #pragma unroll
    for (int q = 0; q < NCOLT; ++q) {
      uint32_t nq = 0;
      if (cc1[q] >= 0) { nq = c4[cc1[q]]; if (cc2[q] >= 0) nq += c4[cc2[q]]; }
      nloc[q] = nq;
      part += nq;
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (THREADS > 32) {
      if (lane == 0) s_total[tid >> 5] = part;
      __syncthreads();
      part = 0;
      #pragma unroll
      for (int i = 0; i < THREADS / 32; ++i) part += s_total[i];
    } else {
      __syncwarp();
    }
    // every thread has read its c4 entries: re-zero for the next item (its first update is behind the final sync)
    for (int i = tid; i < 256; i += THREADS) c4[i] = 0;
    const uint32_t total = part;
    if (slot == 0) {
      const double inv = 1.0 / (double)total;
      #pragma unroll
      for (int q = 0; q < NCOLT; ++q) {
        const int col = tid + q * THREADS;
        if (col < DBB_NUM_KMERS && counts) counts[(size_t)item.scaffold * DBB_NUM_KMERS + col] = nloc[q];
        if (freq && col < (int)freq_stride)
          freq[(size_t)item.scaffold * freq_stride + col] = col < DBB_NUM_KMERS ? (float)((double)nloc[q] * inv) : 0.f;
      }
    } else {
      #pragma unroll
      for (int q = 0; q < NCOLT; ++q) {
        const int col = tid + q * THREADS;
        if (col < DBB_NUM_KMERS && nloc[q]) atomicAdd(&split_counts[(size_t)(slot - 1) * DBB_NUM_KMERS + col], nloc[q]);
      }
    }
    group_sync<THREADS>();
    it = it_n;
    item = item_n;
  }
}

// counts/frequencies of scaffolds that were split into several items
__global__ void __launch_bounds__(160) finalize_kernel(const uint32_t* __restrict__ split_counts,
                                                       const uint32_t* __restrict__ split_scaffold,
                                                       uint32_t n_split, uint32_t* __restrict__ counts,
                                                       float* __restrict__ freq, int64_t freq_stride) {
  __shared__ unsigned long long s_part[5];
  const uint32_t s = blockIdx.x;
  if (s >= n_split) return;
  const int col = threadIdx.x;
  const uint32_t sc = split_scaffold[s];
  uint32_t n = col < DBB_NUM_KMERS ? split_counts[(size_t)s * DBB_NUM_KMERS + col] : 0u;
  unsigned long long part = n;
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if ((col & 31) == 0) s_part[col >> 5] = part;
  __syncthreads();
  unsigned long long total = s_part[0] + s_part[1] + s_part[2] + s_part[3] + s_part[4];
  if (col < DBB_NUM_KMERS && counts) counts[(size_t)sc * DBB_NUM_KMERS + col] = n;
  if (freq) {
    for (int c = col; c < (int)freq_stride; c += 160) {
      freq[(size_t)sc * freq_stride + c] =
          c < DBB_NUM_KMERS ? (float)((double)split_counts[(size_t)s * DBB_NUM_KMERS + c] / (double)total) : 0.f;
    }
  }
}

struct ClassCfg { int S; int threads; };

}  // namespace

struct dbb_count_plan {
  int64_t n_seq = 0;
  int64_t total_blocks = 0;
  // per class (S = 1..4): item range in d_items
  uint32_t class_begin[4] = {0, 0, 0, 0};
  uint32_t class_count[4] = {0, 0, 0, 0};
  Item* d_items = nullptr;
  size_t cap_items = 0;
  uint32_t* d_counters = nullptr;       // [4] (kept for the ABI's memset; the schedule is static)
  uint32_t n_split = 0;
  size_t cap_split = 0;
  uint32_t* d_split_scaffold = nullptr;
  uint32_t* d_split_counts = nullptr;
  std::vector<Item> h_items;            // host staging of the last build (source of the async upload)
  std::vector<uint32_t> h_split;
  int sm_count = 0;
  int launches = 0;
  // the four class kernels run concurrently on their own streams so that the tail of one overlaps the others
  cudaStream_t class_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

bool g_tables_uploaded[64] = {false};   // per device ordinal (constants live in each device's module image)

int upload_tables() {
  // per device would be more general; one device per process is the deployment model (one rank per GPU)
  int dev = 0;
  DBB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { dbb::set_error("device ordinal out of range"); return DBB_ERR_UNSUPPORTED; }
  if (g_tables_uploaded[dev]) return DBB_OK;
  const dbb::KmerTable& t = dbb::kmer_table();
  DBB_CUDA_TRY(cudaMemcpyToSymbol(c_col_code1, t.col_code1, sizeof(t.col_code1)));
  DBB_CUDA_TRY(cudaMemcpyToSymbol(c_col_code2, t.col_code2, sizeof(t.col_code2)));
  { const char* e = getenv("DBB_K1_DEBUG_SKIP_FLUSH"); int v = (e && *e == '1') ? 1 : 0; DBB_CUDA_TRY(cudaMemcpyToSymbol(c_debug_skip_flush, &v, sizeof(int))); }
  g_tables_uploaded[dev] = true;
  return DBB_OK;
}

int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoll(v) : dflt;
}

template <int S, int THREADS>
int launch_class(const dbb_count_plan* plan, int cls, const void* d_codes, const void* d_valid,
                 uint32_t* d_counts, float* d_freq, int64_t freq_stride, cudaStream_t st) {
  const uint32_t n = plan->class_count[cls];
  if (n == 0) return DBB_OK;
  constexpr int BINS = 1 << (2 * (3 + S));
  const size_t smem = (S == 1 ? 256 : 256 + BINS) * sizeof(uint32_t);
  auto kern = count_kernel<S, THREADS>;
  DBB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  DBB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (occ < 1) { dbb::set_error("count kernel S=%d does not fit on the device", S); return DBB_ERR_CUDA; }
  // DBB_K1_OCC<S> caps the CTAs per SM of a class so that the class kernels share the SMs from the start instead of
  // overlapping only at their tails (tuning switch; 0 = the occupancy limit)
  static const int64_t cap = env_i64(S == 1 ? "DBB_K1_OCC1" : S == 2 ? "DBB_K1_OCC2" : S == 3 ? "DBB_K1_OCC3" : "DBB_K1_OCC4", 0);
  if (cap > 0 && cap < occ) occ = (int)cap;
  uint32_t grid = (uint32_t)std::min<int64_t>((int64_t)occ * plan->sm_count, (int64_t)n);
  kern<<<grid, THREADS, smem, st>>>(plan->d_items + plan->class_begin[cls], n, plan->d_counters + cls,
                                    (const uint4*)d_codes, (const uint64_t*)d_valid, d_counts, d_freq,
                                    freq_stride, plan->d_split_counts);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

}  // namespace

namespace dbb {

// (Re)build the work list of `plan` for a new batch, reusing its device allocations, streams and events when
// they are large enough.  The upload is enqueued on `st`.
int count_plan_rebuild(dbb_count_plan* plan, const int64_t* h_blk_off, int64_t n_seq, cudaStream_t st) {
  DBB_REQUIRE(plan != nullptr && n_seq >= 0 && (n_seq == 0 || h_blk_off != nullptr), "bad arguments");
  DBB_REQUIRE(n_seq < (1ll << 32) - 1, "too many scaffolds");
  plan->n_seq = n_seq;
  plan->total_blocks = n_seq ? h_blk_off[n_seq] : 0;
  if (plan->total_blocks >= (1ll << 32)) { dbb::set_error("batch exceeds 2^32 blocks"); return DBB_ERR_INVALID; }

  // class thresholds in blocks (64 bases); env overrides are for tuning runs only
  static const int64_t t1 = env_i64("DBB_K1_T1", 128);         // <= 8 kb    : S = 1, one warp per scaffold
  static const int64_t t2 = env_i64("DBB_K1_T2", 512);         // <= 33 kb   : S = 2
  static const int64_t t3 = env_i64("DBB_K1_T3", 4300);        // <= 275 kb  : S = 3
  static const int64_t chunk = env_i64("DBB_K1_CHUNK", 16384); // S = 4 chunk: 1 Mb
  std::vector<Item> cls[4];
  plan->h_split.clear();
  for (int64_t s = 0; s < n_seq; ++s) {
    const int64_t b0 = h_blk_off[s], nb = h_blk_off[s + 1] - b0;
    if (nb < 0) { dbb::set_error("block offsets must be non-decreasing"); return DBB_ERR_INVALID; }
    if (nb <= chunk) {
      int c = nb <= t1 ? 0 : nb <= t2 ? 1 : nb <= t3 ? 2 : 3;
      cls[c].push_back(Item{(uint32_t)b0, (uint32_t)nb, (uint32_t)s, ITEM_LAST});
    } else {
      const uint32_t slot = (uint32_t)plan->h_split.size() + 1;
      if (slot >= ITEM_SLOT_MASK) { dbb::set_error("too many split scaffolds"); return DBB_ERR_INVALID; }
      plan->h_split.push_back((uint32_t)s);
      for (int64_t o = 0; o < nb; o += chunk) {
        const int64_t len = std::min(chunk, nb - o);
        uint32_t flags = slot | (uint32_t)((o % 3) << 29) | (o + len == nb ? ITEM_LAST : 0u);
        cls[3].push_back(Item{(uint32_t)(b0 + o), (uint32_t)len, (uint32_t)s, flags});
      }
    }
  }
  plan->h_items.clear();
  plan->h_items.reserve((size_t)n_seq + 16);
  for (int c = 0; c < 4; ++c) {
    std::stable_sort(cls[c].begin(), cls[c].end(), [](const Item& a, const Item& b) { return a.n_blocks > b.n_blocks; });
    plan->class_begin[c] = (uint32_t)plan->h_items.size();
    plan->class_count[c] = (uint32_t)cls[c].size();
    plan->h_items.insert(plan->h_items.end(), cls[c].begin(), cls[c].end());
  }
  plan->n_split = (uint32_t)plan->h_split.size();
  plan->launches = 0;
  for (int c = 0; c < 4; ++c) plan->launches += plan->class_count[c] ? 1 : 0;
  if (plan->n_split) plan->launches += 1;

  cudaError_t e = cudaSuccess;
  if (plan->h_items.size() > plan->cap_items) {
    cudaFree(plan->d_items); plan->d_items = nullptr;
    plan->cap_items = plan->h_items.size() + plan->h_items.size() / 4 + 64;
    e = cudaMalloc(&plan->d_items, plan->cap_items * sizeof(Item));
    if (e != cudaSuccess) plan->cap_items = 0;
  }
  if (e == cudaSuccess && !plan->h_items.empty())
    e = cudaMemcpyAsync(plan->d_items, plan->h_items.data(), plan->h_items.size() * sizeof(Item), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && plan->n_split > plan->cap_split) {
    cudaFree(plan->d_split_scaffold); cudaFree(plan->d_split_counts);
    plan->d_split_scaffold = nullptr; plan->d_split_counts = nullptr;
    plan->cap_split = (size_t)plan->n_split + 16;
    e = cudaMalloc(&plan->d_split_scaffold, plan->cap_split * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&plan->d_split_counts, plan->cap_split * DBB_NUM_KMERS * sizeof(uint32_t));
    if (e != cudaSuccess) plan->cap_split = 0;
  }
  if (e == cudaSuccess && plan->n_split)
    e = cudaMemcpyAsync(plan->d_split_scaffold, plan->h_split.data(), plan->n_split * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    dbb::set_error("count plan allocation/upload failed: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? DBB_ERR_NOMEM : DBB_ERR_CUDA;
  }
  return DBB_OK;
}

}  // namespace dbb

DBB_EXPORT int dbb_count_plan_create(const int64_t* h_blk_off, int64_t n_seq, dbb_count_plan** out) {
  DBB_REQUIRE(out != nullptr, "bad arguments");
  *out = nullptr;
  int rc = upload_tables();
  if (rc != DBB_OK) return rc;
  dbb_count_plan* plan = new dbb_count_plan();
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaMalloc(&plan->d_counters, 4 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming);
  for (int c = 0; c < 4 && e == cudaSuccess; ++c) {
    e = cudaStreamCreateWithFlags(&plan->class_stream[c], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&plan->ev_join[c], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    dbb::set_error("count plan setup failed: %s", cudaGetErrorString(e));
    dbb_count_plan_destroy(plan);
    return DBB_ERR_CUDA;
  }
  rc = dbb::count_plan_rebuild(plan, h_blk_off, n_seq, nullptr);
  if (rc == DBB_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) { dbb::set_error("count plan upload failed"); rc = DBB_ERR_CUDA; }
  if (rc != DBB_OK) { dbb_count_plan_destroy(plan); return rc; }
  *out = plan;
  return DBB_OK;
}

DBB_EXPORT int dbb_count_plan_destroy(dbb_count_plan* plan) {
  if (!plan) return DBB_OK;
  cudaFree(plan->d_items);
  cudaFree(plan->d_counters);
  cudaFree(plan->d_split_scaffold);
  cudaFree(plan->d_split_counts);
  for (int c = 0; c < 4; ++c) {
    if (plan->class_stream[c]) cudaStreamDestroy(plan->class_stream[c]);
    if (plan->ev_join[c]) cudaEventDestroy(plan->ev_join[c]);
  }
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  delete plan;
  return DBB_OK;
}

DBB_EXPORT int dbb_count_plan_launches(const dbb_count_plan* plan) { return plan ? plan->launches : 0; }

DBB_EXPORT int dbb_tetra_count_dev(dbb_count_plan* plan, const void* d_codes, const void* d_valid,
                                   uint32_t* d_counts, float* d_freq, int64_t freq_stride, void* stream) {
  DBB_REQUIRE(plan != nullptr, "plan is NULL");
  if (plan->n_seq == 0) return DBB_OK;
  DBB_REQUIRE(plan->total_blocks == 0 || (d_codes && d_valid), "NULL packed planes");
  DBB_REQUIRE(d_counts || d_freq, "no output requested");
  DBB_REQUIRE(!d_freq || freq_stride >= DBB_NUM_KMERS, "freq_stride < 136");
  if (!d_freq) freq_stride = DBB_NUM_KMERS;
  cudaStream_t st = (cudaStream_t)stream;
  DBB_CUDA_TRY(cudaMemsetAsync(plan->d_counters, 0, 4 * sizeof(uint32_t), st));
  if (plan->n_split)
    DBB_CUDA_TRY(cudaMemsetAsync(plan->d_split_counts, 0, (size_t)plan->n_split * DBB_NUM_KMERS * sizeof(uint32_t), st));
  // fork: every class kernel on its own stream behind the memsets, join back into the caller's stream
  DBB_CUDA_TRY(cudaEventRecord(plan->ev_fork, st));
  int rc = DBB_OK;
  for (int c = 3; c >= 0 && rc == DBB_OK; --c) {            // longest-running class first
    if (plan->class_count[c] == 0) continue;
    cudaStream_t cs = plan->class_stream[c];
    DBB_CUDA_TRY(cudaStreamWaitEvent(cs, plan->ev_fork, 0));
    // CTA sizes: env overrides exist for tuning sweeps (profiles/), the defaults are the measured best
    static const int thr4 = (int)env_i64("DBB_K1_THR4", 512), thr3 = (int)env_i64("DBB_K1_THR3", 256),
                     thr2 = (int)env_i64("DBB_K1_THR2", 64);
#define DBB_LAUNCH(S_, T_) rc = launch_class<S_, T_>(plan, c, d_codes, d_valid, d_counts, d_freq, freq_stride, cs)
    switch (c) {
      case 3: if (thr4 >= 1024) DBB_LAUNCH(4, 1024); else DBB_LAUNCH(4, 512); break;
      case 2: if (thr3 >= 512) DBB_LAUNCH(3, 512); else if (thr3 >= 256) DBB_LAUNCH(3, 256); else if (thr3 >= 128) DBB_LAUNCH(3, 128); else DBB_LAUNCH(3, 64); break;
      case 1: if (thr2 >= 256) DBB_LAUNCH(2, 256); else if (thr2 >= 128) DBB_LAUNCH(2, 128); else if (thr2 >= 64) DBB_LAUNCH(2, 64); else DBB_LAUNCH(2, 32); break;
      default: DBB_LAUNCH(1, 32); break;
    }
#undef DBB_LAUNCH
    if (rc != DBB_OK) return rc;
    DBB_CUDA_TRY(cudaEventRecord(plan->ev_join[c], cs));
    DBB_CUDA_TRY(cudaStreamWaitEvent(st, plan->ev_join[c], 0));
  }
  if (plan->n_split) {
    finalize_kernel<<<plan->n_split, 160, 0, st>>>(plan->d_split_counts, plan->d_split_scaffold, plan->n_split,
                                                   d_counts, d_freq, freq_stride);
    DBB_CUDA_TRY(cudaGetLastError());
  }
  return DBB_OK;
}
