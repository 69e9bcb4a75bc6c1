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

template <int THREADS>
__device__ __forceinline__ void group_sync() {
  if (THREADS == 32) __syncwarp(); else __syncthreads();
}

__device__ __forceinline__ void red_inc_if(uint32_t smem_addr, uint32_t mask_word, uint32_t bit) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.b32 q, t, 0;\n\t"
      "@q red.shared.add.u32 [%0], 1;\n\t}"
      :: "r"(smem_addr), "r"(mask_word), "r"(bit) : "memory");
}

// Update the histograms with one 64-base block.
//   w[0..3]: code words of the block, w[4]: first code word of the next block (halo)
//   V: validity of the 64 bases (bit 63-p), Vh: validity of the next block's first 32 bases (bit 31-p)
//   hS: shared address of the (3+S)-mer histogram, c4: shared address of the 256-bin tetramer histogram
template <int S>
__device__ __forceinline__ void process_block(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                              uint32_t w4, uint64_t V, uint32_t Vh, int n_anchor,
                                              uint32_t hS, uint32_t c4) {
  constexpr int R = S + 3;
  constexpr int KB = 2 * R;
  constexpr uint32_t BINMASK4 = ((1u << KB) - 1u) << 2;
  const uint32_t w[5] = {w0, w1, w2, w3, w4};

  // wv bit (63-p): the 4 bases p..p+3 are valid; sv: the R bases p..p+R-1 are valid
  uint64_t wv = V;
  #pragma unroll
  for (int t = 1; t < 4; ++t) wv &= (V << t) | (uint64_t)(Vh >> (32 - t));
  uint64_t sv = wv;
  #pragma unroll
  for (int t = 4; t < R; ++t) sv &= (V << t) | (uint64_t)(Vh >> (32 - t));
  if (S == 3 && n_anchor < 22) sv &= ~1ull;            // anchor 63 belongs to the next block
  const uint32_t sv_hi = (uint32_t)(sv >> 32), sv_lo = (uint32_t)sv;

  #pragma unroll
  for (int p = 0; p < 64; p += S) {
    const int k = p >> 4, o = p & 15;
    const int sh = 64 - 2 * o - KB - 2;                // >= 18
    uint32_t x = (sh >= 32) ? (w[k] >> (sh - 32)) : __funnelshift_r(w[k + 1], w[k], sh);
    uint32_t off = x & BINMASK4;
    if (p < 32) red_inc_if(hS + off, sv_hi, 1u << (31 - p));
    else        red_inc_if(hS + off, sv_lo, 1u << (63 - p));
  }

  if (S > 1) {
    // anchors whose super-k-mer is not fully valid but which still own >= 1 valid window
    uint32_t wvh = Vh & (Vh << 1) & (Vh << 2) & (Vh << 3);
    uint64_t anyw = wv;
    #pragma unroll
    for (int t = 1; t < S; ++t) anyw |= (wv << t) | (uint64_t)(wvh >> (32 - t));
    uint64_t anchors = 0;
    #pragma unroll
    for (int p = 0; p < 64; p += S) anchors |= 1ull << (63 - p);
    if (S == 3 && n_anchor < 22) anchors &= ~1ull;
    uint64_t need = ~sv & anchors & anyw;
    if (need != 0) {
      #pragma unroll
      for (int p = 0; p < 64; p += S) {
        if ((need >> (63 - p)) & 1ull) {
          #pragma unroll
          for (int o2 = 0; o2 < S; ++o2) {
            const int q = p + o2;
            bool ok = (q < 64) ? ((wv >> (63 - q)) & 1ull) : ((wvh >> (31 - (q - 64))) & 1u);
            if (ok) {
              const int k = q >> 4, o = q & 15;
              const int sh = 64 - 2 * o - 8 - 2;
              uint32_t x = (sh >= 32) ? (w[k] >> (sh - 32)) : __funnelshift_r(w[k + 1], w[k], sh);
              asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(c4 + (x & 0x3FCu)) : "memory");
            }
          }
        }
      }
    }
  }
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
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t c4_addr = (uint32_t)__cvta_generic_to_shared(c4);
  const uint32_t hS_addr = (uint32_t)__cvta_generic_to_shared(hS);

  for (int i = tid; i < (S == 1 ? 256 : 256 + BINS); i += THREADS) smem[i] = 0;
  group_sync<THREADS>();

  while (true) {
    if (tid == 0) s_item = atomicAdd(next_item, 1u);
    group_sync<THREADS>();
    const uint32_t it = s_item;
    if (it >= n_items) break;
    const Item item = items[it];
    const uint32_t nb = item.n_blocks;
    const bool last_chunk = (item.flags & ITEM_LAST) != 0;
    const uint32_t slot = item.flags & ITEM_SLOT_MASK;
    const uint32_t first_mod3 = (item.flags >> 29) & 3u;
    const uint4* cp = codes + item.blk_begin;
    const uint64_t* vp = valid + item.blk_begin;

    for (uint32_t base = 0; base < nb; base += THREADS) {
      const uint32_t b = base + tid;
      const bool act = b < nb;
      uint4 c = make_uint4(0, 0, 0, 0);
      uint64_t V = 0;
      if (act) { c = __ldg(cp + b); V = __ldg(vp + b); }
      uint32_t hw = __shfl_down_sync(0xffffffffu, c.x, 1);
      uint32_t hv = __shfl_down_sync(0xffffffffu, (uint32_t)(V >> 32), 1);
      if (act && (lane == 31 || b + 1 == nb)) {
        const bool has = (b + 1 < nb) || !last_chunk;
        hw = has ? __ldg(reinterpret_cast<const uint32_t*>(cp + b + 1)) : 0u;
        hv = has ? (uint32_t)(__ldg(vp + b + 1) >> 32) : 0u;
      }
      if (act) {
        if (S == 3) {
          // anchors are the positions P = 0 mod 3 counted from the scaffold start; rotate the block so
          // that its first anchor sits at bit position 0
          const uint32_t bs = (first_mod3 + b) % 3u;          // (block index within scaffold) mod 3
          const uint32_t p0 = (3u - bs) % 3u;                 // 64 = 1 mod 3
          const uint32_t r2 = 2 * p0;
          uint32_t w0 = __funnelshift_l(c.y, c.x, r2), w1 = __funnelshift_l(c.z, c.y, r2);
          uint32_t w2 = __funnelshift_l(c.w, c.z, r2), w3 = __funnelshift_l(hw, c.w, r2);
          uint32_t w4 = hw << r2;
          uint64_t V2 = (V << p0) | (uint64_t)(p0 ? (hv >> (32 - p0)) : 0u);
          uint32_t hv2 = hv << p0;
          process_block<S>(w0, w1, w2, w3, w4, V2, hv2, p0 == 0 ? 22 : 21, hS_addr, c4_addr);
        } else {
          process_block<S>(c.x, c.y, c.z, c.w, hw, V, hv, 64 / S, hS_addr, c4_addr);
        }
      }
    }
    group_sync<THREADS>();

    if (S > 1) {
      // fold the (3+S)-mer histogram into the 256 tetramer codes, re-zeroing it on the way
      for (int x = tid; x < BINS; x += THREADS) {
        uint32_t v = hS[x];
        if (v) {
          hS[x] = 0;
          #pragma unroll
          for (int o = 0; o < S; ++o) atomicAdd(&c4[(x >> (2 * (S - 1 - o))) & 255], v);
        }
      }
      group_sync<THREADS>();
    }

    // total, then canonical fold 256 -> 136 and output
    uint32_t part = 0;
    for (int i = tid; i < 256; i += THREADS) part += c4[i];
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (THREADS > 32) {
      if (lane == 0) s_total[tid >> 5] = part;
      __syncthreads();
      part = 0;
      #pragma unroll
      for (int i = 0; i < THREADS / 32; ++i) part += s_total[i];
    }
    const uint32_t total = part;

    if (slot == 0) {
      for (int col = tid; col < (int)freq_stride; col += THREADS) {
        uint32_t n = 0;
        if (col < DBB_NUM_KMERS) {
          n = c4[c_col_code1[col]];
          const int c2 = c_col_code2[col];
          if (c2 >= 0) n += c4[c2];
          if (counts) counts[(size_t)item.scaffold * DBB_NUM_KMERS + col] = n;
        }
        if (freq)
          freq[(size_t)item.scaffold * freq_stride + col] =
              col < DBB_NUM_KMERS ? (float)((double)n / (double)total) : 0.f;
      }
    } else {
      for (int col = tid; col < DBB_NUM_KMERS; col += THREADS) {
        uint32_t n = c4[c_col_code1[col]];
        const int c2 = c_col_code2[col];
        if (c2 >= 0) n += c4[c2];
        if (n) atomicAdd(&split_counts[(size_t)(slot - 1) * DBB_NUM_KMERS + col], n);
      }
    }
    group_sync<THREADS>();
    for (int i = tid; i < 256; i += THREADS) c4[i] = 0;
    group_sync<THREADS>();
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
  uint32_t* d_counters = nullptr;       // [4] dynamic work counters
  uint32_t n_split = 0;
  uint32_t* d_split_scaffold = nullptr;
  uint32_t* d_split_counts = nullptr;
  int sm_count = 0;
  int launches = 0;
};

namespace {

bool g_tables_uploaded = false;

int upload_tables() {
  // per device would be more general; one device per process is the deployment model (one rank per GPU)
  if (g_tables_uploaded) return DBB_OK;
  const dbb::KmerTable& t = dbb::kmer_table();
  DBB_CUDA_TRY(cudaMemcpyToSymbol(c_col_code1, t.col_code1, sizeof(t.col_code1)));
  DBB_CUDA_TRY(cudaMemcpyToSymbol(c_col_code2, t.col_code2, sizeof(t.col_code2)));
  g_tables_uploaded = true;
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
  uint32_t grid = (uint32_t)std::min<int64_t>((int64_t)occ * plan->sm_count, (int64_t)n);
  kern<<<grid, THREADS, smem, st>>>(plan->d_items + plan->class_begin[cls], n, plan->d_counters + cls,
                                    (const uint4*)d_codes, (const uint64_t*)d_valid, d_counts, d_freq,
                                    freq_stride, plan->d_split_counts);
  DBB_CUDA_TRY(cudaGetLastError());
  return DBB_OK;
}

}  // namespace

DBB_EXPORT int dbb_count_plan_create(const int64_t* h_blk_off, int64_t n_seq, dbb_count_plan** out) {
  DBB_REQUIRE(out != nullptr && n_seq >= 0 && (n_seq == 0 || h_blk_off != nullptr), "bad arguments");
  DBB_REQUIRE(n_seq < (1ll << 32) - 1, "too many scaffolds");
  *out = nullptr;
  int rc = upload_tables();
  if (rc != DBB_OK) return rc;
  dbb_count_plan* plan = new dbb_count_plan();
  plan->n_seq = n_seq;
  plan->total_blocks = n_seq ? h_blk_off[n_seq] : 0;
  if (plan->total_blocks >= (1ll << 32)) { delete plan; dbb::set_error("batch exceeds 2^32 blocks"); return DBB_ERR_INVALID; }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, dev);

  // class thresholds in blocks (64 bases); env overrides are for tuning runs only
  const int64_t t1 = env_i64("DBB_K1_T1", 40);          // <= 2.5 kb  : S = 1, one warp per scaffold
  const int64_t t2 = env_i64("DBB_K1_T2", 320);         // <= 20 kb   : S = 2
  const int64_t t3 = env_i64("DBB_K1_T3", 3072);        // <= 196 kb  : S = 3
  const int64_t chunk = env_i64("DBB_K1_CHUNK", 16384); // S = 4 chunk: 1 Mb
  std::vector<Item> cls[4];
  std::vector<uint32_t> split_scaffold;
  for (int64_t s = 0; s < n_seq; ++s) {
    const int64_t b0 = h_blk_off[s], nb = h_blk_off[s + 1] - b0;
    if (nb < 0) { delete plan; dbb::set_error("block offsets must be non-decreasing"); return DBB_ERR_INVALID; }
    if (nb <= chunk) {
      int c = nb <= t1 ? 0 : nb <= t2 ? 1 : nb <= t3 ? 2 : 3;
      cls[c].push_back(Item{(uint32_t)b0, (uint32_t)nb, (uint32_t)s, ITEM_LAST});
    } else {
      const uint32_t slot = (uint32_t)split_scaffold.size() + 1;
      if (slot >= ITEM_SLOT_MASK) { delete plan; dbb::set_error("too many split scaffolds"); return DBB_ERR_INVALID; }
      split_scaffold.push_back((uint32_t)s);
      for (int64_t o = 0; o < nb; o += chunk) {
        const int64_t len = std::min(chunk, nb - o);
        uint32_t flags = slot | (uint32_t)((o % 3) << 29) | (o + len == nb ? ITEM_LAST : 0u);
        cls[3].push_back(Item{(uint32_t)(b0 + o), (uint32_t)len, (uint32_t)s, flags});
      }
    }
  }
  std::vector<Item> all;
  all.reserve((size_t)n_seq + 16);
  for (int c = 0; c < 4; ++c) {
    std::stable_sort(cls[c].begin(), cls[c].end(), [](const Item& a, const Item& b) { return a.n_blocks > b.n_blocks; });
    plan->class_begin[c] = (uint32_t)all.size();
    plan->class_count[c] = (uint32_t)cls[c].size();
    all.insert(all.end(), cls[c].begin(), cls[c].end());
  }
  plan->n_split = (uint32_t)split_scaffold.size();
  plan->launches = 0;
  for (int c = 0; c < 4; ++c) plan->launches += plan->class_count[c] ? 1 : 0;
  if (plan->n_split) plan->launches += 1;

  cudaError_t e = cudaSuccess;
  if (!all.empty()) {
    e = cudaMalloc(&plan->d_items, all.size() * sizeof(Item));
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_items, all.data(), all.size() * sizeof(Item), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&plan->d_counters, 4 * sizeof(uint32_t));
  if (e == cudaSuccess && plan->n_split) {
    e = cudaMalloc(&plan->d_split_scaffold, plan->n_split * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_split_scaffold, split_scaffold.data(), plan->n_split * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&plan->d_split_counts, (size_t)plan->n_split * DBB_NUM_KMERS * sizeof(uint32_t));
  }
  if (e != cudaSuccess) {
    dbb::set_error("count plan allocation failed: %s", cudaGetErrorString(e));
    dbb_count_plan_destroy(plan);
    return e == cudaErrorMemoryAllocation ? DBB_ERR_NOMEM : DBB_ERR_CUDA;
  }
  *out = plan;
  return DBB_OK;
}

DBB_EXPORT int dbb_count_plan_destroy(dbb_count_plan* plan) {
  if (!plan) return DBB_OK;
  cudaFree(plan->d_items);
  cudaFree(plan->d_counters);
  cudaFree(plan->d_split_scaffold);
  cudaFree(plan->d_split_counts);
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
  int rc;
  // longest-running class first
  if ((rc = launch_class<4, 512>(plan, 3, d_codes, d_valid, d_counts, d_freq, freq_stride, st)) != DBB_OK) return rc;
  if ((rc = launch_class<3, 256>(plan, 2, d_codes, d_valid, d_counts, d_freq, freq_stride, st)) != DBB_OK) return rc;
  if ((rc = launch_class<2, 64>(plan, 1, d_codes, d_valid, d_counts, d_freq, freq_stride, st)) != DBB_OK) return rc;
  if ((rc = launch_class<1, 32>(plan, 0, d_codes, d_valid, d_counts, d_freq, freq_stride, st)) != DBB_OK) return rc;
  if (plan->n_split) {
    finalize_kernel<<<plan->n_split, 160, 0, st>>>(plan->d_split_counts, plan->d_split_scaffold, plan->n_split,
                                                   d_counts, d_freq, freq_stride);
    DBB_CUDA_TRY(cudaGetLastError());
  }
  return DBB_OK;
}
