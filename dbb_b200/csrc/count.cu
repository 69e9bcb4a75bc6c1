// Stage 1 (K1): canonical tetranucleotide counting + normalisation on packed sequence.
//
// Reference semantics (genomicSignatures.py:134-150): slide a 4-base window with stride 1, skip any
// window that contains a byte outside {A,C,G,T}, increment the canonical column, then divide by the
// sum (0/0 -> NaN).  The integer counts must be bit-exact.
//
// Hardware facts this design follows (profiles/microbench/mb2.cu, mb4.cu, measured on B200):
//   * the SM's shared-memory pipe retires one wavefront per clock; a shared-memory atomic of 32 random bins costs
//     ~3-3.7 wavefronts (bank conflicts of random addresses), so one atomic per WINDOW caps counting near 9-12
//     windows / clk / SM, a quarter of what the HBM target needs;
//   * so one atomic covers S windows: a (3+S)-mer histogram (4^(3+S) bins) is updated once every S bases
//     ("super-k-mer") and folded to the 256 tetramer codes when the scaffold ends.  The fold is exact integer
//     arithmetic:  count4[c] = sum over window offsets o < S of the marginal of the (3+S)-mer histogram.
//     Measured ceilings without the fold (mb4, 16-bit counters): 17 / 26 / 35 / 38 windows / clk / SM for S = 2..5 on
//     uniform sequence, 14 / 22 / 31 / 33 on a 30 % GC genome (skewed bins collide more);
//   * the fold costs O(4^(3+S)) wavefronts per scaffold, so S is chosen per scaffold from its length.
//
// Round 2 layout: ONE WARP owns one work item and one histogram; nothing in the kernel synchronises wider than a
// warp, so a warp that folds or writes its outputs never stalls the atomics of the others (round 1 used one CTA per
// item and lost 3-5 issue slots per instruction to barriers around the fold).  Counters are 16 bits wide, two per
// 32-bit word (word = bin without its top bit, half = the top bit): the histogram of S = 4 is 32 KB instead of 64 KB,
// six of them fit an SM, every fold reads and re-zeroes half the bytes; an item is at most 65 535 super-k-mers
// (4 095 blocks for S = 4), so no counter can wrap.
//
// Work decomposition: a work ITEM is a run of 64-base blocks of one scaffold (whole scaffold, or a chunk of a very
// long one).  Lane i of the warp loads block i of the current 32-block stripe with a single 128-bit load (512
// contiguous bytes per warp), the halo of the next block arrives by warp shuffle, loads run DEPTH stripes ahead of
// the atomics.  Windows whose super-k-mer is only partly valid (N, lowercase, scaffold end) take a rare per-window
// path.  Whole-scaffold items write counts and frequencies directly; chunks of split scaffolds add into a scratch row
// that a finalize kernel normalises.
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

template <int S> struct Geo {
  static constexpr int R = S + 3;                       // bases per super-k-mer
  static constexpr int KB = 2 * R;                      // bits per bin index
  static constexpr int BINS = 1 << KB;
  static constexpr int WORDS = S == 1 ? 0 : BINS / 2;   // 16-bit counters, two per word (S = 1 counts straight into c4)
};

// All shared memory of the kernel is the dynamic array below, addressed by WORD OFFSETS (not by derived pointers:
// a generic pointer into shared memory costs an S2UR + ULEA to build every time it is materialised, which was 30 % of
// the warp time of the short-contig class).  Layout per CTA: c4[256] u32 | histogram[WORDS] | column table[160] |
// WPI x slow-path scratch[16].
extern __shared__ __align__(16) uint32_t smem[];
constexpr int OFF_C4 = 0;
constexpr int OFF_HS = 256;

// Shared-memory accesses by 32-bit shared-space BYTE address (the address of smem[] is taken once per thread; the
// compiler otherwise rebuilds the CTA's shared window -- S2UR SR_CgaCtaId + ULEA -- in every basic block that
// touches shared memory).
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128_zero(uint32_t a) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
}
__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }

// one update of the (3+S)-mer histogram; x holds the bin in bits [2, 2 + KB)
template <int S>
__device__ __forceinline__ void bump(uint32_t sb, uint32_t x) {
  constexpr int KB = Geo<S>::KB;
  if (S == 1) {
    red_inc(sb + (x & (255u << 2)));
  } else {
    constexpr uint32_t WMASK4 = ((1u << (KB - 1)) - 1u) << 2;     // byte offset of the word: the bin without its top bit
    constexpr uint32_t TOP = 1u << (KB + 1);                       // top bit of the bin selects the half
    red_add(sb + 4 * OFF_HS + (x & WMASK4), (x & TOP) ? 0x10000u : 1u);
  }
}

// Rare path: blocks that are not fully valid (N / IUPAC / lowercase, the padded last block of a scaffold, a
// block whose halo is invalid).  The warp handles them one at a time and COOPERATIVELY: the owning lane
// publishes the block through a 10-word shared scratch, lane a then treats anchor a (a, a+32 for S = 1).  A
// fully valid anchor still updates the (3+S)-mer histogram; otherwise its valid windows go one by one to the
// tetramer histogram.
template <int S>
__device__ __noinline__ void slow_blocks(uint32_t slowmask, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                         uint32_t w4, uint64_t V, uint32_t Vh, int n_anchor, uint32_t sb, int off_sc, int lane) {
  constexpr int R = Geo<S>::R;
  constexpr int KB = Geo<S>::KB;
  const uint32_t sc = sb + 4 * off_sc;
  while (slowmask) {
    const int src = __ffs(slowmask) - 1;
    slowmask &= slowmask - 1;
    if (lane == src) {
      sts(sc, w0); sts(sc + 4, w1); sts(sc + 8, w2); sts(sc + 12, w3); sts(sc + 16, w4); sts(sc + 20, 0u);
      sts(sc + 24, (uint32_t)(V >> 32)); sts(sc + 28, (uint32_t)V); sts(sc + 32, Vh); sts(sc + 36, (uint32_t)n_anchor);
    }
    __syncwarp();
    const uint64_t V64 = ((uint64_t)lds(sc + 24) << 32) | lds(sc + 28);
    const uint32_t vh = lds(sc + 32);
    const int na = (int)lds(sc + 36);
    for (int a = lane; a < na; a += 32) {
      const int p = a * S;
      // validity of positions p, p+1, ... in the top bits of x
      const uint64_t x = (V64 << p) | (p ? (((uint64_t)vh << 32) >> (64 - p)) : 0ull);
      const int k = p >> 4, o = p & 15;
      if ((x >> (64 - R)) == ((1ull << R) - 1ull)) {
        const uint64_t pair = ((uint64_t)lds(sc + 4 * k) << 32) | lds(sc + 4 * k + 4);
        bump<S>(sb, ((uint32_t)(pair >> (64 - 2 * o - KB)) & ((1u << KB) - 1u)) << 2);
      } else {
        #pragma unroll
        for (int o2 = 0; o2 < S; ++o2) {
          if (((x >> (60 - o2)) & 15ull) == 15ull) {
            const int q = p + o2, k2 = q >> 4, oq = q & 15;
            const uint64_t pair = ((uint64_t)lds(sc + 4 * k2) << 32) | lds(sc + 4 * k2 + 4);
            red_inc(sb + 4 * (OFF_C4 + ((uint32_t)(pair >> (64 - 2 * oq - 8)) & 255u)));
          }
        }
      }
    }
    __syncwarp();
  }
}

// Fast path of one 64-base block whose bases (and the R-1 halo bases) are all valid: one unpredicated
// shared-memory atomic per anchor.
//   w0..w3: code words of the block, w4: first code word of the next block (halo)
//   n_anchor: super-k-mers anchored in this block (64/S; 21 or 22 for S = 3)
template <int S>
__device__ __forceinline__ void fast_block(uint32_t sb, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, int n_anchor) {
  constexpr int KB = Geo<S>::KB;
  constexpr int NA = (64 + S - 1) / S;                 // 64, 32, 22, 16
  const uint32_t w[5] = {w0, w1, w2, w3, w4};
  #pragma unroll
  for (int p = 0; p < 64; p += S) {
    const int k = p >> 4, o = p & 15;
    const int sh = 64 - 2 * o - KB - 2;                // >= 18
    const uint32_t x = (sh >= 32) ? (w[k] >> (sh - 32)) : __funnelshift_r(w[k + 1], w[k], sh);
    if (S == 3 && p == 63 && n_anchor != NA) continue; // S = 3: the anchor at 63 may belong to the next block
    bump<S>(sb, x);
  }
}

template <int S>
__device__ __forceinline__ bool block_all_valid(uint64_t V, uint32_t Vh) {
  constexpr uint32_t HALO = ~0u << (32 - (S + 2));     // the R-1 halo bases an anchor of this block may touch
  return (V == ~0ull) && ((Vh & HALO) == HALO);
}

__device__ __forceinline__ uint32_t halves(uint32_t v) { return (v & 0xffffu) + (v >> 16); }
__device__ __forceinline__ uint32_t sum4(const uint4& v) { return (v.x + v.y) + (v.z + v.w); }   // packed: no half can wrap

// Fold of the CTA's (3+S)-mer histogram into its 256 tetramer codes.  Thread <-> code: warp w takes the codes
// c = lane + 32 q of its q's; no shuffles.  Bin = (b0 .. b_{R-1}), b0 most significant; word = bin without its top bit,
// low / high half = top bit 0 / 1.  For window offset o the count of code c = (b_o .. b_{o+3}) is the sum over the o
// bases before it and the e = S-1-o bases after it:
//   o = 0 : the top bit belongs to the code -- the codes c and c | 128 share their words (low / high halves); a run of
//           4^e consecutive words each;
//   o >= 1: the top bit belongs to the summed prefix -- both halves of every word are added; 4^o / 2 runs of 4^e words.
// Runs of >= 4 words are read as 128-bit words in an order rotated by the lane so that the eight lanes of a
// 128-bit wavefront hit eight different 16-byte bank groups.
template <int S, int WPI>
__device__ __forceinline__ void fold_histogram(uint32_t sb, int lane, int warp) {
  #pragma unroll
  for (int o = 0; o < S; ++o) {
    const int e = S - 1 - o;                    // bases after the window
    const int run = 1 << (2 * e);               // consecutive words per run
    const int nq = run >> 2;                    // 128-bit chunks per run (0: single words)
    if (o == 0) {
      #pragma unroll 1
      for (int q = warp; q < 4; q += WPI) {
        const int c = lane + 32 * q;            // < 128: low halves -> code c, high halves -> code c | 128
        const int base = OFF_HS + (c << (2 * e));
        uint32_t s = 0;
        if (nq == 0) {
          s = lds(sb + 4 * base);
        } else {
          #pragma unroll
          for (int r4 = 0; r4 < nq; ++r4) {
            const int ch = nq == 4 ? ((r4 + (lane >> 1)) & 3) : ((r4 + lane) & (nq - 1));
            s += sum4(lds128(sb + 4 * (base + 4 * ch)));
          }
        }
        red_add(sb + 4 * (OFF_C4 + c), s & 0xffffu);
        red_add(sb + 4 * (OFF_C4 + c + 128), s >> 16);
      }
    } else {
      const int npfx = 1 << (2 * o - 1);        // prefix values without the top bit
      #pragma unroll 1
      for (int q = warp; q < 8; q += WPI) {
        const int c = lane + 32 * q;
        uint32_t s = 0;
        #pragma unroll 4
        for (int pfx = 0; pfx < npfx; ++pfx) {
          const int base = OFF_HS + ((pfx << (8 + 2 * e)) | (c << (2 * e)));
          if (nq == 0) {
            s += lds(sb + 4 * base);
          } else {
            #pragma unroll
            for (int r4 = 0; r4 < nq; ++r4) {
              const int ch = nq == 4 ? ((r4 + (lane >> 1)) & 3) : ((r4 + lane) & (nq - 1));
              s += sum4(lds128(sb + 4 * (base + 4 * ch)));
            }
          }
        }
        red_add(sb + 4 * (OFF_C4 + c), halves(s));
      }
    }
  }
}

// Fold of the 7-mer histogram (S = 4) in TWO read passes instead of four (the generic fold above reads the whole
// histogram once per window offset: 4 x 32 KB = 1 024 wavefronts, 35 % of the shared-memory pipe of a 96 kb scaffold):
//   pass "suffix": thread <-> t = (b3 b4 b5 b6), the low 8 bits of the word index; the 32 words (b0_lo b1 b2 | t) are 32
//     conflict-free 32-bit loads.  Offset 3 (code t) is the sum of all of them, offset 2 (code b2 b3 b4 b5) sums over
//     b0, b1 in the thread and over b6 across the four neighbouring lanes;
//   pass "prefix": thread <-> rho = (b0_lo b1 b2 b3 b4), a run of 16 consecutive words (b5 b6) read as four rotated
//     128-bit loads.  Offset 1 (code b1 b2 b3 b4) adds both halves (b0_hi), b0_lo arrives from the other rho through the
//     atomic; offset 0 (codes c and c | 128 in the low / high halves) sums over b4 across the four neighbouring lanes.
// Every bin is read exactly twice: 512 wavefronts + ~100 for shuffles and the updates of c4.  Read-only (the caller
// re-zeroes the histogram behind a barrier, as for the other S).
template <int WPI>
__device__ __forceinline__ void fold_histogram4(uint32_t sb, int tid, int lane) {
  constexpr int NT = 32 * WPI;
  static_assert(256 % NT == 0, "whole warps must run every iteration of the loops below (they shuffle)");
  const uint32_t hs = sb + 4 * OFF_HS;
  #pragma unroll 1
  for (int t = tid; t < 256; t += NT) {
    uint32_t m3 = 0, m2[4] = {0u, 0u, 0u, 0u};
    #pragma unroll
    for (int it = 0; it < 32; ++it) {               // it = (b0_lo b1 b2)
      const uint32_t sv = halves(lds(hs + 4 * ((it << 8) | t)));
      m3 += sv;
      m2[it & 3] += sv;
    }
    red_add(sb + 4 * (OFF_C4 + t), m3);
    #pragma unroll
    for (int j = 0; j < 4; ++j) {                   // sum over b6 = the two low lane bits (t = tid + i NT: lane bits = t bits)
      m2[j] += __shfl_xor_sync(0xffffffffu, m2[j], 1);
      m2[j] += __shfl_xor_sync(0xffffffffu, m2[j], 2);
    }
    const int j = t & 3;                            // lane j of the four publishes b2 = j
    const uint32_t v = j == 0 ? m2[0] : j == 1 ? m2[1] : j == 2 ? m2[2] : m2[3];
    red_add(sb + 4 * (OFF_C4 + ((j << 6) | (t >> 2))), v);
  }
  #pragma unroll 1
  for (int rho = tid; rho < 512; rho += NT) {
    uint32_t a = 0;                                 // packed halves: b0_hi = 0 / 1 (no half can wrap: an item is <= 65 535 super-k-mers)
    #pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ch = (r + (lane >> 1)) & 3;         // the eight lanes of a 128-bit wavefront hit eight bank groups
      a += sum4(lds128(hs + 4 * (rho * 16 + 4 * ch)));
    }
    red_add(sb + 4 * (OFF_C4 + (rho & 255)), halves(a));
    a += __shfl_xor_sync(0xffffffffu, a, 1);        // sum over b4 = the two low lane bits
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if ((rho & 3) == 0) {
      red_add(sb + 4 * (OFF_C4 + (rho >> 2)), a & 0xffffu);
      red_add(sb + 4 * (OFF_C4 + (rho >> 2) + 128), a >> 16);
    }
  }
}

template <int WPI>
__device__ __forceinline__ void group_sync() {
  if (WPI == 1) __syncwarp(); else __syncthreads();
}

// this lane's block of a stripe, loaded one stripe ahead of its use
struct Fetch { uint4 c; uint64_t V; uint32_t hw, hv; };

template <int S, int WPI, int MINB>
__global__ void __launch_bounds__(32 * WPI, MINB) count_kernel(const Item* __restrict__ items, uint32_t n_items,
                                                               const uint4* __restrict__ codes,
                                                               const uint64_t* __restrict__ valid,
                                                               uint32_t* __restrict__ counts,
                                                               float* __restrict__ freq, int64_t freq_stride,
                                                               uint32_t* __restrict__ split_counts) {
  // One CTA = the WPI warps that share one item and one histogram (a warp is latency-bound around its atomics, so a
  // histogram needs several warps to keep the shared-memory pipe fed; the warps only meet at the end of an item).
  constexpr int WORDS = Geo<S>::WORDS;
  constexpr int OFF_CC = 256 + WORDS;                     // column table: the (one or two) codes of column i, 16 bits each
  constexpr int OFF_SC = OFF_CC + 160;                    // slow-path scratch, 16 words per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < OFF_SC + 16 * WPI; i += 32 * WPI) {
    uint32_t v = 0;
    const int col = i - OFF_CC;
    if (col >= 0 && col < 160)
      v = col < DBB_NUM_KMERS ? ((uint32_t)(uint16_t)c_col_code1[col] | ((uint32_t)(uint16_t)c_col_code2[col] << 16)) : 0xffffffffu;
    sts(sb + 4 * i, v);
  }
  group_sync<WPI>();

  // Static schedule: CTA g takes items g, g + G, ... of a list sorted by descending size, so every load below is
  // known one step ahead: the next stripe (of this item or of the next one) is requested before the current one is
  // processed, the record of the item after the next when an item begins.
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
  const uint32_t tid = threadIdx.x;
  uint32_t it = blockIdx.x;
  Item item = Item{0, 0, 0, 0};
  Fetch nxt;
  nxt.c = make_uint4(0, 0, 0, 0); nxt.V = 0; nxt.hw = 0; nxt.hv = 0;
  if (it < n_items) { item = items[it]; nxt = fetch(item, tid); }

  while (it < n_items) {
    const uint32_t it_n = it + gridDim.x;
    const bool has_n = it_n < n_items && it_n > it;
    Item item_n = Item{0, 0, 0, 0};
    if (has_n) item_n = items[it_n];
    const uint32_t nb = item.n_blocks;
    const uint32_t first_mod3 = (item.flags >> 29) & 3u;

    for (uint32_t base = 0; base < nb || base == 0; base += 32 * WPI) {
      const uint32_t b = base + tid;
      const bool act = b < nb;
      const Fetch cur = nxt;
      if (base + 32 * WPI < nb) nxt = fetch(item, base + 32 * WPI + tid);
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
      if (ok) fast_block<S>(sb, w0, w1, w2, w3, w4, n_anchor);
      const uint32_t slow = __ballot_sync(0xffffffffu, act && !ok);
      if (slow) slow_blocks<S>(slow, w0, w1, w2, w3, w4, Vb, hvb, n_anchor, sb, OFF_SC + 16 * warp, lane);
    }
    group_sync<WPI>();
    if (c_debug_skip_flush) { it = it_n; item = item_n; if (!has_n) break; continue; }

    if (S > 1) {
      if (S == 4) fold_histogram4<WPI>(sb, (int)tid, lane);
      else fold_histogram<S, WPI>(sb, lane, warp);
      group_sync<WPI>();                                    // every marginal is in c4, nobody reads the histogram any more
      for (int x = tid * 4; x < WORDS; x += 128 * WPI) sts128_zero(sb + 4 * (OFF_HS + x));
    }

    // canonical fold 256 -> 136 by warp 0, total = sum of the folded counts, one reciprocal per scaffold:
    // freq = (float)(n * (1.0 / total))  (0 * inf = NaN when nothing was counted, as genomicSignatures.py:147-148 yields)
    constexpr int NCOLT = (DBB_SIG_STRIDE + 31) / 32;       // output columns per lane (136 + 4 padding)
    uint32_t nloc[NCOLT];
    uint32_t part = 0;
    if (warp == 0) {
      #pragma unroll
      for (int q = 0; q < NCOLT; ++q) {
        const uint32_t cc = lds(sb + 4 * (OFF_CC + lane + 32 * q));
        uint32_t nq = 0;
        if (cc != 0xffffffffu) { nq = lds(sb + 4 * (OFF_C4 + (cc & 0xffffu))); if (!(cc >> 31)) nq += lds(sb + 4 * (OFF_C4 + (cc >> 16))); }
        nloc[q] = nq;
        part += nq;
      }
      #pragma unroll
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
      __syncwarp();
      // every lane has read its c4 entries: re-zero for the next item (its first update is behind the sync below)
      #pragma unroll
      for (int i = 0; i < 2; ++i) sts128_zero(sb + 4 * (OFF_C4 + 4 * lane + 128 * i));
    }
    group_sync<WPI>();                                      // histogram and c4 are zero again: the next item may start
    if (warp == 0) {
      const uint32_t total = part;
      const uint32_t slot = item.flags & ITEM_SLOT_MASK;
      if (slot == 0) {
        const double inv = 1.0 / (double)total;
        #pragma unroll
        for (int q = 0; q < NCOLT; ++q) {
          const int col = lane + q * 32;
          if (col < DBB_NUM_KMERS && counts) counts[(size_t)item.scaffold * DBB_NUM_KMERS + col] = nloc[q];
          if (freq && col < (int)freq_stride)
            freq[(size_t)item.scaffold * freq_stride + col] = col < DBB_NUM_KMERS ? (float)((double)nloc[q] * inv) : 0.f;
        }
      } else {
        #pragma unroll
        for (int q = 0; q < NCOLT; ++q) {
          const int col = lane + q * 32;
          if (col < DBB_NUM_KMERS && nloc[q]) atomicAdd(&split_counts[(size_t)(slot - 1) * DBB_NUM_KMERS + col], nloc[q]);
        }
      }
    }
    if (!has_n) break;
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

int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoll(v) : dflt;
}

// Class limits in 64-base blocks: a scaffold of nb blocks is counted with S = 1 if nb <= limit[0], S = 2 if
// nb <= limit[1], S = 3 if nb <= limit[2], else S = 4 (whole if nb <= limit[3], otherwise in chunks of limit[3]
// blocks).  The 16-bit counters bound every item to 65 535 super-k-mers: 2 047 / 3 071 / 4 095 blocks for S = 2 / 3 / 4.
// Env overrides are for tuning runs only and are clamped to those bounds.
struct Limits { int64_t t[4]; };
const Limits& limits() {
  static const Limits L = [] {
    Limits l;
    l.t[0] = std::max<int64_t>(0, env_i64("DBB_K1_T1", 32));                                   // <= 2 kb   : S = 1
    l.t[1] = std::min<int64_t>(2047, std::max(l.t[0], env_i64("DBB_K1_T2", 512)));             // <= 33 kb  : S = 2
    l.t[2] = std::min<int64_t>(3071, std::max(l.t[1], env_i64("DBB_K1_T3", 1696)));            // <= 108 kb : S = 3
    l.t[3] = std::min<int64_t>(4095, std::max<int64_t>(l.t[2] + 1, env_i64("DBB_K1_CHUNK", 4032)));   // S = 4 chunk: 258 kb
    return l;
  }();
  return L;
}

}  // namespace

struct dbb_count_plan {
  int64_t n_seq = 0;
  int64_t total_blocks = 0;
  // per class (S = 1..4): item range in d_items
  uint32_t class_begin[4] = {0, 0, 0, 0};
  uint32_t class_count[4] = {0, 0, 0, 0};
  uint32_t class_grid[4] = {0, 0, 0, 0};       // CTAs of the class kernel: CTA g takes items g, g + grid, ... of its class list
  Item* d_items = nullptr;
  size_t cap_items = 0;
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

// The kernel instantiation of a class: warps per item (= per CTA) and minimum CTAs per SM are the measured best
// (profiles/r02_k1_classes_by_length.txt; 16 instead of 12 CTAs per SM, i.e. 62 instead of 77 registers: S = 2 class +1.5 %,
// S = 3 class -2 % on C2); DBB_K1_WPI<S> / DBB_K1_MINB<S> overrides exist for tuning sweeps.
struct ClassCfg { const void* fn; int threads; size_t smem; int S; };
template <int S, int WPI, int MINB>
ClassCfg make_cfg() {
  return ClassCfg{(const void*)count_kernel<S, WPI, MINB>, 32 * WPI, (size_t)(256 + Geo<S>::WORDS + 160 + WPI * 16) * sizeof(uint32_t), S};
}
const ClassCfg& class_cfg(int c) {
  static const ClassCfg cfg[4] = {
      make_cfg<1, 1, 16>(),
      env_i64("DBB_K1_WPI2", 2) >= 2 ? (env_i64("DBB_K1_MINB2", 16) >= 16 ? make_cfg<2, 2, 16>() : make_cfg<2, 2, 12>()) : make_cfg<2, 1, 16>(),
      env_i64("DBB_K1_WPI3", 2) >= 4 ? make_cfg<3, 4, 8>() : env_i64("DBB_K1_WPI3", 2) >= 2 ? (env_i64("DBB_K1_MINB3", 12) >= 16 ? make_cfg<3, 2, 16>() : make_cfg<3, 2, 12>()) : make_cfg<3, 1, 16>(),
      env_i64("DBB_K1_WPI4", 8) >= 8 ? make_cfg<4, 8, 6>() : env_i64("DBB_K1_WPI4", 8) >= 4 ? make_cfg<4, 4, 6>() : make_cfg<4, 2, 6>()};
  return cfg[c];
}

// CTAs per SM of a class kernel on the current device (0: does not fit); DBB_K1_OCC<S> caps it (tuning switch)
int class_occupancy(int c) {
  const ClassCfg& k = class_cfg(c);
  if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem) != cudaSuccess) return 0;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.threads, k.smem) != cudaSuccess) return 0;
  static const int64_t cap[4] = {env_i64("DBB_K1_OCC1", 0), env_i64("DBB_K1_OCC2", 0), env_i64("DBB_K1_OCC3", 0), env_i64("DBB_K1_OCC4", 0)};
  if (cap[c] > 0 && cap[c] < occ) occ = (int)cap[c];
  return occ;
}

int launch_class(const dbb_count_plan* plan, int cls, const void* d_codes, const void* d_valid,
                 uint32_t* d_counts, float* d_freq, int64_t freq_stride, cudaStream_t st) {
  uint32_t n = plan->class_count[cls];
  if (n == 0) return DBB_OK;
  const ClassCfg& k = class_cfg(cls);
  const Item* items = plan->d_items + plan->class_begin[cls];
  const uint4* codes = (const uint4*)d_codes;
  const uint64_t* valid = (const uint64_t*)d_valid;
  uint32_t* split_counts = plan->d_split_counts;
  void* args[] = {&items, &n, &codes, &valid, &d_counts, &d_freq, &freq_stride, &split_counts};
  DBB_CUDA_TRY(cudaLaunchKernel(k.fn, dim3(plan->class_grid[cls]), dim3(k.threads), args, k.smem, st));
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

  const Limits& lim = limits();
  const int64_t t1 = lim.t[0], t2 = lim.t[1], t3 = lim.t[2], chunk = lim.t[3];
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
    plan->class_grid[c] = 0;
    if (!cls[c].empty()) {
      const int occ = class_occupancy(c);
      if (occ < 1) { dbb::set_error("count kernel S=%d does not fit on the device", c + 1); return DBB_ERR_CUDA; }
      plan->class_grid[c] = (uint32_t)std::min<int64_t>((int64_t)occ * plan->sm_count, (int64_t)cls[c].size());
    }
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

DBB_EXPORT int dbb_count_class_limits(int64_t* limits4) {
  DBB_REQUIRE(limits4 != nullptr, "NULL argument");
  for (int i = 0; i < 4; ++i) limits4[i] = limits().t[i];
  return DBB_OK;
}

DBB_EXPORT int dbb_count_plan_create(const int64_t* h_blk_off, int64_t n_seq, dbb_count_plan** out) {
  DBB_REQUIRE(out != nullptr, "bad arguments");
  *out = nullptr;
  int rc = upload_tables();
  if (rc != DBB_OK) return rc;
  dbb_count_plan* plan = new dbb_count_plan();
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming);
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
  if (plan->n_split)
    DBB_CUDA_TRY(cudaMemsetAsync(plan->d_split_counts, 0, (size_t)plan->n_split * DBB_NUM_KMERS * sizeof(uint32_t), st));
  // fork: every class kernel on its own stream behind the memset, join back into the caller's stream
  DBB_CUDA_TRY(cudaEventRecord(plan->ev_fork, st));
  int rc = DBB_OK;
  for (int c = 3; c >= 0 && rc == DBB_OK; --c) {            // longest-running class first
    if (plan->class_count[c] == 0) continue;
    cudaStream_t cs = plan->class_stream[c];
    DBB_CUDA_TRY(cudaStreamWaitEvent(cs, plan->ev_fork, 0));
    rc = launch_class(plan, c, d_codes, d_valid, d_counts, d_freq, freq_stride, cs);
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
