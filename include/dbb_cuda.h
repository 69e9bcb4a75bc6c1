/*
 * dbb_cuda.h -- C ABI of libdbbcuda.so: DBB's data-parallel hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (donovan-h-parks/DBB) is pure Python with no FFI layer; the "plugin boundary" for this
 * path is the method surface of four modules.  Each entry point below names the reference code it
 * replaces (file:line into /root/reference/dbb).  The Python mirror of those modules lives in
 * dbb_b200/ and binds these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns int: DBB_OK (0) or a negative DBB_ERR_*; never throws, never aborts;
 *     dbb_last_error() returns a thread-local message for the last failure.
 *   - "_dev" functions take DEVICE pointers and a cudaStream_t (as void*; NULL = default stream),
 *     enqueue work on the CURRENT device and return without synchronising.
 *   - "_host" functions take HOST pointers, do their own H2D/D2H and return when results are in the
 *     output buffers.  The caller owns every buffer passed in; the library owns only its scratch.
 *   - no CPU fallback exists: without a CUDA device every compute entry point returns DBB_ERR_CUDA.
 *
 * Packed sequence layout: a scaffold of L bases occupies ceil(L/64) BLOCKS.
 *   codes plane : 16 bytes per block = 4 little-endian uint32 words, word w holds bases 16w..16w+15,
 *                 2 bits per base (A=0 C=1 G=2 T=3), FIRST base in the MOST significant bits.
 *   valid plane :  8 bytes per block = one uint64, bit (63-p) set iff base p of the block is one of
 *                 the four UPPERCASE letters A C G T (anything else, and the padding past the end of
 *                 the scaffold, is invalid -- genomicSignatures.py:139-144 skips such windows).
 *   blk_off[s]  : first block of scaffold s; blk_off[n_seq] = total blocks.  Windows never span
 *                 scaffolds.
 */
#ifndef DBB_CUDA_H
#define DBB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBB_NUM_KMERS 136      /* canonical tetramers, genomicSignatures.py:44-72 */
#define DBB_SIG_STRIDE 140     /* floats per row of the padded signature matrix used by stage 2 */
#define DBB_BLOCK_BASES 64
#define DBB_MAX_K 32

enum {
  DBB_OK = 0,
  DBB_ERR_INVALID = -1,        /* bad argument */
  DBB_ERR_CUDA = -2,           /* CUDA runtime/driver failure (message in dbb_last_error) */
  DBB_ERR_NOMEM = -3,
  DBB_ERR_UNSUPPORTED = -4     /* e.g. K != 4, k > DBB_MAX_K, device is not sm_100 */
};

const char* dbb_last_error(void);
int dbb_version(void);
/* device facts for roofline arithmetic: SM count, compute capability, free/total bytes */
int dbb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* free_bytes, int64_t* total_bytes);

/* ---- contexts (SURVEY.md section 8b "Ownership"): a context owns the stream, the device scratch and the cached
 * plans of ONE device.  Every entry point that takes a dbb_ctx* accepts NULL = the default context of the current
 * device (created on first use, alive until the process ends); the *_host entry points without a context argument
 * use that default context too.  A context must be used with its own device current (else DBB_ERR_INVALID) and by one
 * thread at a time.  device < 0 = the current device. */
typedef struct dbb_ctx dbb_ctx;
int dbb_ctx_create(int device, dbb_ctx** ctx);
int dbb_ctx_destroy(dbb_ctx* ctx);

/* ---- canonical tetramer table: genomicSignatures.py:35-84 (__init__, __makeKmerColNames,
 * ---- __lexicographicallyLowest, __revComp) ------------------------------------------------------ */
/* lut256[code] = canonical column of the 4-mer with 2-bit code `code` (first base most significant) */
int dbb_kmer_lut(uint8_t* lut256);
/* names: 136 NUL-terminated 4-letter names, 5 bytes each, in column order (canonicalKmerOrder, :128) */
int dbb_kmer_names(char* names680);

/* ---- stage 1: tetranucleotide signatures -- genomicSignatures.py:134-150 (seqSignature) and the
 * ---- fan-out of :152-181 (calculate) ---------------------------------------------------------------- */
int64_t dbb_packed_blocks(int64_t seq_len);
/* blk_off[0..n_seq] from lengths (host helper) */
int dbb_block_offsets(const int64_t* seq_len, int64_t n_seq, int64_t* blk_off);

/* ASCII -> packed planes.  d_seq_off[n_seq+1] are byte offsets of each scaffold in d_ascii. */
int dbb_pack_ascii_dev(const uint8_t* d_ascii, const int64_t* d_seq_off, const int64_t* d_blk_off,
                       int64_t n_seq, int64_t total_blocks, void* d_codes, void* d_valid, void* stream);

/* The same planes computed on the HOST (AVX2 + BMI2 when the CPU has them, a table loop otherwise; bit-identical to
 * dbb_pack_ascii_dev): all pointers are host pointers, codes = 16 bytes and valid = 8 bytes per block, scaffold s at
 * block blk_off[s].  No CUDA call is made.  dbb_tetra_signatures_host uses it on spare host threads for part of every
 * batch, because packed planes cost 0.375 bytes per base on the host-to-device link against 1 for ASCII. */
int dbb_pack_ascii_host(const uint8_t* ascii, const int64_t* seq_off, const int64_t* blk_off, int64_t n_seq,
                        void* codes, void* valid);

/* Same for sequences that are arbitrary slices ascii[begin[s], end[s]) of one buffer (contigs cut out of their
 * scaffolds, preprocess.py:158).  Any of d_codes / d_valid / d_nmask may be NULL.  d_nmask is a third bit plane
 * with the geometry of the validity plane: bit (63 - p) of block word = base p is 'N' or 'n' (splitScaffolds.py:34). */
int dbb_pack_ranges_dev(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end, const int64_t* d_blk_off,
                        int64_t n_seq, int64_t total_blocks, void* d_codes, void* d_valid, void* d_nmask, void* stream);

/* ---- N2: sequence side of the preprocess worker (preprocess.py:118-161) ------------------------------------
 * dbb_split_scaffolds_dev replaces SplitScaffolds.splits (splitScaffolds.py:27-53) for a batch: scaffold s is cut at
 * every run of N/n LONGER than min_n that is followed by another base.  Two passes over the N plane of
 * dbb_pack_ranges_dev: with d_contig_off == NULL it writes the number of contigs per scaffold to d_n_contigs[n_seq];
 * with d_contig_off[n_seq] (exclusive prefix sum of those counts) it writes d_begin / d_end (positions relative to
 * the scaffold start, contig k of scaffold s at d_contig_off[s] + k = the reference's '<id>_c<k>').  The reference's
 * edge behaviour is kept (empty leading contig, trimmed trailing run, dropped one-base tail). */
int dbb_split_scaffolds_dev(const void* d_nmask, const int64_t* d_seq_len, const int64_t* d_blk_off, int64_t n_seq,
                            int64_t min_n, const int64_t* d_contig_off, int32_t* d_n_contigs, int64_t* d_begin,
                            int64_t* d_end, void* stream);
/* baseCount (seqUtils.py:258-265) for n_ranges slices ascii[begin, end): d_acgt[r] = {A, C, G, T+U}, case-insensitive.
 * d_ascii must be 16-byte aligned. */
int dbb_base_count_dev(const uint8_t* d_ascii, const int64_t* d_begin, const int64_t* d_end, int64_t n_ranges,
                       int64_t* d_acgt, void* stream);

/* A count plan = the device-resident work list (chunks of scaffolds, 16 bytes of metadata each)
 * derived from HOST block offsets.  Build once per batch, run many times. */
typedef struct dbb_count_plan dbb_count_plan;
int dbb_count_plan_create(const int64_t* h_blk_off, int64_t n_seq, dbb_count_plan** plan);
int dbb_count_plan_destroy(dbb_count_plan* plan);
/* number of kernels one dbb_tetra_count_dev call launches (for launch accounting) */
int dbb_count_plan_launches(const dbb_count_plan* plan);
/* The kernel classes of stage 1, in 64-base blocks: a scaffold of nb blocks is counted with super-k-mers of
 * S = 1 / 2 / 3 windows when nb <= limits4[0] / [1] / [2], else S = 4, whole if nb <= limits4[3] and otherwise in
 * chunks of limits4[3] blocks (informational: tests straddle these boundaries, results do not depend on them). */
int dbb_count_class_limits(int64_t* limits4);

/* packed planes -> counts uint32[n_seq][136] (bit-exact) and freq float[n_seq][freq_stride]
 * (freq = count / sum(count); all-zero counts give NaN as genomicSignatures.py:147-148 does;
 * freq_stride >= 136, columns 136..freq_stride-1 are written as 0).  Either output may be NULL. */
int dbb_tetra_count_dev(dbb_count_plan* plan, const void* d_codes, const void* d_valid,
                        uint32_t* d_counts, float* d_freq, int64_t freq_stride, void* stream);

/* Whole stage 1 from host memory: ASCII in, counts/freq out (freq is [n_seq][136], dense).
 * Replaces the per-sequence loop of genomicSignatures.py:152-181.
 * The input goes up in batches (DBB_HOST_BATCH_MB; default 32 MB for pinned, 512 MB for pageable input), two in flight.
 * Pinned (or registered) input is copied as ASCII.  Pageable input of 8 MB and more per batch is uploaded by two routes at once: the calling thread copies
 * ASCII from the front of the batch while a few host threads (DBB_HOST_PACK_THREADS; default: the hardware threads /
 * LOCAL_WORLD_SIZE - 1, at most 16; 0 or DBB_HOST_PACK=0 = off) pack scaffolds from its back with
 * dbb_pack_ascii_host's code and send 0.375 bytes per base; the split follows the speeds of the two.  Results do not
 * depend on the route (the planes are bit-identical). */
int dbb_tetra_signatures_host(const uint8_t* ascii, const int64_t* seq_off, int64_t n_seq,
                              uint32_t* counts, float* freq);
/* The same for sequences that are NOT back to back: one host pointer and one length per scaffold (the reference keeps its
 * sequences as separate Python strings, seqUtils.readFasta -> genomicSignatures.py:152-181).  Nothing is concatenated:
 * the packer threads read every scaffold where it lies and only packed planes cross the link. */
int dbb_tetra_signatures_host_ptrs(const uint8_t* const* seqs, const int64_t* seq_len, int64_t n_seq,
                                   uint32_t* counts, float* freq);
/* bytes the last dbb_tetra_signatures_host call on the current device sent over the link as ASCII and as host-packed
 * planes, and the number of packer threads in use (any pointer may be NULL) */
int dbb_host_upload_stats(int64_t* ascii_bytes, int64_t* plane_bytes, int32_t* pack_threads);

/* ---- stage 2: all-pairs tetranucleotide distance, neighbour predicates and fused top-k ------------
 * tdNeighbours.py:41-51, gcNeighbours.py:43-52, coverageNeighbours.py:41-50 with the bounds of
 * distributions.py:75-127 resolved ONCE PER SCAFFOLD on the host (the algebraic restatement in
 * SURVEY.md section 8a): every per-pair bound depends only on per-scaffold scalars.
 * dbb_pairs_dev: sig and every [n] array must be 16-byte aligned (cudaMalloc / torch allocations are);
 * dbb_pairs_host has no alignment requirement.                                                  */
#define DBB_METRIC_L1 0
#define DBB_METRIC_L2 1
typedef struct {
  int64_t n;                   /* scaffolds (columns) */
  const float* sig;            /* [n][DBB_SIG_STRIDE] float32 frequencies, cols 136..139 zero */
  const int32_t* length;       /* [n] */
  /* TD predicate (tdNeighbours.py:43-51): pass iff L1 < bound of the SHORTER scaffold (ties -> i).
   * td_bound[s] = tdDist[nearest(len_s or fixedSeqLen)][tdKey]; NULL disables the predicate. */
  const float* td_bound;       /* [n] or NULL */
  /* GC predicate (gcNeighbours.py:44-52, distributions.py:75-93); gc == NULL disables it. */
  const double* gc;            /* [n] */
  const int32_t* gc_mbin;      /* [n] nearest mean-GC key index of the scaffold (used when it is the core) */
  const int32_t* gc_lbin;      /* [n] nearest length key index (used when it is the unbinned one) */
  const double* gc_lo;         /* [gc_nm][gc_nl] lower bounds at the lower percentile key */
  const double* gc_hi;         /* [gc_nm][gc_nl] */
  int32_t gc_nm, gc_nl;
  /* coverage predicate (coverageNeighbours.py:42-50, distributions.py:106-127); cov == NULL disables */
  const double* cov;           /* [n] */
  const int32_t* cov_lbin;     /* [n] */
  const double* cov_lo;        /* [cov_nl] */
  const double* cov_hi;        /* [cov_nl] */
  int32_t cov_nl;
  /* Distance: DBB_METRIC_L1 = Manhattan, the reference's metric (genomicSignatures.py:183-184, tdNeighbours.py:43).
   * DBB_METRIC_L2 = Euclidean, an addition for k-NN callers (k > 0; td_bound then bounds the Euclidean distance
   * by the same shorter-scaffold rule; GC / coverage predicates and masks are not available with it).
   * Both run on the FP32 CUDA cores; tolerance 2e-6 absolute against FP64 (tests/test_pairs_gpu.py). */
  int32_t metric;
  /* Optional sparse sweep, dbb_pairs_dev only (all NULL = every row visits every column).  The caller may know that
   * whole 128 x 128 tiles of the pair matrix cannot hold a neighbour or a top-k candidate -- e.g. with the scaffolds
   * sorted by a projection p = sum_k w_k sig_k, w_k in [0, 1], L1(a, b) >= 2 |p_a - p_b| because signatures sum
   * to 1 (dbb_b200/device.py::pairs_pruned builds exactly that) -- and passes, for row tile t of [row0, row1)
   * (rows row0 + 128 t ...), the ascending list of column tiles it must visit: tile_list[tile_off[t] .. tile_off[t+1]).
   * item_order[n_row_tiles]: a permutation of the row tiles, most work first (scheduling only).  col_id[n]: the
   * caller's own id of column j, used instead of j for the (distance, id) order of the top-k lists and in knn_idx,
   * so that results do not depend on how the caller sorted the scaffolds.  Masks cannot be requested with tile lists. */
  const int32_t* col_id;
  const int32_t* item_order;
  const int64_t* tile_off;
  const int32_t* tile_list;
  int32_t split_all;           /* > 0: every row tile is swept as this many segments (items), merged afterwards; 0: the
                                  library splits only the row tiles of the last partial wave */
  int32_t n_active_row_tiles;  /* > 0: only the first so many row tiles of item_order are processed (the outputs of
                                  the other rows are left untouched); 0: all */
  int32_t split_tiles;         /* > 0 (with split_all and tile lists): a row tile whose list holds L tiles is cut into
                                  min(split_all, ceil(L / split_tiles)) segments instead of split_all -- long lists are
                                  cut finely (the sweep ends with its longest item), short ones not at all (every
                                  segment refills its own top-k lists); 0: split_all segments for every row tile */
  int32_t reserved0;
} dbb_pairs_input;

typedef struct {
  /* top-k (NEW output, not in the reference): for row i the candidates are j != i that pass the
   * predicates selected by topk_filter, ordered by (distance asc, j asc).  k == 0 disables. */
  int32_t k;                   /* 0..DBB_MAX_K */
  int32_t topk_filter;         /* bit0: TD, bit1: GC, bit2: coverage must pass to be a candidate */
  int32_t* knn_idx;            /* [rows][k] padded with -1, or NULL */
  float* knn_dist;             /* [rows][k] padded with +inf, or NULL */
  /* neighbourhood of the intersection of all ENABLED predicates, j == i included, i.e. what
   * greedy.py:407-417 derives from the three masks */
  int32_t* count;              /* [rows] or NULL */
  int64_t* neighbour_bp;       /* [rows] sum of length[j] over that intersection, or NULL */
  /* bit masks, one bit per column, row stride mask_words uint32 (bit j%32 of word j/32):
   * the dense 0/1 rows the reference stores in .tdNeighbours / .gcNeighbours / .covNeighbours */
  uint32_t* td_mask;           /* [rows][mask_words] or NULL */
  uint32_t* gc_mask;
  uint32_t* cov_mask;
  int64_t mask_words;
} dbb_pairs_output;

/* rows [row0, row1) of the all-pairs problem against all n columns.  All pointers are device
 * pointers.  Enqueues on `stream`. */
int dbb_pairs_dev(const dbb_pairs_input* in, int64_t row0, int64_t row1, const dbb_pairs_output* out,
                  void* stream);
/* bytes of stream-ordered device scratch dbb_pairs_dev allocates internally for (n, rows, k), and the number
 * of kernels it launches (pair kernel + merge of column-split row tiles) -- informational */
int64_t dbb_pairs_scratch_bytes(int64_t n, int64_t rows, int32_t k);
int dbb_pairs_launches(int64_t n, int64_t rows);
/* same with HOST pointers everywhere (sig is dense [n][136] float32 on the host) */
int dbb_pairs_host(const dbb_pairs_input* in_host, int64_t row0, int64_t row1,
                   const dbb_pairs_output* out_host);

/* ---- planner of the exact pruned sweep (the sparse sweep of dbb_pairs_input, built on the device) -------------
 * Sorts the scaffolds by (class, p = <sig, weights>), carries the signature rows and per-scaffold tables into that
 * order, and writes the tile lists dbb_pairs_dev takes: a (row tile, column tile) pair is visited iff
 * 2 * gap(p intervals) - margin <= max(largest td_bound of the row tile, of the column tile)  [no td_bound: <= 0].
 * Exact for any weights in [0, 1] because signatures sum to 1: L1(a, b) >= 2 |p_a - p_b|.  PRECONDITION: every finite
 * row of sig sums to 1 within 0.3 * margin (frequencies do; raw counts do not) -- checked on the device, see info[4].  Lists are ordered
 * nearest tile first.  world > 1: the row tiles are dealt to the ranks by decreasing list length; a rank's lists
 * cover its own row tiles only (mine[t] = 1).  All pointers are device pointers, caller-allocated. */
typedef struct {
  int64_t n;
  const float* sig;            /* [n][DBB_SIG_STRIDE] */
  const float* weights;        /* [DBB_SIG_STRIDE], in [0, 1], padding columns 0 */
  const float* cls;            /* [n] sort class 0, 1, 2, ... as float (e.g. quantile of td_bound), or NULL */
  const int32_t* length; const float* td_bound; const double* gc; const int32_t* gc_mbin; const int32_t* gc_lbin;
  const double* cov; const int32_t* cov_lbin;     /* [n] each; NULL = absent */
  int64_t row0, row1;          /* rows of the SORTED order the sweep will cover */
  int32_t rank, world;         /* world <= 1: single GPU; else row0 = 0, row1 = n */
  float margin;                /* slack: covers |sum(sig) - 1| <= 0.3 * margin of every row and the FP32 rounding of p; e.g. 1e-5 */
} dbb_prune_input;

typedef struct {
  float* sig_sorted;           /* [n][DBB_SIG_STRIDE] */
  int32_t* length; float* td_bound; double* gc; int32_t* gc_mbin; int32_t* gc_lbin; double* cov; int32_t* cov_lbin;
  int32_t* col_id;             /* [n] original index of sorted position j */
  float* p_sorted;             /* [n] */
  int32_t* item_order;         /* [row tiles] by decreasing list length */
  int64_t* tile_off;           /* [row tiles + 1] */
  int32_t* tile_list;          /* capacity: row tiles of this rank x column tiles */
  int32_t* tile_count;         /* [row tiles] */
  uint8_t* mine;               /* [row tiles] */
  int64_t* info;               /* [8]: [0..1] tiles listed, row tiles with a non-empty list -- first plan; [2..3] second plan;
                                  [4] rows whose signature does not sum to 1 within 0.3 * margin (finite rows only): the
                                  bound does not hold for them, the caller must refuse the plan's results if it is non-zero */
  void* scratch; int64_t scratch_bytes;   /* dbb_prune_scratch_bytes; must be left alone between the plan calls */
} dbb_prune_buffers;

int64_t dbb_prune_scratch_bytes(int64_t n, int64_t rows, int32_t world);
int dbb_prune_plan_dev(const dbb_prune_input* in, const dbb_prune_buffers* buf, void* stream);
/* Unfiltered top-k: after the first sweep, the tiles not yet visited whose lower bound does not exceed the largest
 * k-th best distance (d_kth[row * kth_stride], row relative to row0) of the row tile.  Overwrites item_order /
 * tile_off / tile_list / tile_count (the affected row tiles sort first; info[3] = how many).  Same `in` as the plan. */
int dbb_prune_second_dev(const dbb_prune_input* in, const dbb_prune_buffers* buf, const float* d_kth, int64_t kth_stride,
                         void* stream);
/* Merges the second sweep's top-k lists into the first sweep's for the rows of the affected row tiles, in the
 * (distance, id) order of dbb_pairs_output ([rows][k] each, ids as written by dbb_pairs_dev, -1 padding). */
int dbb_prune_merge_dev(const dbb_prune_buffers* buf, int64_t rows, int32_t k, int32_t* d_idx1, float* d_dist1,
                        const int32_t* d_idx2, const float* d_dist2, void* stream);

/* ---- k nearest neighbours + neighbourhood sizes of all rows in ONE call: the default stage 2 -----------------
 * (tdNeighbours.py:41-51 with gcNeighbours.py:43-52 / coverageNeighbours.py:41-50 as dbb_pairs_input describes them;
 * the top-k is the new output of dbb_pairs_output.)  With prune != 0 the library plans and runs the exact
 * tile-skipping sweep itself -- class quantiles of td_bound, dbb_prune_plan_dev, sparse dbb_pairs_dev, for a top-k that
 * is not TD-filtered the second plan + sweep + list merge -- and returns the rows in the caller's order; results are
 * identical to the full sweep (prune == 0), ties included.  prune needs td_bound, the Manhattan metric and signature
 * rows that sum to 1: rows that do not are counted on the device and the call fails with DBB_ERR_INVALID (prune == 1)
 * or falls back to the full sweep on the GPU (prune == -1, "auto").
 * Multi-GPU (world > 1): this rank takes its share of the 128-row tiles of the sorted order, dealt by decreasing work
 * (prune) or the equal row block [n*rank/world, n*(rank+1)/world) (full sweep); the rows come back in ascending
 * sorted position with their original ids in row_ids.  The call synchronises `stream` once (list sizes). */
typedef struct {
  int32_t k;                   /* 1..DBB_MAX_K */
  int32_t topk_filter;         /* as dbb_pairs_output.topk_filter */
  int32_t prune;               /* 1: tile-skipping sweep, 0: full sweep, -1: auto */
  int32_t n_classes;           /* sort classes of td_bound for the pruned sweep; <= 0: default (8) */
  float margin;                /* slack of the projection bound; <= 0: default (1e-5) */
  int32_t rank, world;         /* world <= 1: single GPU */
} dbb_knn_params;

typedef struct {
  int64_t cap_rows;            /* in: rows the arrays below can hold (n for world <= 1; >= 128 * ceil(ceil(n/128)/world) else) */
  int64_t* row_ids;            /* [cap_rows] original id of output row r; may be NULL when world <= 1 (row r = scaffold r) */
  int32_t* knn_idx;            /* [cap_rows][k] */
  float* knn_dist;             /* [cap_rows][k] */
  int32_t* count;              /* [cap_rows] or NULL */
  int64_t* neighbour_bp;       /* [cap_rows] or NULL */
  int64_t n_rows;              /* out: rows written */
  int64_t tiles_visited, tiles_total;   /* out: 128 x 128 tiles executed / in this rank's share of the pair matrix */
  int64_t tiles_second, row_tiles_second;   /* out: of those, tiles / row tiles of the second sweep */
  int32_t pruned;              /* out: 1 if the tile-skipping sweep ran */
  int32_t launches;            /* out: kernels launched by the call */
  float ms_plan, ms_sweep, ms_second;   /* out: device time of planning, first sweep, second plan + sweep + merge */
} dbb_knn_output;

/* all pointers of `in` and `out` are DEVICE pointers (in->sig is [n][DBB_SIG_STRIDE], 16-byte aligned) */
int dbb_knn_dev(dbb_ctx* ctx, const dbb_pairs_input* in, const dbb_knn_params* par, dbb_knn_output* out, void* stream);
/* all pointers are HOST pointers (in->sig is dense [n][136] float32); runs on the context's own stream */
int dbb_knn_host(dbb_ctx* ctx, const dbb_pairs_input* in_host, const dbb_knn_params* par, dbb_knn_output* out_host);

/* ---- "next" row N1: rectangular filtered nearest-core search -- refineBins.py:42-96 (__workerThread), the same
 * ---- loop in unbinned.py:156-176.  For every unbinned contig u over all cores c (in array order):
 * valid(u,c) = withinDistGC(u,c) && withinDistCov(u,c) && L1(u,c) < boundDistTD(u) (distributions.py:75-127 with
 * fixedSeqLen = None); closest valid core in GC / TD / coverage space with strict '<' (first core wins ties);
 * assigned iff the three agree.  All FP64, as in the reference.  Bounds are resolved per contig / per core on the
 * host: u_td_bound[u] = tdDist[nearest(len_u)][tdKey], u_gc_lbin / u_cov_lbin = nearest length key of u,
 * c_gc_mbin = nearest mean-GC key of the core. */
typedef struct {
  int64_t n_unbinned, n_cores;
  const double* u_sig;         /* [U][136] */
  const double* c_sig;         /* [B][136] */
  const double* u_gc;  const double* u_cov;   /* [U] */
  const double* c_gc;  const double* c_cov;   /* [B] */
  const double* u_td_bound;    /* [U] */
  const int32_t* u_gc_lbin; const int32_t* u_cov_lbin;   /* [U] */
  const int32_t* c_gc_mbin;    /* [B] */
  const double* gc_lo; const double* gc_hi; int32_t gc_nm, gc_nl;
  const double* cov_lo; const double* cov_hi; int32_t cov_nl;
} dbb_refine_input;

typedef struct {
  int32_t* closest;            /* [U][3] closest valid core in GC, TD, coverage space; -1 if no core is valid; or NULL */
  double* min_dist;            /* [U][3] the three minimum distances (1e9 if none); or NULL */
  int32_t* assigned;           /* [U] the core if the three agree, else -1; or NULL */
  uint8_t* within;             /* [U][B] bit0 GC, bit1 coverage, bit2 TD predicate (unbinned.py:137-152); or NULL */
  double* dist;                /* [U][B] the L1 distance of every (contig, core) pair (compare.py:108,117); or NULL */
} dbb_refine_output;

int dbb_refine_dev(const dbb_refine_input* in, const dbb_refine_output* out, void* stream);
int dbb_refine_host(const dbb_refine_input* in_host, const dbb_refine_output* out_host);

/* ---- "next" row N5: bin-to-bin merge table -- Merge.run, merge.py:96-156 ------------------------------------
 * The same filtered search with the BINNED contigs as queries: `search.n_unbinned` contigs against the
 * `search.n_cores` bins (array order = the order merge.py iterates `bins`), own[u] = position of the contig's own
 * bin.  For every ordered pair of different bins (b1, b2), over the contigs u of b1 (merge.py:101-151):
 *   table[b1][b2][0..2] = how many are withinDistGC / withinDistCov / witinDistTD of b2 (:112-119),
 *   table[b1][b2][3..5] = for how many b2 is the closest valid bin (valid = all three predicates, the contig's own
 *                         bin skipped, :124-145) in GC / TD / coverage space -- the order of merge.py's
 *                         `distances = [gcDistance, tdDistance, covDistance]` (:140); strict '<', first bin wins ties.
 * table is int32 [B][B][6], zeroed by the call; the diagonal stays zero.  closest: [U][3] per contig, or NULL. */
typedef struct {
  dbb_refine_input search;
  const int32_t* own;          /* [U], each in [0, B) */
} dbb_merge_input;
int dbb_merge_table_dev(const dbb_merge_input* in, int32_t* d_table, int32_t* d_closest, void* stream);
int dbb_merge_table_host(const dbb_merge_input* in_host, int32_t* table, int32_t* closest);

/* ---- N6: statistics of unbinned scaffolds -- Unbinned.run, unbinned.py:111-186 ----------------------------------
 * The same search once more: queries = the contigs of the scaffolds under study, group[u] = index of the contig's
 * scaffold, cores = the bins (array order = the order unbinned.py iterates `bins`), nothing excluded.  Per
 * (scaffold s, bin b), over the contigs of s:
 *   table[s][b][0..2] = within the GC / coverage / TD distribution of b (:137-152),
 *   table[s][b][3..5] = b is the closest valid bin in GC / TD / coverage space (:156-178; the search only runs for
 *                       contigs within all three of b, which "b is the closest VALID bin" implies),
 *   table[s][b][6]    = b is the closest in TD and in coverage space (the binning score's numerator, :181-182).
 * table is int32 [n_groups][B][7], zeroed by the call. */
typedef struct {
  dbb_refine_input search;
  const int32_t* group;        /* [U], each in [0, n_groups) */
  int64_t n_groups;
} dbb_unbinned_input;
int dbb_unbinned_table_dev(const dbb_unbinned_input* in, int32_t* d_table, void* stream);
int dbb_unbinned_table_host(const dbb_unbinned_input* in_host, int32_t* table);

/* dense [n][136] -> padded [n][DBB_SIG_STRIDE] on device (used after the all-gather) */
int dbb_pad_signatures_dev(const float* d_dense, int64_t n, float* d_padded, void* stream);

/* ---- synthetic input (benchmark/test support; dbb_b200/synthetic.py is the host twin) ------------ */
int dbb_synth_ascii_dev(uint64_t seed, int64_t n_seq, const int64_t* d_seq_off, const int64_t* d_blk_off,
                        int64_t total_blocks, const int32_t* d_genome, const uint32_t* d_cum /* [G][64][3] */,
                        const int64_t* d_n_start, const int64_t* d_n_len, const int64_t* d_lc_start,
                        const int64_t* d_lc_len, uint8_t* d_ascii, void* stream);

/* ---- "next" row N3: greedy core growth -- Greedy.__greedy, greedy.py:167-251 ----------------------------------
 * Contigs in the caller's order (greedy.py walks `sortedContigs`).  Per contig (all resolved on the host as in the
 * refine row): td_bound[i] = tdDist[nearest(len_i)][tdKey] (witinDistTD(d, contig)), gc_lbin / cov_lbin = nearest
 * length key of the contig; gc_mean_keys[gc_nm] are the mean-GC keys of the GC table in ascending order (the core's
 * key is re-resolved at every step because its GC moves).  Outputs: bin_id[n] (-1 = unbinned, Greedy.UNBINNED),
 * the number of cores kept, and per kept core its length, GC, coverage and signature exactly as the running
 * weighted means of greedy.py:229-234 leave them ([n] / [n][136] capacity; any may be NULL).  The loop over growth
 * steps runs on the host, the scan over the candidates of each step on the GPU (one launch per step). */
typedef struct {
  int64_t n;
  const double* sig;           /* [n][136] float64 signatures */
  const int64_t* length;       /* [n] */
  const double* gc; const double* cov;               /* [n] */
  const double* td_bound;      /* [n] */
  const int32_t* gc_lbin; const int32_t* cov_lbin;   /* [n] */
  const double* gc_mean_keys;  /* [gc_nm] ascending */
  const double* gc_lo; const double* gc_hi; int32_t gc_nm, gc_nl;
  const double* cov_lo; const double* cov_hi; int32_t cov_nl;
} dbb_greedy_input;
int dbb_greedy_host(const dbb_greedy_input* in_host, int64_t min_bin_size, int32_t* bin_id, int64_t* n_cores,
                    int64_t* core_len, double* core_gc, double* core_cov, double* core_sig, int64_t* n_steps);

/* ---- "next" row N4: PCA of the signature matrix -- PCA.pcaMatrix, pca.py:56-75, with Center, pca.py:113-126 ----
 * dbb_pca_gram_host: mean[w] = A.mean(0); std[w] = population std of the centred columns if scale != 0 (0 -> 1,
 * pca.py:121-122) else ones; gram[w][w] = X^T X of X = (A - mean) / std, all FP64.  The caller solves the w x w
 * symmetric eigenproblem (singular values d = sqrt(eigenvalues), rows of Vt = eigenvectors) and gets the scores
 * pc() = X . V[:, :npc] from dbb_pca_project_host (V column-major-by-component: V[c][k], [w][npc]). */
int dbb_pca_gram_host(const double* A, int64_t n, int32_t w, int32_t scale, double* mean, double* std_out, double* gram);
int dbb_pca_project_host(const double* A, int64_t n, int32_t w, const double* mean, const double* std_in, const double* V,
                         int32_t npc, double* scores);

#ifdef __cplusplus
}
#endif
#endif /* DBB_CUDA_H */
