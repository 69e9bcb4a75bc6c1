/* C restatement of the reference's hot-path arithmetic -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Built by oracle/Makefile into oracle/_build/libdbb_oracle.so and loaded by tests/ and bench.py only (ctypes).
 * It exists so that full-size inputs (1 Gbp, 50 000 x 50 000 pairs) can be checked exactly in seconds; it is validated
 * against the Python literal oracle (oracle/dbb_oracle.py), which is pinned to the reference's own output.
 * Every function cites the /root/reference/dbb lines it follows. */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* genomicSignatures.py:74-84 : reverse complement / lexicographically lowest, on 4-letter strings */
static void rev_comp4(const char* s, char* out) {
  for (int i = 0; i < 4; ++i) {
    char c = s[3 - i];
    out[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
  }
  out[4] = 0;
}

/* genomicSignatures.py:44-72 : enumerate k-mers lexicographically, canonical = min(kmer, revcomp) by string compare,
 * column = order of first appearance; map every k-mer and its reverse complement to the column.
 * table[256] is indexed by the 2-bit code (A0 C1 G2 T3, first base most significant). */
int oracle_kmer_table(uint8_t* table, char* names680) {
  static const char B[4] = {'A', 'C', 'G', 'T'};
  char cols[136][5];
  int ncol = 0;
  for (int code = 0; code < 256; ++code) {
    char mer[5], rc[5];
    for (int p = 0; p < 4; ++p) mer[p] = B[(code >> (2 * (3 - p))) & 3];
    mer[4] = 0;
    rev_comp4(mer, rc);
    const char* canon = strcmp(mer, rc) < 0 ? mer : rc;           /* :74-79 */
    int found = -1;
    for (int c = 0; c < ncol; ++c) if (strcmp(cols[c], canon) == 0) { found = c; break; }   /* :60-62 "not in retList" */
    if (found < 0) { strcpy(cols[ncol], canon); found = ncol++; }
    table[code] = (uint8_t)found;
  }
  if (names680) for (int c = 0; c < ncol; ++c) memcpy(names680 + 5 * c, cols[c], 5);
  return ncol;
}

static inline int base_code(unsigned char ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1; }

/* genomicSignatures.py:134-144 : sliding window, stride 1; a window holding any byte outside uppercase ACGT is a
 * KeyError in the reference and is skipped.  counts[136] int64. */
void oracle_seq_counts(const uint8_t* seq, int64_t len, const uint8_t* table, int64_t* counts) {
  memset(counts, 0, 136 * sizeof(int64_t));
  for (int64_t i = 0; i + 4 <= len; ++i) {
    int c0 = base_code(seq[i]), c1 = base_code(seq[i + 1]), c2 = base_code(seq[i + 2]), c3 = base_code(seq[i + 3]);
    if ((c0 | c1 | c2 | c3) < 0) continue;
    counts[table[(c0 << 6) | (c1 << 4) | (c2 << 2) | c3]] += 1;
  }
}

/* batch form over a concatenated buffer (OpenMP over sequences) */
void oracle_counts_batch(const uint8_t* flat, const int64_t* off, int64_t n, int64_t* counts /* [n][136] */) {
  uint8_t table[256];
  oracle_kmer_table(table, 0);
  #pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < n; ++s) oracle_seq_counts(flat + off[s], off[s + 1] - off[s], table, counts + s * 136);
}

/* genomicSignatures.py:146-150 : float64 normalisation (0/0 -> NaN) */
void oracle_signature(const int64_t* counts, double* sig) {
  double tot = 0;
  for (int c = 0; c < 136; ++c) tot += (double)counts[c];
  for (int c = 0; c < 136; ++c) sig[c] = (double)counts[c] / tot;
}

/* tdNeighbours.py:41-51 with distributions.py:95-104 resolved per scaffold: for the given rows, the float64 L1
 * distance to every column and the 0/1 neighbour flag (bound of the SHORTER scaffold, ties -> i; strict '<').
 * sig [n][136] float64; td_bound[n]; out_dist / out_mask [nrows][n] (either may be NULL). */
void oracle_td_rows(const double* sig, const int64_t* length, const double* td_bound, int64_t n, const int64_t* rows,
                    int64_t nrows, double* out_dist, uint8_t* out_mask) {
  #pragma omp parallel for schedule(dynamic, 1)
  for (int64_t r = 0; r < nrows; ++r) {
    const int64_t i = rows[r];
    const double* a = sig + i * 136;
    for (int64_t j = 0; j < n; ++j) {
      const double* b = sig + j * 136;
      double d = 0;
      for (int c = 0; c < 136; ++c) d += fabs(a[c] - b[c]);
      if (out_dist) out_dist[r * n + j] = d;
      if (out_mask) {
        const double bound = length[i] > length[j] ? td_bound[j] : td_bound[i];
        out_mask[r * n + j] = d < bound ? 1 : 0;
      }
    }
  }
}
