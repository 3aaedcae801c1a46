/*
 * bz_oracle.c -- CPU restatement of lbzip2's per-block compressor.
 *
 * TEST INFRASTRUCTURE ONLY (see bz_oracle.h).  Plain C, written from the
 * behavioural description of the reference and
 * pinned against the compiled reference stage by stage.  The structure is
 * deliberately different from the reference's (piece-wise RLE1 instead of a
 * goto state machine, prefix-doubling BWT instead of divsufsort, explicit
 * package-merge lists instead of the lazy boundary variant); only the
 * results are the same.  Each function cites the reference lines it restates.
 */
#include "bz_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ CRC -- */

static uint32_t crc_tab[256];
static int crc_tab_ready;

/* Table as produced by build-aux/make-crctab.pl:29-33 (== crctab.c:26). */
static void
crc_tab_init(void)
{
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t r = i << 24;
    for (int k = 0; k < 8; k++)
      r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0);
    crc_tab[i] = r;
  }
  crc_tab_ready = 1;
}

uint32_t
orc_crc_update(uint32_t crc, const uint8_t *p, size_t n)
{
  if (!crc_tab_ready) crc_tab_init();
  for (size_t i = 0; i < n; i++)      /* encode.c:103 */
    crc = (crc << 8) ^ crc_tab[(crc >> 24) ^ p[i]];
  return crc;
}

/* ----------------------------------------------------------------- RLE1 -- */

/* Piece-wise view of collect() (encode.c:135-336): the input is cut into
   "pieces" of 1..259 equal bytes (a longer run simply starts a new piece,
   encode.c:266-272); a piece of length r emits min(r,4) literals and, when
   r >= 4, the count byte r-4.  The block closes as soon as it holds `cap`
   bytes after any write (encode.c:162,176,202,218,256-264), or when, right
   after the third literal, exactly one slot is left and the run goes on
   (encode.c:218) -- the 4th literal is only written with room for its count
   byte.  A count byte that is still pending at end of input is appended by
   encode() (encode.c:443-447) and does not make collect() report "full".  */
int
orc_rle1(const uint8_t *in, size_t n, uint32_t cap, uint8_t *block,
         uint32_t *nblock, size_t *consumed, uint8_t used[256], uint32_t *crc)
{
  size_t i = 0;
  uint32_t m = 0;
  int full = 0;

  memset(used, 0, 256);
  while (i < n && !full) {
    uint8_t c = in[i];
    size_t r = 1;
    while (r < 259 && i + r < n && in[i + r] == c) r++;

    used[c] = 1;
    block[m++] = c;                                   /* 1st copy */
    if (m >= cap) { i += 1; full = 1; break; }
    if (r == 1) { i += 1; continue; }
    block[m++] = c;                                   /* 2nd copy */
    if (m >= cap) { i += 2; full = 1; break; }
    if (r == 2) { i += 2; continue; }
    block[m++] = c;                                   /* 3rd copy */
    if (m >= cap || (m == cap - 1 && r >= 4)) { i += 3; full = 1; break; }
    if (r == 3) { i += 3; continue; }
    block[m++] = c;                                   /* 4th copy */
    block[m++] = (uint8_t)(r - 4);                    /* count byte */
    used[r - 4] = 1;
    i += r;
    /* The count byte is written by collect() itself unless the piece is cut
       short by the end of the input (then encode() appends it).  */
    if (m >= cap && (r == 259 || i < n)) full = 1;
  }
  *nblock = m;
  *consumed = i;
  *crc = orc_crc_update(0xFFFFFFFFu, in, i);
  return full;
}

/* ------------------------------------------------------------------ BWT -- */

/* Cyclic-rotation sort by prefix doubling with two counting-sort passes per
   round (contract: SURVEY.md A.2; reference divbwt.c:1707-1726 obtains the
   same last column through divsufsort).  O(n log n) time, 5 int arrays.  */
uint32_t
orc_bwt(const uint8_t *t, uint32_t n, uint8_t *bwt, uint32_t *tie_count)
{
  if (tie_count) *tie_count = 1;
  if (n == 0) return 0;
  if (n == 1) { bwt[0] = t[0]; return 0; }

  uint32_t nk = n > 65536 ? n : 65536;
  uint32_t *sa = malloc(sizeof(uint32_t) * n);
  uint32_t *tmp = malloc(sizeof(uint32_t) * n);
  uint32_t *rk = malloc(sizeof(uint32_t) * n);
  uint32_t *nr = malloc(sizeof(uint32_t) * n);
  uint32_t *cnt = malloc(sizeof(uint32_t) * ((size_t)nk + 1));
  uint32_t h, distinct = 0, i;

  /* Round 0: rank = first two bytes (cyclic). */
  for (i = 0; i < n; i++)
    rk[i] = ((uint32_t)t[i] << 8) | t[i + 1 < n ? i + 1 : 0];
  memset(cnt, 0, sizeof(uint32_t) * 65537);
  for (i = 0; i < n; i++) cnt[rk[i] + 1]++;
  for (i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
  for (i = 0; i < n; i++) sa[cnt[rk[i]]++] = i;
  /* densify: rank = start position of the group */
  {
    uint32_t start = 0;
    distinct = 0;
    for (i = 0; i < n; i++) {
      if (i == 0 || rk[sa[i]] != rk[sa[i - 1]]) { start = i; distinct++; }
      nr[sa[i]] = start;
    }
    memcpy(rk, nr, sizeof(uint32_t) * n);
  }

  for (h = 2; distinct < n && h < n; h *= 2) {
    /* sort by (rk[i], rk[i+h]): LSD -- second key first */
    memset(cnt, 0, sizeof(uint32_t) * ((size_t)n + 1));
    for (i = 0; i < n; i++) cnt[rk[i] + 1]++;
    for (i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    uint32_t *c2 = nr;                     /* reuse as second counter set */
    memcpy(c2, cnt, sizeof(uint32_t) * n);
    for (i = 0; i < n; i++) {
      uint32_t j = i + h; if (j >= n) j -= n;
      tmp[c2[rk[j]]++] = i;                /* ordered by rk[i+h], stable by i */
    }
    for (i = 0; i < n; i++) sa[cnt[rk[tmp[i]]]++] = tmp[i];
    /* new ranks */
    uint32_t start = 0;
    distinct = 0;
    for (i = 0; i < n; i++) {
      uint32_t a = sa[i];
      if (i == 0) { start = 0; distinct = 1; }
      else {
        uint32_t b = sa[i - 1];
        uint32_t a2 = a + h; if (a2 >= n) a2 -= n;
        uint32_t b2 = b + h; if (b2 >= n) b2 -= n;
        if (rk[a] != rk[b] || rk[a2] != rk[b2]) { start = i; distinct++; }
      }
      tmp[a] = start;
    }
    memcpy(rk, tmp, sizeof(uint32_t) * n);
  }

  uint32_t pidx = rk[0];     /* first position of rotation 0's group */
  if (tie_count) {
    uint32_t k = 0;
    for (i = 0; i < n; i++) k += (rk[i] == pidx);
    *tie_count = k;
  }
  for (i = 0; i < n; i++) bwt[i] = t[sa[i] ? sa[i] - 1 : n - 1];

  free(sa); free(tmp); free(rk); free(nr); free(cnt);
  return pidx;
}

/* ------------------------------------------------------------------ MTF -- */

uint32_t
orc_mtf(const uint8_t *bwt, uint32_t n, const uint8_t used[256],
        uint16_t *mtfv, uint32_t *freq, uint32_t *alpha_size)
{
  uint8_t dense[256], list[256];
  uint32_t ninuse = 0, nm = 0, run = 0, i;

  for (i = 0; i < 256; i++) { dense[i] = (uint8_t)ninuse; ninuse += used[i] != 0; }
  uint32_t eob = ninuse + 1;               /* encode.c:450 */
  memset(freq, 0, sizeof(uint32_t) * (ORC_MAX_ALPHA + 1));
  for (i = 0; i < 256; i++) list[i] = (uint8_t)i;

#define FLUSH_RUN()                                                    \
  do {                                                                 \
    /* bijective base-2, LSB first: bits of run+1 below its top bit    \
       (encode.c:381-386) */                                           \
    for (uint32_t v = run + 1; v > 1; v >>= 1) {                       \
      mtfv[nm] = (uint16_t)(v & 1); freq[v & 1]++; nm++;               \
    }                                                                  \
    run = 0;                                                           \
  } while (0)

  for (i = 0; i < n; i++) {
    uint8_t c = dense[bwt[i]];
    if (list[0] == c) { run++; continue; }
    FLUSH_RUN();
    uint32_t p = 1;
    while (list[p] != c) p++;
    memmove(list + 1, list, p);
    list[0] = c;
    mtfv[nm++] = (uint16_t)(p + 1);
    freq[p + 1]++;
  }
  FLUSH_RUN();
#undef FLUSH_RUN
  mtfv[nm++] = (uint16_t)eob;
  freq[eob]++;
  *alpha_size = eob + 1;
  return nm;
}

/* ------------------------------------------------------- prefix codes -- */

/* Sort symbol indices so that the "heaviest" comes first: by frequency
   descending, ties by symbol ascending.  This is the order sort_alphabet()
   (encode.c:553-567) produces for keys freq<<32 | 1<<16 | (258-sym).  */
static void
order_symbols(const uint32_t *f, uint32_t as, uint16_t *ord)
{
  for (uint32_t i = 0; i < as; i++) {
    uint32_t j = i;
    while (j > 0 && f[ord[j - 1]] < f[i]) { ord[j] = ord[j - 1]; j--; }
    ord[j] = (uint16_t)i;
  }
}

/* Unlimited-length Huffman code lengths used inside the EM loop
   (make_code_lengths() encode.c:713-766 = sort_alphabet + build_tree :574 +
   compute_depths :619).  Greedy two-queue Huffman; on equal weight a leaf is
   taken before an internal node (the reference encodes this in bits 24..31 of
   its 64-bit keys, encode.c:609-610).  Only the NUMBER of leaves per depth
   matters: depths are then dealt out to the symbols in sorted order, the
   heaviest symbol receiving the shortest code (encode.c:750-763).  */
static void
huffman_lengths(uint8_t *length, const uint32_t *frequency, uint32_t as)
{
  uint32_t f[ORC_MAX_ALPHA];
  uint16_t ord[ORC_MAX_ALPHA];
  uint64_t w[2 * ORC_MAX_ALPHA];          /* node weights: leaves then internals */
  int parent[2 * ORC_MAX_ALPHA];
  uint32_t depth_count[64];
  uint32_t i;

  for (i = 0; i < as; i++) f[i] = frequency[i] ? frequency[i] : 1;   /* :739 */
  order_symbols(f, as, ord);

  /* leaves in ascending weight order: leaf k = ord[as-1-k] */
  for (i = 0; i < as; i++) w[i] = f[ord[as - 1 - i]];
  uint32_t li = 0, ii = as, ni = as;      /* next leaf, next internal, next free */
  while (ni < 2 * as - 1) {
    uint32_t pick[2];
    for (int k = 0; k < 2; k++) {
      int take_leaf;
      if (li >= as) take_leaf = 0;
      else if (ii >= ni) take_leaf = 1;
      else take_leaf = w[li] <= w[ii];    /* tie -> leaf first */
      pick[k] = take_leaf ? li++ : ii++;
    }
    w[ni] = w[pick[0]] + w[pick[1]];
    parent[pick[0]] = (int)ni;
    parent[pick[1]] = (int)ni;
    ni++;
  }
  parent[ni - 1] = -1;
  memset(depth_count, 0, sizeof(depth_count));
  for (i = 0; i < as; i++) {
    uint32_t d = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) d++;
    depth_count[d]++;
  }
  uint32_t k = 0;
  for (uint32_t d = 0; d < 64; d++)
    for (uint32_t c = depth_count[d]; c > 0; c--)
      length[ord[k++]] = (uint8_t)d;
}

/* Length-limited code by package-merge (assign_codes() encode.c:882-987 with
   package_merge() :660-710).  Textbook formulation: list N_1 = leaves in
   ascending weight order, N_{k+1} = merge(leaves, pair-wise packages of N_k);
   for a height limit h take the first 2*as-2 items of N_h and follow the
   packages down.  On equal weight a leaf sorts before a package
   (weight_add() encode.c:652-654 puts a non-zero depth in bits 24..31).
   The height is chosen to minimise payload + tree transmission cost, lowest
   height on ties (encode.c:913-945).  Returns that cost.  */
static uint32_t
limited_lengths_and_codes(uint32_t *code, uint8_t *length,
                          const uint32_t *frequency, uint32_t as)
{
  enum { L = 20 };
  uint16_t ord[ORC_MAX_ALPHA];
  uint64_t leafw[ORC_MAX_ALPHA];
  /* per level: item weights and whether an item is a package */
  static __thread uint64_t lw[L + 1][2 * ORC_MAX_ALPHA];
  static __thread uint8_t ispkg[L + 1][2 * ORC_MAX_ALPHA];
  uint32_t cnt[L + 1];
  uint32_t i, k;

  order_symbols(frequency, as, ord);
  for (i = 0; i < as; i++) leafw[i] = frequency[ord[as - 1 - i]];  /* ascending */

  cnt[1] = as;
  for (i = 0; i < as; i++) { lw[1][i] = leafw[i]; ispkg[1][i] = 0; }
  for (k = 2; k <= L; k++) {
    uint32_t np = cnt[k - 1] / 2, a = 0, b = 0, m = 0;
    while (a < as || b < np) {
      uint64_t pw = b < np ? lw[k - 1][2 * b] + lw[k - 1][2 * b + 1] : 0;
      int take_pkg = (a >= as) || (b < np && pw < leafw[a]);   /* strict: tie -> leaf */
      if (take_pkg) { lw[k][m] = pw; ispkg[k][m] = 1; b++; }
      else { lw[k][m] = leafw[a]; ispkg[k][m] = 0; a++; }
      m++;
    }
    cnt[k] = m;
  }

  uint32_t best_cost = 0xFFFFFFFFu, best_h = L;
  uint8_t best_len[ORC_MAX_ALPHA];
  uint8_t len[ORC_MAX_ALPHA];
  memset(best_len, 0, sizeof best_len);
  for (uint32_t h = 2; h <= L; h++) {
    if ((1ul << h) < as) continue;                              /* :914 */
    /* ge[j] = number of leaves with code length >= j, j = 1..h */
    uint32_t ge[L + 2];
    uint32_t take = 2 * as - 2;
    for (k = h; k >= 1; k--) {
      uint32_t leaves = 0;
      for (i = 0; i < take; i++) leaves += !ispkg[k][i];
      ge[h - k + 1] = leaves;
      take = 2 * (take - leaves);
    }
    ge[h + 1] = 0;
    if (ge[h] == 0) break;              /* no code uses the full height: :916-920 */

    /* heaviest symbols get the shortest codes (encode.c:924-933) */
    uint32_t cost = 0, s = 0;
    for (uint32_t d = 1; d <= h; d++)
      for (uint32_t c = ge[d] - ge[d + 1]; c > 0; c--) {
        len[ord[s]] = (uint8_t)d;
        cost += frequency[ord[s]] * d;
        s++;
      }
    for (i = 1; i < as; i++)
      cost += 2 * (uint32_t)abs((int)len[i] - (int)len[i - 1]);  /* :935-937 */
    cost += 5 + as;
    if (cost < best_cost) {
      best_cost = cost; best_h = h;
      memcpy(best_len, len, as);
    }
  }
  (void)best_h;
  memcpy(length, best_len, as);

  /* canonical codes: ascending within each length in symbol order (:948-969) */
  uint32_t next[L + 2], nlen[L + 2];
  memset(nlen, 0, sizeof nlen);
  for (i = 0; i < as; i++) nlen[length[i]]++;
  uint32_t c = 0;
  for (uint32_t d = 1; d <= L; d++) { next[d] = c; c = (c + nlen[d]) << 1; }
  for (i = 0; i < as; i++) code[i] = next[length[i]]++;
  return best_cost;
}

/* Initial partition of the alphabet into frequency-balanced classes
   (generate_initial_trees() encode.c:779-841).  */
static void
initial_classes(uint8_t length[ORC_MAX_TREES][ORC_MAX_ALPHA + 1],
                const uint32_t *freq, uint32_t nm, uint32_t nt)
{
  uint32_t live = 0, cum, a, t;

  memset(length, 1, ORC_MAX_TREES * (ORC_MAX_ALPHA + 1));
  for (a = 0, cum = 0; cum < nm; a++) { cum += freq[a]; live += freq[a] != 0; }
  if (nt > live) nt = live;

  a = 0;
  for (t = 0; nt > 0; t++, nt--) {
    uint32_t f = freq[a], b = a + 1;
    cum = f;
    live -= f != 0;
    while (live > nt - 1 && cum * nt < nm) {
      f = freq[b]; cum += f; live -= f != 0; b++;
    }
    if (cum > f && (2 * cum - f) * nt > 2 * nm) {
      cum -= f; live += f != 0; b--;
    }
    memset(&length[t][a], 0, b - a);
    a = b;
    nm -= cum;
  }
}

void
orc_prefix_code(uint16_t *mtfv, uint32_t nm, uint32_t as,
                const uint32_t *mtffreq, const uint8_t used[256],
                unsigned cluster_factor, struct orc_coding *out)
{
  /* working state indexed by OLD tree number */
  static __thread uint8_t length[ORC_MAX_TREES][ORC_MAX_ALPHA + 1];
  static __thread uint32_t code[ORC_MAX_TREES][ORC_MAX_ALPHA + 1];
  static __thread uint32_t frequency[ORC_MAX_TREES][ORC_MAX_ALPHA + 1];
  static __thread uint8_t sel[ORC_MAX_SELECTORS];
  uint32_t ng = (nm + ORC_GROUP - 1) / ORC_GROUP;
  uint32_t nt = nm > 2400 ? 6 : nm > 1200 ? 5 : nm > 600 ? 4 :
                nm > 300 ? 3 : nm > 150 ? 2 : 1;            /* :1027-1031 */
  uint32_t i, g, t, v;

  for (i = nm; i < ng * ORC_GROUP; i++) mtfv[i] = (uint16_t)as;  /* :1034 */
  initial_classes(length, mtffreq, nm, nt);

  for (unsigned iter = 0; iter < cluster_factor; iter++) {   /* :1043-1084 */
    uint64_t pack[ORC_MAX_ALPHA + 1];
    for (v = 0; v < as; v++) {
      pack[v] = 0;
      for (t = 0; t < ORC_MAX_TREES; t++)
        pack[v] += (uint64_t)length[t][v] << (10 * t);
    }
    pack[as] = 0;
    memset(frequency, 0, nt * sizeof(frequency[0]));
    for (g = 0; g < ng; g++) {
      const uint16_t *gs = mtfv + g * ORC_GROUP;
      uint64_t cp = 0;
      for (i = 0; i < ORC_GROUP; i++) cp += pack[gs[i]];   /* plain u64 adds */
      uint32_t best = cp & 0x3ff, bt = 0;
      for (t = 1; t < nt; t++) {
        cp >>= 10;
        if ((cp & 0x3ff) < best) { best = cp & 0x3ff; bt = t; }
      }
      sel[g] = (uint8_t)bt;
      for (i = 0; i < ORC_GROUP; i++) frequency[bt][gs[i]]++;
    }
    for (t = 0; t < nt; t++) huffman_lengths(length[t], frequency[t], as);
  }

  /* renumber trees by first use; finalise the used ones (:1088-1111) */
  uint32_t old2new[ORC_MAX_TREES], new2old[ORC_MAX_TREES], nn = 0, cost = 0;
  uint32_t unseen = (1u << nt) - 1;
  for (t = 0; t < ORC_MAX_TREES; t++) old2new[t] = 0xFF;
  for (g = 0; g < ng && unseen; g++) {
    t = sel[g];
    if (unseen & (1u << t)) {
      unseen &= ~(1u << t);
      old2new[t] = nn; new2old[nn] = t; nn++;
      cost += limited_lengths_and_codes(code[t], length[t], frequency[t], as);
      code[t][as] = 0; length[t][as] = 0;
    }
  }
  if (nn == 1) {                                             /* :1117-1132 */
    uint32_t cl0 = 0;
    while ((2u << cl0) <= as) cl0++;                         /* floor(log2 as) */
    t = new2old[0] ^ 1;
    old2new[t] = 1; new2old[1] = t; nn = 2;
    for (v = 0; v < (2u << cl0) - as; v++) length[t][v] = (uint8_t)cl0;
    if (v < as) cost += 2;
    for (; v < as; v++) length[t][v] = (uint8_t)(cl0 + 1);
    cost += as + 5;
    memset(code[t], 0, sizeof(code[t]));
  }

  memset(out, 0, sizeof *out);
  out->num_trees = nn;
  out->num_groups = ng;
  for (t = 0; t < nn; t++) {
    memcpy(out->length[t], length[new2old[t]], ORC_MAX_ALPHA + 1);
    memcpy(out->code[t], code[new2old[t]], sizeof(code[0]));
  }

  /* selector MTF + size arithmetic of encode() (encode.c:460-544) */
  uint32_t bits = 48 + 32 + 1 + 24 + 3 + 15 + cost;
  uint8_t order[ORC_MAX_TREES] = {0, 1, 2, 3, 4, 5};
  for (g = 0; g < ng; g++) {
    uint8_t c = (uint8_t)old2new[sel[g]];
    uint32_t j = 0;
    while (order[j] != c) j++;
    memmove(order + 1, order, j);
    order[0] = c;
    out->selector[g] = c;
    out->selector_mtf[g] = (uint8_t)j;
    bits += j + 1;
  }
  uint32_t pad = (8 - (bits & 7)) & 7;
  bits += pad;
  out->tree_pad = pad >> 1;
  out->num_selectors = ng + (pad & 1);
  if (pad & 1) out->selector_mtf[ng] = 0;
  for (i = 0; i < 16; i++) {
    uint32_t any = 0;
    for (v = 0; v < 16; v++) any |= used[16 * i + v];
    bits += any ? 16 : 0;
  }
  bits += 16;
  out->out_len = bits >> 3;
}

/* ------------------------------------------------------------ bit writer -- */

struct bitw { uint8_t *p; uint64_t acc; unsigned n; size_t len; };

static void
put(struct bitw *w, unsigned nbits, uint32_t v)
{
  w->acc = (w->acc << nbits) | v;
  w->n += nbits;
  while (w->n >= 8) {
    w->n -= 8;
    w->p[w->len++] = (uint8_t)(w->acc >> w->n);
  }
}

size_t
orc_pack(const struct orc_coding *c, const uint16_t *mtfv, uint32_t nmtf,
         uint32_t as, const uint8_t used[256], uint32_t block_crc,
         uint32_t bwt_idx, uint8_t *out)
{
  struct bitw w = { out, 0, 0, 0 };
  uint32_t i, j, t, v;
  (void)nmtf;

  put(&w, 24, 0x314159); put(&w, 24, 0x265359);             /* :1185-1186 */
  put(&w, 16, (block_crc ^ 0xFFFFFFFFu) >> 16);
  put(&w, 16, (block_crc ^ 0xFFFFFFFFu) & 0xFFFF);
  put(&w, 1, 0);
  put(&w, 24, bwt_idx);

  uint32_t rows = 0, row[16];
  for (i = 0; i < 16; i++) {
    row[i] = 0;
    for (j = 0; j < 16; j++) row[i] = (row[i] << 1) | (used[16 * i + j] != 0);
    rows = (rows << 1) | (row[i] != 0);
  }
  put(&w, 16, rows);
  for (i = 0; i < 16; i++) if (row[i]) put(&w, 16, row[i]);

  put(&w, 3, c->num_trees);
  put(&w, 15, c->num_selectors);
  for (i = 0; i < c->num_selectors; i++) {
    v = c->selector_mtf[i] + 1;
    put(&w, v, (1u << v) - 2);
  }

  for (t = 0; t < c->num_trees; t++) {                       /* :1231-1255 */
    int a = c->length[t][0];
    if (t == 0) a += (a < 4) ? (int)c->tree_pad : -(int)c->tree_pad;
    put(&w, 5, (uint32_t)a);
    for (v = 0; v < as; v++) {
      int l = c->length[t][v];
      while (a < l) { put(&w, 2, 2); a++; }
      while (a > l) { put(&w, 2, 3); a--; }
      put(&w, 1, 0);
    }
  }

  for (i = 0; i < c->num_groups; i++) {                      /* :1258-1272 */
    t = c->selector[i];
    for (j = 0; j < ORC_GROUP; j++) {
      v = mtfv[i * ORC_GROUP + j];
      put(&w, c->length[t][v], c->code[t][v]);
    }
  }
  /* blocks end on a byte boundary by construction (encode.c:1275-1277) */
  return w.len;
}

/* ---------------------------------------------------------- whole block -- */

size_t
orc_encode_block(const uint8_t *in, size_t n, uint32_t cap, uint8_t *out,
                 struct orc_block_info *info)
{
  uint8_t *block = malloc((size_t)cap + 8);
  uint8_t *bwt = malloc((size_t)cap + 8);
  uint16_t *mtfv = malloc(sizeof(uint16_t) * ((size_t)cap + 64));
  uint32_t freq[ORC_MAX_ALPHA + 1];
  uint8_t used[256];
  struct orc_coding *cd = malloc(sizeof *cd);
  struct orc_block_info bi;
  size_t consumed, len = 0;

  memset(&bi, 0, sizeof bi);
  orc_rle1(in, n, cap, block, &bi.nblock, &consumed, used, &bi.block_crc);
  bi.consumed = consumed;
  if (bi.nblock > 0) {
    bi.bwt_idx = orc_bwt(block, bi.nblock, bwt, &bi.tie_count);
    bi.nmtf = orc_mtf(bwt, bi.nblock, used, mtfv, freq, &bi.alpha_size);
    orc_prefix_code(mtfv, bi.nmtf, bi.alpha_size, freq, used, 8, cd);
    bi.num_trees = cd->num_trees;
    bi.num_selectors = cd->num_selectors;
    bi.tree_pad = cd->tree_pad;
    bi.out_len = cd->out_len;
    len = orc_pack(cd, mtfv, bi.nmtf, bi.alpha_size, used, bi.block_crc,
                   bi.bwt_idx, out);
  }
  if (info) *info = bi;
  free(block); free(bwt); free(mtfv); free(cd);
  return len;
}

size_t
orc_stream_bound(size_t n)
{
  return n + n / 50 + 2048 * (n / 100000 + 2) + 64;
}

size_t
orc_compress_stream(const uint8_t *in, size_t n, int level, uint8_t *out,
                    struct orc_block_info *infos, size_t max_infos,
                    size_t *num_blocks)
{
  uint32_t cap = (uint32_t)level * 100000u;
  size_t o = 0, nb = 0, pos = 0;
  uint32_t cc = 0;

  out[o++] = 'B'; out[o++] = 'Z'; out[o++] = 'h'; out[o++] = (uint8_t)('0' + level);
  while (pos < n) {
    /* one I/O chunk of `cap` raw bytes -> 1..2 blocks (compress.c:93-110) */
    size_t chunk_end = pos + cap < n ? pos + cap : n;
    while (pos < chunk_end) {
      struct orc_block_info bi;
      size_t len = orc_encode_block(in + pos, chunk_end - pos, cap, out + o, &bi);
      o += len;
      pos += bi.consumed;
      cc = ((cc << 1) ^ (cc >> 31)) ^ (bi.block_crc ^ 0xFFFFFFFFu);  /* encode.h:38 */
      if (infos && nb < max_infos) infos[nb] = bi;
      nb++;
    }
  }
  static const uint8_t eos[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};
  memcpy(out + o, eos, 6); o += 6;
  out[o++] = (uint8_t)(cc >> 24); out[o++] = (uint8_t)(cc >> 16);
  out[o++] = (uint8_t)(cc >> 8); out[o++] = (uint8_t)cc;
  if (num_blocks) *num_blocks = nb;
  return o;
}
