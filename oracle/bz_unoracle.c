/*
 * bz_unoracle.c -- CPU restatement of lbzip2's per-block DECOMPRESSOR and of
 * the stream walk around it.
 *
 * TEST INFRASTRUCTURE ONLY (see bz_oracle.h).  Written independently of the
 * reference's table-driven decoder: bits are read by absolute position,
 * prefix codes are decoded canonically one length at a time, the inverse MTF
 * is a plain list, the inverse BWT is a counting sort + successor walk and
 * the final run expansion is a two-variable loop.
 *
 * Parity status: PINNED against the compiled reference CLI
 * (oracle/_ref/lbzip2 -d) on every .bz2 fixture of the reference
 * (tests/ and tests/suite/manual-expand) -- accept/reject, error kind and
 * output bytes -- and on round trips of the compress fixtures; see
 * tools/pin_unoracle.py, tests/test_unoracle.py and tests/golden/decode/.
 */
#include "bz_oracle.h"

#include <stdlib.h>
#include <string.h>

#define MAXBLK 900000u

/* ------------------------------------------------------------------ bits */

struct bitr {
  const uint8_t *p;
  size_t nbytes;      /* real bytes                                        */
  uint64_t nbits;     /* bits as the reference sees them: whole 32-bit words,
                         the tail zero-filled (expand.c:840-858)            */
  uint64_t pos;
};

static uint32_t
peekbits(const struct bitr *b, unsigned k)
{
  uint32_t v = 0;
  unsigned i;
  for (i = 0; i < k; i++) {
    uint64_t q = b->pos + i;
    unsigned bit = 0;
    if ((q >> 3) < b->nbytes)
      bit = (b->p[q >> 3] >> (7 - (q & 7))) & 1;
    v = (v << 1) | bit;
  }
  return v;
}

static uint32_t
takebits(struct bitr *b, unsigned k)
{
  uint32_t v = peekbits(b, k);
  b->pos += k;
  return v;
}

/* The reference refills its 64-bit window whenever fewer than 32 bits are
   buffered and reports "unexpected end of file" if it cannot
   (NEED(), decode.c:387-407).  In absolute terms that is: fewer than 32 bits
   of (word-padded) input remain.  */
#define NEED(b) do { if ((b)->nbits - (b)->pos < 32) return ORC_ERR_EOF; } while (0)

/* ------------------------------------------------------------- retrieve */

struct dtree {
  int status;                     /* tree number, or ORC_ERR_PREFIX/INCOMPLT */
  uint32_t first[22];             /* first code of each length              */
  uint32_t count[22];
  uint32_t offset[22];
  uint16_t perm[ORC_MAX_ALPHA];
};

/* decode.c:181-305 make_tree(): Kraft check, canonical code order.  */
static void
build_dtree(struct dtree *t, int tno, const uint8_t *len, unsigned n)
{
  uint64_t kraft = 0;
  uint32_t code = 0, off = 0;
  unsigned k, s;
  memset(t->count, 0, sizeof t->count);
  for (s = 0; s < n; s++) t->count[len[s]]++;
  for (k = 1; k <= 20; k++) kraft += (uint64_t)t->count[k] << (20 - k);
  if (kraft != (1u << 20)) {
    t->status = kraft < (1u << 20) ? ORC_ERR_INCOMPLT : ORC_ERR_PREFIX;
    return;
  }
  for (k = 1; k <= 20; k++) {
    t->first[k] = code;
    t->offset[k] = off;
    code = (code + t->count[k]) << 1;
    off += t->count[k];
  }
  {
    uint32_t fill[22];
    memcpy(fill, t->offset, sizeof fill);
    for (s = 0; s < n; s++) t->perm[fill[len[s]]++] = (uint16_t)s;
  }
  t->status = tno;
}

int
orc_d_retrieve(const uint8_t *in, size_t nbytes, uint64_t bitpos,
               uint8_t *bwt, struct orc_dblock *bi)
{
  struct bitr B;
  uint8_t list[256];
  uint8_t selector[32768];
  struct dtree *tree = NULL;
  unsigned big, nsym = 0, alpha, ntrees, nsel, i, j, t, g;
  unsigned sellist[6];
  uint32_t run = 0, shift = 0, n = 0;
  uint8_t runch;
  int rv;

  memset(bi, 0, sizeof *bi);
  B.p = in; B.nbytes = nbytes; B.nbits = 32 * (uint64_t)((nbytes + 3) / 4);
  B.pos = bitpos;

  /* decode.c:527-531 */
  NEED(&B);
  bi->rand = takebits(&B, 1);
  bi->bwt_idx = takebits(&B, 24);

  /* byte map, decode.c:533-553 */
  NEED(&B);
  big = takebits(&B, 16);
  for (i = 0; i < 16; i++) {
    if (big & (0x8000u >> i)) {
      unsigned small = takebits(&B, 16);
      NEED(&B);
      for (j = 0; j < 16; j++)
        if (small & (0x8000u >> j)) list[nsym++] = (uint8_t)(16 * i + j);
    }
  }
  if (nsym == 0) return bi->status = ORC_ERR_BITMAP;
  alpha = nsym + 2;
  bi->alpha_size = alpha;

  ntrees = takebits(&B, 3);
  bi->num_trees = ntrees;
  if (ntrees < 2 || ntrees > 6) return bi->status = ORC_ERR_TREES;
  nsel = takebits(&B, 15);
  bi->num_selectors = nsel;
  if (nsel == 0) return bi->status = ORC_ERR_GROUPS;

  /* unary selector ranks, decode.c:566-575 (6-bit look-ahead) */
  for (i = 0; i < nsel; i++) {
    uint32_t w = peekbits(&B, 6);
    unsigned k = 1;
    while (k <= 6 && (w & (0x40u >> k))) k++;   /* first zero bit, 7 = none */
    if (k > ntrees) return bi->status = ORC_ERR_SELECTOR;
    selector[i] = (uint8_t)(k - 1);
    B.pos += k;
    NEED(&B);
  }

  /* delta-coded lengths, decode.c:577-601: up to three +-1 steps are taken
     per 6-bit window and the range is checked once per window */
  tree = malloc(6 * sizeof *tree);
  for (t = 0; t < ntrees; t++) {
    uint8_t len[ORC_MAX_ALPHA];
    int cur = (int)takebits(&B, 5);
    j = 0;
    while (j < alpha) {
      uint32_t w = peekbits(&B, 6);
      unsigned used = 0;
      int done = 0;
      while (used + 2 <= 6 && (w & (0x20u >> used))) {
        cur += (w & (0x20u >> (used + 1))) ? -1 : 1;
        used += 2;
      }
      if (used < 6) { used += 1; done = 1; }
      if (cur < 1 || cur > 20) { free(tree); return bi->status = ORC_ERR_DELTA; }
      if (done) len[j++] = (uint8_t)cur;
      B.pos += used;
      NEED(&B);
    }
    build_dtree(&tree[t], (int)t, len, alpha);
  }
  for (t = 0; t < ntrees; t++) sellist[t] = (unsigned)tree[t].status;

  if (nsel > 18001) nsel = 18001;   /* decode.c:631-632 */
  runch = list[0];

  rv = ORC_ERR_UNTERM;
  for (g = 0; g < nsel && rv == ORC_ERR_UNTERM; g++) {
    const struct dtree *T;
    unsigned r = selector[g];
    t = sellist[r];
    if (t >= 6) { rv = (int)t; break; }    /* bad tree used, decode.c:640-642 */
    for (; r > 0; r--) sellist[r] = sellist[r - 1];
    sellist[0] = t;
    T = &tree[t];

    for (j = 0; j < ORC_GROUP; j++) {
      uint32_t code = 0, s = 0;
      unsigned k;
      if (B.nbits - B.pos < 32) { rv = ORC_ERR_EOF; break; }
      for (k = 1; k <= 20; k++) {
        code = (code << 1) | takebits(&B, 1);
        if (code - T->first[k] < T->count[k]) {
          s = T->perm[T->offset[k] + code - T->first[k]];
          break;
        }
      }
      if (s == alpha - 1) {                 /* EOB, decode.c:731-752 */
        if (run > MAXBLK - n) { rv = ORC_ERR_OVERFLOW; break; }
        while (run--) bwt[n++] = runch;
        if (n == 0) rv = ORC_ERR_EMPTY;
        else if (bi->bwt_idx >= n) rv = ORC_ERR_BWTIDX;
        else rv = ORC_OK;
        break;
      }
      if (s < 2 && run <= MAXBLK) {         /* RUNA/RUNB, decode.c:761-764 */
        run += (s + 1) << shift++;
        continue;
      }
      if (run > MAXBLK - n) { rv = ORC_ERR_OVERFLOW; break; }
      while (run--) bwt[n++] = runch;
      {
        unsigned rank = s - 1;
        uint8_t c = list[rank];
        memmove(list + 1, list, rank);
        list[0] = c;
        runch = c;
      }
      shift = 0;
      run = 1;
    }
  }
  free(tree);
  bi->block_size = n;
  bi->end_bit = B.pos;
  return bi->status = rv;
}

/* ----------------------------------------------------------------- IBWT */

static const uint16_t rnums[512] = {
#include "bz_randtab.inc"
};

/* decode.c:840-917 decode(): out = the initial-RLE coded text.  */
void
orc_d_ibwt(const uint8_t *bwt, uint32_t n, uint32_t idx, int rand,
           uint8_t *out)
{
  uint32_t cnt[257], i, x;
  uint32_t *succ = malloc((size_t)(n ? n : 1) * sizeof *succ);
  memset(cnt, 0, sizeof cnt);
  for (i = 0; i < n; i++) cnt[bwt[i] + 1]++;
  for (i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
  /* succ[j] = i : row j of the sorted matrix starts with the byte that ends
     row i (stable), so following succ from idx spells the text forward */
  for (i = 0; i < n; i++) succ[cnt[bwt[i]]++] = i;
  x = succ[idx];
  for (i = 0; i < n; i++) { out[i] = bwt[x]; x = succ[x]; }
  free(succ);
  if (rand) {                    /* decode.c:893-899 */
    uint32_t k = 0, j = 617;
    while (j < n) {
      out[j] ^= 1;
      k = (k + 1) & 511;
      j += rnums[k];
    }
  }
}

/* ---------------------------------------------------------------- unRLE */

/* emit(), decode.c:936-1143: four equal bytes are followed by a repeat count.
   Returns ORC_ERR_RUNLEN if the input ends right after the fourth equal byte,
   ORC_ERR_OVERFLOW (oracle-only) if cap is too small.  *crc is the final,
   inverted CRC.  */
int
orc_d_unrle(const uint8_t *src, uint32_t n, uint8_t *out, size_t cap,
            size_t *out_len, uint32_t *crc)
{
  size_t o = 0;
  uint32_t i = 0, c = 0xFFFFFFFFu;
  unsigned same = 0;
  int prev = -1;
  while (i < n) {
    uint8_t b = src[i++];
    if (o >= cap) return ORC_ERR_OVERFLOW;
    out[o++] = b;
    same = (b == prev) ? same + 1 : 1;
    prev = b;
    if (same == 4) {
      unsigned k;
      if (i >= n) { *out_len = o; return ORC_ERR_RUNLEN; }
      k = src[i++];
      if (k > cap - o) return ORC_ERR_OVERFLOW;
      memset(out + o, b, k);
      o += k;
      same = 0;
      prev = -1;
    }
  }
  c = orc_crc_update(c, out, o);
  *crc = ~c;
  *out_len = o;
  return ORC_OK;
}

/* --------------------------------------------------------------- stream */

static uint32_t
get16(const uint8_t *in, size_t nbytes, uint64_t bit)
{
  struct bitr B;
  B.p = in; B.nbytes = nbytes; B.nbits = 0; B.pos = bit;
  return peekbits(&B, 16);
}

/* main.c:664-683 (first header), parse.c:147-263 parse(), expand.c:395-440
   (EOF rule), :703-741 (per-block checks in stream order).  */
int
orc_decompress_stream(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                      size_t *out_len, struct orc_dstream *si)
{
  uint64_t nbits = 32 * (uint64_t)((n + 3) / 4);
  uint64_t pos;
  uint32_t strm_crc = 0;
  int bs100k;
  size_t o = 0;
  uint8_t *bwt, *txt;
  int rv = ORC_OK;

  memset(si, 0, sizeof *si);
  *out_len = 0;
  if (n < 4 || in[0] != 'B' || in[1] != 'Z' || in[2] != 'h' ||
      in[3] < '1' || in[3] > '9')
    return si->status = ORC_ERR_MAGIC;
  bs100k = in[3] - '0';
  pos = 32;
  bwt = malloc(MAXBLK);
  txt = malloc(MAXBLK);

#define UNIT(v) do { if (nbits - pos < 16) { rv = ORC_ERR_EOF; goto done; } \
                     (v) = get16(in, n, pos); pos += 16; } while (0)
  for (;;) {
    uint32_t w, stored;
    UNIT(w);
    if (w == 0x1772) {                              /* end of stream */
      uint64_t q;
      UNIT(w); if (w != 0x4538) { rv = ORC_ERR_HEADER; break; }
      UNIT(w); if (w != 0x5090) { rv = ORC_ERR_HEADER; break; }
      UNIT(stored); UNIT(w); stored = (stored << 16) | w;
      if (stored != strm_crc) { rv = ORC_ERR_STRMCRC; break; }
      si->num_streams++;
      strm_crc = 0;
      pos = (pos + 7) & ~(uint64_t)7;
      /* next stream header, trailing garbage or clean end of file: whatever
         was accepted as data must lie inside the real file (expand.c:428-436) */
      q = pos;
      if (nbits - pos >= 16 && get16(in, n, pos) == 0x425A &&
          nbits - pos >= 32 && (w = get16(in, n, pos + 16)) >= 0x6831 && w <= 0x6839) {
        bs100k = (int)(w & 15);
        pos += 32;
        continue;
      }
      if (q > 8 * (uint64_t)n) rv = ORC_ERR_EOF;
      else if (q < 8 * (uint64_t)n) si->garbage = 1;
      break;
    }
    if (w != 0x3141) { rv = ORC_ERR_HEADER; break; }
    UNIT(w); if (w != 0x5926) { rv = ORC_ERR_HEADER; break; }
    UNIT(w); if (w != 0x5359) { rv = ORC_ERR_HEADER; break; }
    UNIT(stored); UNIT(w); stored = (stored << 16) | w;
    {
      struct orc_dblock bi;
      uint32_t crc = 0;
      size_t got = 0;
      strm_crc = ((strm_crc << 1) | (strm_crc >> 31)) ^ stored;
      rv = orc_d_retrieve(in, n, pos, bwt, &bi);
      /* expand.c:725-726 tests the size before it looks at the block's status, so
         an over-long block whose only other fault is its primary index (the one
         retrieve() error that is raised after the size is known, decode.c:741-747)
         is reported as an overflow */
      if ((rv == ORC_OK || rv == ORC_ERR_BWTIDX) &&
          bi.block_size > (uint32_t)bs100k * 100000u) rv = ORC_ERR_OVERFLOW;
      if (rv == ORC_OK)
        orc_d_ibwt(bwt, bi.block_size, bi.bwt_idx, (int)bi.rand, txt);
      if (rv == ORC_OK) {
        rv = orc_d_unrle(txt, bi.block_size, out + o, cap - o, &got, &crc);
        if (rv == ORC_ERR_OVERFLOW) rv = ORC_ERR_OUTCAP;
      }
      if (rv == ORC_OK && crc != stored) rv = ORC_ERR_BLKCRC;
      if (rv != ORC_OK) { si->bad_block = si->num_blocks; break; }
      o += got;
      si->num_blocks++;
      pos = bi.end_bit;
    }
  }
done:
  free(bwt); free(txt);
  *out_len = o;
  si->end_bit = pos;
  return si->status = rv;
}
