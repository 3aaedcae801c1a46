/*
 * ref_stages.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin harness around the UNMODIFIED reference sources, compiled where they
 * lie under /root/reference (see oracle/Makefile, target _ref/libref_stages.so).
 * It textually includes the reference's src/encode.c at build time so that the
 * per-stage state (struct encoder_state, file-static helpers) is reachable,
 * and dumps the intermediate results of ONE block so that both the oracle
 * restatement (oracle/bz_oracle.c) and the CUDA path can be pinned against
 * the real thing stage by stage.
 *
 * Nothing from the reference is copied into this repository: the include is
 * resolved by the compiler from REFERENCE_SRC at build time, and the outputs
 * live under oracle/_ref/ (git-ignored).
 */
#define _XOPEN_SOURCE 700
#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include REF_ENCODE_C           /* "/root/reference/src/encode.c" */

struct ref_block_dump {
  uint64_t consumed;            /* raw bytes eaten by collect() */
  uint32_t full;                /* collect() return value */
  uint32_t nblock;              /* n' after encode()'s run finalisation */
  uint32_t block_crc;           /* un-inverted, as encode() hands it back */
  uint32_t bwt_idx;
  uint32_t nmtf;
  uint32_t alpha_size;          /* EOB + 1 */
  uint32_t num_selectors;       /* incl. the optional padding selector */
  uint32_t num_trees;
  uint32_t tree_pad;
  uint32_t out_len;             /* bytes, == encode() return */
  uint8_t  used[256];
  uint8_t  tmap_new2old[6];
  uint8_t  length[6][259];      /* indexed by OLD tree number */
  uint32_t code[6][259];
};

/* Encode one block through the reference API (encode.h:29-36):
   encoder_init -> collect (once) -> encode -> transmit.
   Optional output arrays may be NULL.  Returns 0 on success.  */
int
ref_encode_block(const uint8_t *in, uint64_t in_len, uint32_t mbs,
                 struct ref_block_dump *d,
                 uint8_t *block_out,     /* >= mbs+1 bytes: RLE1 output   */
                 uint8_t *bwt_out,       /* >= mbs bytes: BWT last column */
                 uint16_t *mtfv_out,     /* >= mbs+51 entries             */
                 uint8_t *selector_out,  /* >= 18002 (old tree numbers)   */
                 uint8_t *selmtf_out,    /* >= 18008                      */
                 uint8_t *bits_out)      /* >= out_len rounded up to 4    */
{
  struct encoder_state *s = malloc(encoder_alloc_size(mbs));
  size_t left = in_len;
  uint32_t crc;
  uint32_t i;
  uint8_t *block;

  if (!s) return -1;
  encoder_init(s, mbs, CLUSTER_FACTOR);
  d->full = collect(s, in, &left);
  d->consumed = in_len - left;

  /* BWT bytes: run divbwt on a scratch copy of the (finalised) block, so the
     real encode() below still sees pristine state.  divbwt is a pure
     function of (T, n) (divbwt.c:1707).  */
  block = (uint8_t *)(s->SA + s->max_block_size + GROUP_SIZE);
  {
    uint32_t nb = s->nblock;
    uint8_t *t = malloc((size_t)mbs + 8);
    int32_t *sa = malloc(((size_t)mbs + 64) * sizeof(int32_t));
    int32_t *bk = malloc((65536 + 256) * sizeof(int32_t));
    if (!t || !sa || !bk) return -1;
    memcpy(t, block, nb);
    if (s->rle_state >= 4)
      t[nb++] = s->rle_state - 4;     /* what encode() will append (encode.c:443-447) */
    if (nb > 0 && bwt_out) {
      (void)divbwt(t, sa, bk, nb);
      for (i = 0; i < nb; i++) bwt_out[i] = (uint8_t)sa[i];
    }
    free(t); free(sa); free(bk);
  }

  if (s->nblock == 0 && s->rle_state < 4) { free(s); return 1; }  /* empty input */

  d->out_len = (uint32_t)encode(s, &crc);
  d->nblock = s->nblock;
  d->block_crc = crc;
  d->bwt_idx = s->bwt_idx;
  d->nmtf = s->nmtf;
  d->num_selectors = s->u.s.num_selectors;
  d->num_trees = s->u.s.num_trees;
  d->tree_pad = s->u.s.tree_pad;
  for (i = 0; i < 256; i++) d->used[i] = s->cmap[i];
  {
    uint16_t *mtfv = (uint16_t *)s->SA;
    d->alpha_size = mtfv[s->nmtf - 1] + 1;
    if (mtfv_out) memcpy(mtfv_out, mtfv, s->nmtf * sizeof(uint16_t));
  }
  if (block_out) memcpy(block_out, block, s->nblock);
  for (i = 0; i < 6; i++) d->tmap_new2old[i] = (uint8_t)s->u.s.tmap_new2old[i];
  memcpy(d->length, s->u.s.length, sizeof(d->length));
  memcpy(d->code, s->u.s.code, sizeof(d->code));
  if (selector_out) memcpy(selector_out, s->u.s.selector, 18002);
  if (selmtf_out) memcpy(selmtf_out, s->u.s.selectorMTF, 18008);
  if (bits_out) transmit(s, bits_out);
  free(s);
  return 0;
}

/* Straight BWT through the reference's divbwt (divbwt.c:1707). */
int32_t
ref_divbwt(const uint8_t *T, int32_t n, uint8_t *bwt_out)
{
  uint8_t *t = malloc((size_t)n + 8);
  int32_t *sa = malloc(((size_t)n + 64) * sizeof(int32_t));
  int32_t *bk = malloc((65536 + 256) * sizeof(int32_t));
  int32_t i, pidx;
  memcpy(t, T, n);
  pidx = divbwt(t, sa, bk, n);
  for (i = 0; i < n; i++) bwt_out[i] = (uint8_t)sa[i];
  free(t); free(sa); free(bk);
  return pidx;
}

uint32_t ref_crc_table_entry(unsigned i) { return crc_table[i & 255]; }
