/*
 * bz_oracle.h -- CPU restatement of lbzip2's per-block compressor.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (lbzip2_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  Every stage below is checked bit-for-bit against
 * the compiled, unmodified reference (oracle/_ref/libref_stages.so, built by
 * oracle/Makefile from /root/reference/src) on the reference's own fixture
 * corpus (tests/suite/{fuzz-collect,fuzz-divbwt,manual-compress}) and on
 * synthetic inputs; see tests/test_oracle_vs_ref.py and tests/golden/.
 * One documented exception: for exactly periodic blocks (block == w^k, k>=2)
 * the BWT primary index is ambiguous (reference tests/incomp:4-18); the
 * reference's choice falls out of divsufsort's internal sort order
 * (divbwt.c:1468-1477).  orc_bwt() reports the tie group so callers can tell.
 */
#ifndef BZ_ORACLE_H
#define BZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TREES 6
#define ORC_MAX_ALPHA 258
#define ORC_GROUP 50
#define ORC_MAX_SELECTORS 18002

/* CRC-32/BZIP2 (poly 0x04C11DB7, MSB first), raw update without final xor.
   Follows encode.c:103 + crctab.c:26 (table regenerated, not copied).  */
uint32_t orc_crc_update(uint32_t crc, const uint8_t *p, size_t n);

/* Initial run-length coding of ONE block out of a raw buffer.
   Follows collect() encode.c:135-336 and the run flush in encode()
   encode.c:443-447, for a fresh state fed by a single collect() call.
   Outputs: block[0..*nblock), *consumed raw bytes, used[256] byte map,
   *crc = un-inverted CRC of the consumed bytes, returns 1 iff the block
   closed because it was full (collect()'s return value).  */
int orc_rle1(const uint8_t *in, size_t n, uint32_t cap, uint8_t *block,
             uint32_t *nblock, size_t *consumed, uint8_t used[256],
             uint32_t *crc);

/* BWT under cyclic-rotation order (contract of divbwt(), divbwt.c:1707-1726).
   Returns the primary index.  If the block is exactly periodic the index is
   the FIRST position of the tie group and *tie_count (if non-NULL) receives
   the number of equal rotations (1 for aperiodic blocks).  */
uint32_t orc_bwt(const uint8_t *t, uint32_t n, uint8_t *bwt,
                 uint32_t *tie_count);

/* MTF + zero-run coding + histogram (do_mtf(), encode.c:360-425, with the
   dense renumbering of make_map_e(), encode.c:340-355).  mtfv must hold
   n+1 entries (+50 slack for the group padding added later).
   Returns nmtf; *alpha_size = EOB+1.  freq must hold 259 entries.  */
uint32_t orc_mtf(const uint8_t *bwt, uint32_t n, const uint8_t used[256],
                 uint16_t *mtfv, uint32_t *freq, uint32_t *alpha_size);

struct orc_coding {
  uint32_t num_trees;                    /* after reordering / dummy tree */
  uint32_t num_groups;                   /* real groups = ceil(nmtf/50)    */
  uint32_t num_selectors;                /* + optional padding selector    */
  uint32_t tree_pad;                     /* 0..3 dummy delta pairs          */
  uint32_t out_len;                      /* block bytes                     */
  uint8_t  length[ORC_MAX_TREES][ORC_MAX_ALPHA + 1]; /* NEW tree order       */
  uint32_t code[ORC_MAX_TREES][ORC_MAX_ALPHA + 1];
  uint8_t  selector[ORC_MAX_SELECTORS];              /* NEW tree numbers    */
  uint8_t  selector_mtf[ORC_MAX_SELECTORS + 8];
};

/* Multi-table prefix code construction (generate_prefix_code()
   encode.c:1005-1137 and the selector MTF / padding / size arithmetic of
   encode() encode.c:460-544).  mtfv must have room for the 50-symbol group
   padding.  */
void orc_prefix_code(uint16_t *mtfv, uint32_t nmtf, uint32_t alpha_size,
                     const uint32_t *freq, const uint8_t used[256],
                     unsigned cluster_factor, struct orc_coding *out);

/* Bit serialisation of one block (transmit(), encode.c:1152-1281).
   Writes exactly c->out_len bytes; returns that number.  */
size_t orc_pack(const struct orc_coding *c, const uint16_t *mtfv,
                uint32_t nmtf, uint32_t alpha_size, const uint8_t used[256],
                uint32_t block_crc, uint32_t bwt_idx, uint8_t *out);

struct orc_block_info {
  uint64_t consumed;
  uint32_t nblock, block_crc, bwt_idx, tie_count, nmtf, alpha_size;
  uint32_t num_trees, num_selectors, tree_pad, out_len;
};

/* One block end to end.  Returns bytes written to out (0 if n == 0).  */
size_t orc_encode_block(const uint8_t *in, size_t n, uint32_t cap,
                        uint8_t *out, struct orc_block_info *info);

/* A whole .bz2 stream the way `lbzip2 -<level>` frames it (compress.c:73-117
   chunking, :290-321 header/trailer, encode.h:38 CRC fold).
   out must hold orc_stream_bound(n) bytes.  Returns bytes written.
   If infos != NULL it receives up to max_infos per-block records and
   *num_blocks the block count.  */
size_t orc_stream_bound(size_t n);
size_t orc_compress_stream(const uint8_t *in, size_t n, int level,
                           uint8_t *out, struct orc_block_info *infos,
                           size_t max_infos, size_t *num_blocks);

/* ------------------------------------------------------------------------
 * Decompression side (bz_unoracle.c): restatement of decode.c / parse.c and
 * of the per-block checks of expand.c.  Status values are the reference's
 * `enum error` (common.h:54-76) in the same order.
 */
enum {
  ORC_OK = 0, ORC_MORE, ORC_FINISH,
  ORC_ERR_MAGIC, ORC_ERR_HEADER, ORC_ERR_BITMAP, ORC_ERR_TREES, ORC_ERR_GROUPS,
  ORC_ERR_SELECTOR, ORC_ERR_DELTA, ORC_ERR_PREFIX, ORC_ERR_INCOMPLT,
  ORC_ERR_EMPTY, ORC_ERR_UNTERM, ORC_ERR_RUNLEN, ORC_ERR_BLKCRC,
  ORC_ERR_STRMCRC, ORC_ERR_OVERFLOW, ORC_ERR_BWTIDX, ORC_ERR_EOF,
  ORC_ERR_OUTCAP = 100          /* ours: caller's output buffer too small */
};

struct orc_dblock {
  uint32_t status, rand, bwt_idx, block_size, alpha_size, num_trees,
           num_selectors, pad;
  uint64_t end_bit;             /* first bit after the block's EOB symbol */
};

struct orc_dstream {
  uint32_t status, num_blocks, num_streams, bad_block, garbage, pad;
  uint64_t end_bit;
};

/* retrieve() decode.c:518-791: header fields, byte map, selectors, code
   lengths, prefix decoding, inverse MTF and zero-run expansion of ONE block
   whose payload starts at absolute bit `bitpos` (right after the 32-bit block
   CRC).  bwt must hold 900000 bytes.  Returns the status.  */
int orc_d_retrieve(const uint8_t *in, size_t nbytes, uint64_t bitpos,
                   uint8_t *bwt, struct orc_dblock *bi);

/* decode() decode.c:840-917: inverse BWT (+ de-randomisation).  */
void orc_d_ibwt(const uint8_t *bwt, uint32_t n, uint32_t idx, int rand,
                uint8_t *out);

/* emit() decode.c:936-1143: undo the initial run-length coding, CRC.  */
int orc_d_unrle(const uint8_t *src, uint32_t n, uint8_t *out, size_t cap,
                size_t *out_len, uint32_t *crc);

/* Whole file: first header (main.c:664-683), parse() parse.c:147-263,
   EOF rule expand.c:428-436, per-block checks expand.c:725-736.
   On error *out_len counts the bytes of the blocks before the bad one.  */
int orc_decompress_stream(const uint8_t *in, size_t n, uint8_t *out,
                          size_t cap, size_t *out_len,
                          struct orc_dstream *si);

#ifdef __cplusplus
}
#endif
#endif
