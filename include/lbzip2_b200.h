/*
 * lbzip2_b200.h -- C ABI of the B200-native bzip2 block-compression engine.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types.  The library
 * (lbzip2_b200/libbz2b200.so) implements everything below with hand-written
 * sm_100a CUDA kernels and has NO CPU fallback: every entry point fails
 * loudly (negative return / abort for the reference-shaped calls, which have
 * no error channel) when no usable GPU is present.
 *
 * Three groups of entry points:
 *
 *  1. The reference's per-block encoder API, verbatim -- what the unmodified
 *     scheduler (reference src/compress.c:73-117,210-228) calls.  Linking the
 *     reference's main.c/process.c/compress.c/... against this library instead
 *     of src/encode.c + src/divbwt.c gives a GPU-driven `lbzip2` binary
 *     (recipe: oracle/Makefile target _ref/lbzip2_gpu).
 *
 *  2. A batch API (many chunks per call) -- what a batch-aware compress.c
 *     would call and what bench.py measures.  One call pushes
 *     a whole batch of <=900 kB chunks through the kernels at once, which is
 *     what fills a B200.
 *
 *  3. Stage-level debug hooks used only by the parity tests (tests/), so that
 *     every kernel can be checked against the oracle on the oracle's own
 *     stage inputs.
 */
#ifndef LBZIP2_B200_H
#define LBZIP2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * 1. Reference-compatible per-block API.
 *    Replaces: reference src/encode.h:22-38 (implemented in src/encode.c,
 *    src/divbwt.c).  Names, argument meaning and return conventions are the
 *    reference's.  `struct encoder_state` is opaque to callers there too
 *    (src/encode.h:27), so its layout is ours: a small host-side handle that
 *    names a pooled device context.
 * ---------------------------------------------------------------------- */
#define CLUSTER_FACTOR 8u        /* src/encode.h:22 */
#define HEADER_SIZE 4u           /* src/encode.h:23 */
#define TRAILER_SIZE 10u         /* src/encode.h:24 */

struct encoder_state;

/* src/encode.h:29  (src/encode.c:108-114).  Bytes the caller must allocate. */
size_t encoder_alloc_size(unsigned long max_block_size);
/* src/encode.h:30  (src/encode.c:117-132). */
void encoder_init(struct encoder_state *e, unsigned long max_block_size, unsigned cluster_factor);
/* src/encode.h:31  (src/encode.c:135-336).  Consumes input until the block is
   full or the buffer is empty; *buf_sz becomes the number of bytes LEFT;
   returns 1 iff the block is full.  Resumable across calls like the reference. */
int collect(struct encoder_state *e, const uint8_t *buf, size_t *buf_sz);
/* src/encode.h:32  (src/encode.c:427-545).  Returns the byte size of the block
   and hands back the un-inverted block CRC. */
size_t encode(struct encoder_state *e, uint32_t *crc);
/* src/encode.h:33  (src/encode.c:1152-1281).  Writes the block into buf
   (>= (size+3)/4*4 bytes).  Releases the device context: it is the last call
   the scheduler makes before free(e) (src/compress.c:220-223). */
void *transmit(struct encoder_state *e, void *buf);
/* src/encode.h:34  (src/encode.c:1005-1137).  In the reference this is a step inside
   encode() (its only caller, src/encode.c:469) that builds the prefix codes of the block
   held by the state and returns their transmission cost in bits.  Here the whole block
   runs on the device in one go, so the call is valid after encode() and returns that
   cost as computed by the Huffman kernel; before encode() it stops loudly. */
unsigned generate_prefix_code(struct encoder_state *s);
/* src/encode.h:36  (src/divbwt.c:1707-1726).  SA[i] receives the BWT byte
   widened to int32, returns the primary index; `bucket` is unused scratch. */
int32_t divbwt(uint8_t *T, int32_t *SA, int32_t *bucket, int32_t n);

#define combine_crc(cc, c) (((cc) << 1) ^ ((cc) >> 31) ^ (c) ^ -1)   /* src/encode.h:38 */

/* ------------------------------------------------------------------------
 * 2. Batch API.
 * ---------------------------------------------------------------------- */
typedef struct lbz_engine lbz_engine;

typedef struct lbz_block_rec {
  uint64_t raw_offset;   /* offset of the block's raw bytes in the call's input   */
  uint32_t raw_len;      /* raw bytes covered                                      */
  uint32_t nblock;       /* n' after the initial RLE                               */
  uint32_t crc;          /* un-inverted block CRC (what encode() returns via *crc) */
  uint32_t bwt_idx;
  uint32_t tie_count;    /* > 1: exactly periodic block (ambiguous primary index)  */
  uint32_t nmtf;
  uint32_t num_trees;
  uint32_t num_selectors;
  uint32_t out_len;      /* bytes of this block in the output                      */
  uint32_t reserved;
} lbz_block_rec;

/* Create an engine on CUDA device `device` for bzip2 level 1..9 (chunk size =
   level*100000 raw bytes, reference src/process.c:631) able to hold
   `max_chunks` chunks per batch.  Returns NULL (and prints why) on failure. */
lbz_engine *lbz_engine_create(int device, int level, int max_chunks);
void lbz_engine_destroy(lbz_engine *e);

/* Upper bound for the output of n raw bytes (blocks only / whole stream). */
size_t lbz_bound(size_t n);

/* Compress n raw bytes of HOST memory, cut into chunks of level*100000 bytes
   exactly like the reference scheduler does (src/compress.c:93-110): writes
   the concatenated, byte-aligned block bitstreams in stream order (no stream
   header/trailer) and one record per block.  Input is processed in batches of
   max_chunks.  Returns 0 on success. */
int lbz_compress_chunks(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                        size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs);

/* Same, but input and output live in DEVICE memory of the engine's GPU
   (d_in: n raw bytes; d_out: >= lbz_bound(n) bytes).  n must fit one batch.
   Nothing is copied to the host except the block records.  Used by bench.py
   for the HBM-resident figure and by the multi-GPU gather. */
int lbz_compress_chunks_device(lbz_engine *e, const void *d_in, size_t n, void *d_out, size_t out_cap,
                               size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs);

/* A complete .bz2 stream, bit-identical to `lbzip2 -<level>` of the same
   input: "BZh"+level, blocks, 0x177245385090, combined CRC
   (src/compress.c:290-321, src/encode.h:38).  Returns 0 on success. */
int lbz_compress_stream(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                        size_t *out_len);

/* Pinned host memory helpers (what the reader thread should fill). */
void *lbz_host_alloc(size_t bytes);
void lbz_host_free(void *p);

/* Kernel launches issued by this engine since creation (for bench.py). */
uint64_t lbz_engine_launches(const lbz_engine *e);
/* Prefix-doubling rounds of the last batch (diagnostics). */
uint32_t lbz_engine_last_rounds(const lbz_engine *e);
/* Bytes of device memory held by the engine. */
size_t lbz_engine_device_bytes(const lbz_engine *e);
const char *lbz_version(void);
/* Device timing of the last lbz_compress_* call, from CUDA events recorded on the
   engine's own stream: whole call; per stage {rle1, initial radix sort, doubling
   rounds, last-column gather, mtf, huffman, pack+gather}; and the dominant kernel
   (one LSD pass of the initial sort, k_scatter): summed launch time, launch count,
   elements moved per launch. */
double lbz_engine_last_ms(const lbz_engine *e);
void lbz_engine_stage_ms(const lbz_engine *e, double *out7);
void lbz_engine_k0_stats(const lbz_engine *e, double *sum_ms, uint32_t *launches, uint64_t *elements);

/* ------------------------------------------------------------------------
 * 3. Stage hooks for the parity tests.  Slots: chunk c owns block slots 2c
 *    and 2c+1 (see lbzip2_b200/csrc/lbz_common.cuh).
 * ---------------------------------------------------------------------- */
enum lbz_stage { LBZ_ST_RLE1 = 0, LBZ_ST_BWT = 1, LBZ_ST_MTF = 2, LBZ_ST_HUFFMAN = 3, LBZ_ST_PACK = 4 };
enum lbz_array {
  LBZ_AR_TEXT = 0,     /* u8  RLE1 output of a slot            */
  LBZ_AR_BWT = 1,      /* u8  last column                      */
  LBZ_AR_MTFV = 2,     /* u16 symbol stream                    */
  LBZ_AR_FREQ = 3,     /* u32[260] symbol histogram            */
  LBZ_AR_CODING = 4,   /* struct LbzCoding                     */
  LBZ_AR_OUT = 5,      /* u8  packed block                     */
  LBZ_AR_META = 6,     /* struct LbzBlockMeta                  */
  LBZ_AR_SA = 7        /* u32 rotation order                   */
};
int lbz_dbg_load(lbz_engine *e, const uint8_t *in, size_t n);          /* H2D + chunk table */
int lbz_dbg_run(lbz_engine *e, int stage);
int lbz_dbg_read(lbz_engine *e, int array, uint32_t slot, void *dst, size_t bytes);
int lbz_dbg_write(lbz_engine *e, int array, uint32_t slot, const void *src, size_t bytes);
uint32_t lbz_dbg_num_slots(const lbz_engine *e);
int lbz_dbg_set_chunks(lbz_engine *e, uint32_t nchunks);   /* slots are then filled with lbz_dbg_write */

#ifdef __cplusplus
}
#endif
#endif /* LBZIP2_B200_H */
