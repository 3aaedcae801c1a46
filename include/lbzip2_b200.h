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
   (>= (size+3)/4*4 bytes) and returns buf; with buf == NULL the block is left in
   the state's own memory and a pointer to it is returned (src/encode.c:1177-1182).
   Releases the device context: it is the last call the scheduler makes before
   free(e) (src/compress.c:220-223). */
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

/* src/decode.h:70  (src/crctab.c:6).  CRC-32/BZIP2 table, poly 0x04C11DB7, MSB first. */
extern uint32_t crc_table[256];

/* The encode side of the reference API has no error channel; unrecoverable errors end in
   the host program's failx() (src/main.h:81-88) when the program defines it -- the library
   holds a weak reference -- so that the CLI's cleanup of partial output (src/main.c:60-74)
   runs; a handler installed here takes precedence; with neither, abort(). */
void lbz_set_fatal_handler(void (*fn)(const char *msg));

/* ------------------------------------------------------------------------
 * 1b. Reference-compatible per-block DECODER API.
 *    Replaces: reference src/decode.h:72-81 as implemented in src/decode.c
 *    (decoder_init :1148, decoder_free :1159, retrieve :519, decode :852,
 *    emit :944).  `struct decoder_state` and `struct bitstream` are the
 *    reference's own public layouts (src/decode.h:39-66): callers compile
 *    against the reference's decode.h and link this library instead of
 *    src/decode.c (oracle/Makefile: _ref/lbzip2_gpu).  The
 *    stream parser and the block scanner (parser_init/parse/scan,
 *    src/parse.c) are host-side framing logic and stay the reference's.
 *    Return values are the reference's `enum error` (src/common.h:54-76).
 * ---------------------------------------------------------------------- */
struct decoder_state;
struct bitstream;
/* src/decode.h:77.  Allocates the per-block state behind ds->internal_state. */
void decoder_init(struct decoder_state *ds);
/* src/decode.h:78. */
void decoder_free(struct decoder_state *ds);
/* src/decode.h:79  (src/decode.c:519-850).  Consumes the bits offered by `bs`
   (one I/O buffer at a time); MORE = the block continues in the next buffer,
   OK = block complete with the cursor on its last bit, else the error. */
int retrieve(struct decoder_state *ds, struct bitstream *bs);
/* src/decode.h:80  (src/decode.c:852-931).  Inverse BWT; here it already ran
   with the rest of the block when retrieve() returned OK. */
void decode(struct decoder_state *ds);
/* src/decode.h:81  (src/decode.c:944-1143).  Writes decoded bytes into buf;
   *buf_sz: in = capacity, out = space left; MORE, OK (ds->crc = block CRC)
   or ERR_RUNLEN. */
int emit(struct decoder_state *ds, void *buf, size_t *buf_sz);

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

/* Same with HOST input and DEVICE output: the multi-process host sink.  Every rank of a sharded
   job (chunk i -> rank i mod N) keeps its blocks in HBM until the block tables of
   all ranks are known, then copies each block to its offset in the ONE stream-ordered output
   (a host buffer shared between the ranks) with lbz_scatter_to_host: the order key is the
   reference's (major = chunk, minor = block in chunk), src/compress.c:85-86,238-252. */
int lbz_compress_chunks_h2d(lbz_engine *e, const uint8_t *in, size_t n, void *d_out, size_t out_cap,
                            size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs);
/* count device->host copies: len[i] bytes from d_src + src_off[i] to h_dst + dst_off[i]; returns
   when all have landed. */
int lbz_scatter_to_host(lbz_engine *e, const void *d_src, const uint64_t *src_off, void *h_dst,
                        const uint64_t *dst_off, const uint64_t *len, size_t count);

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

/* ------------------------------------------------------------------------
 * 4. Batch DECOMPRESSION.
 *    Replaces, for a whole file at a time, the reference's scan()
 *    (src/parse.c:281-342, src/decode.h:76), retrieve()/decode()/emit()
 *    (src/decode.h:78-81, src/decode.c:518-1143) and the stream walk of
 *    parse() (src/parse.c:147-263) with the per-block checks of
 *    src/expand.c:725-736.  This is the fast path: many whole blocks per
 *    launch.  The per-block decode.h calls themselves are offered too
 *    (section 1b: the unmodified src/expand.c drives them), but retrieve()
 *    is a resumable bit-serial automaton fed 256 KiB at a time, which the
 *    device can only honour one block per call -- a device version that
 *    performs has to see many whole blocks
 *    at once.  Status values are the reference's
 *    `enum error` (src/common.h:54-76), same numbering.
 * ---------------------------------------------------------------------- */
enum lbz_status {
  LBZ_OK = 0, LBZ_MORE = 1, LBZ_FINISH = 2,
  LBZ_ERR_MAGIC = 3, LBZ_ERR_HEADER = 4, LBZ_ERR_BITMAP = 5, LBZ_ERR_TREES = 6, LBZ_ERR_GROUPS = 7,
  LBZ_ERR_SELECTOR = 8, LBZ_ERR_DELTA = 9, LBZ_ERR_PREFIX = 10, LBZ_ERR_INCOMPLT = 11,
  LBZ_ERR_EMPTY = 12, LBZ_ERR_UNTERM = 13, LBZ_ERR_RUNLEN = 14, LBZ_ERR_BLKCRC = 15,
  LBZ_ERR_STRMCRC = 16, LBZ_ERR_OVERFLOW = 17, LBZ_ERR_BWTIDX = 18, LBZ_ERR_EOF = 19,
  LBZ_ERR_OUTCAP = 100,      /* ours: the caller's output buffer is too small */
  LBZ_NEED_INPUT = 101       /* ours: lbz_decoder_next of a streaming session wants lbz_decoder_feed first */
};
/* The reference's message for a status (src/expand.c:70-94, src/process.c:680). */
const char *lbz_strerror(int status);

typedef struct lbz_decoder lbz_decoder;

/* One candidate block of a wave (mirror of struct UbBlock, csrc/unbz_kernels.cuh). */
typedef struct lbz_dblock {
  uint64_t pos, end_bit, out_len, out_off;
  uint32_t status, rand, bwt_idx, block_size, alpha_size, num_trees, num_selectors;
  uint32_t period, rl_state, crc_acc, crc, ntok, nsym, ngrp;
  uint64_t sym_bit;
} lbz_dblock;

typedef struct lbz_dstream_info {
  uint32_t status;           /* same as the return value (>= 0)                        */
  uint32_t num_blocks;       /* blocks decoded and verified                            */
  uint32_t num_streams;      /* concatenated streams completed                         */
  uint32_t bad_block;        /* index of the block the error belongs to                */
  uint32_t garbage;          /* 1: trailing garbage after the last stream was ignored  */
  uint32_t candidates;       /* block magics found by the scanner (any bit offset)     */
  uint32_t false_candidates; /* ... that turned out to lie inside other blocks         */
  uint32_t waves;            /* batches of candidates pushed through the kernels       */
  uint64_t end_bit;          /* where the stream walk stopped                          */
} lbz_dstream_info;

/* A decoder on CUDA device `device` that works on waves of up to `max_blocks`
   candidate blocks (about 8 MB of device memory each), accepts compressed
   inputs of up to `in_cap` bytes and stages up to `out_cap` decoded bytes per
   wave (>= 47 MB so that any single block fits).  NULL (and a message) on
   failure; there is no CPU path. */
/* A decoder is used by one thread at a time; several decoders may run concurrently. */
lbz_decoder *lbz_decoder_create(int device, int max_blocks, size_t in_cap, size_t out_cap);
void lbz_decoder_destroy(lbz_decoder *d);

/* Decompress a whole .bz2 file (concatenated streams, trailing garbage and
   bit-aligned blocks of foreign compressors included) from HOST memory into
   HOST memory.  Returns LBZ_OK or the reference's error kind; negative for
   CUDA / capacity failures.  On error *out_len counts the bytes of the blocks
   that precede the bad one (those bytes are valid). */
int lbz_decompress_stream(lbz_decoder *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                          size_t *out_len, lbz_dstream_info *info);

/* Variants for measurements and pipelines: LBZ_D_RESIDENT_INPUT = the same n
   bytes were already placed on the device by lbz_decoder_load (the host copy
   is still needed for the framing walk, 10 bytes per block);
   LBZ_D_DEVICE_OUTPUT = leave the decoded bytes in the decoder's device
   buffer (out may be NULL; only the last wave's bytes remain readable with
   lbz_decoder_read). */
#define LBZ_D_RESIDENT_INPUT 1u
#define LBZ_D_DEVICE_OUTPUT 2u
int lbz_decoder_load(lbz_decoder *d, const uint8_t *in, size_t n);
int lbz_decompress_ex(lbz_decoder *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                      size_t *out_len, lbz_dstream_info *info, unsigned flags);

/* The same work one wave at a time, for callers that stream the output (the
   expansion task graph, lbzip2_b200/host/expand_b200.c): open = upload, scan,
   first framing; every next call decodes one wave of up to max_blocks
   candidates into `out` and returns LBZ_MORE while blocks remain, LBZ_OK at
   the clean end of the file, or the error kind -- in every case *out_len bytes
   of `out` are valid.  LBZ_ERR_OUTCAP: not even the next block fits out_cap
   (a single block can expand to 46.6 MB).  `in` must stay valid until the
   last call. */
int lbz_decoder_open(lbz_decoder *d, const uint8_t *in, size_t n, unsigned flags);
int lbz_decoder_next(lbz_decoder *d, uint8_t *out, size_t out_cap, size_t *out_len, lbz_dstream_info *info);

/* The same for a file that ARRIVES IN PIECES (a pipe, a file larger than memory): only a window of
   at most in_cap compressed bytes is resident, on the device for the kernels and mirrored on the
   host for the framing walk -- memory is bounded by the decoder's capacities, not by the file, and
   reading overlaps decoding.  open_stream starts the session; feed appends up to n bytes (*taken of
   them were accepted: the window is full until further waves have consumed its front; the consumed
   front is dropped when room is needed) and says with eof != 0 that these were the last bytes;
   next works as above on the blocks that are completely resident and returns LBZ_NEED_INPUT (no
   output) when the next block or the framing behind it has not arrived yet.  While eof has not been
   announced, running out of input is never an error (the reference's retrieve() answers MORE in the
   same situation, src/decode.c:387-396); afterwards the end-of-file rules of lbz_decoder_open apply.
   info->end_bit stays an absolute position in the file. */
int lbz_decoder_open_stream(lbz_decoder *d, unsigned flags);
int lbz_decoder_feed(lbz_decoder *d, const uint8_t *in, size_t n, int eof, size_t *taken);

/* Building blocks for sharding the blocks of ONE file over several decoders / GPUs (blocks are
   independent once the scanner has found their start bits; lbzip2_b200/sharding.py
   sharded_decompress): decode a share of the candidates (count <= max_blocks; table[i]
   describes candidate i), run the framing walk on the merged table of all shares (pure host
   code; sorted by pos; returns LBZ_OK, LBZ_MORE if a block is missing from the table, or the
   error that ends the stream, to be reported after the CRCs of chain[] have been checked), then
   write the confirmed blocks of the share at out_off[i] (~0 = skip) and collect their CRCs. */
int lbz_decoder_decode_at(lbz_decoder *d, const uint8_t *in, size_t n, const uint64_t *magic_bits, uint32_t count,
                          lbz_dblock *table, unsigned flags);
int lbz_walk_table(const uint8_t *in, size_t n, const lbz_dblock *table, size_t count, uint32_t *chain,
                   uint32_t *chain_crc, size_t *nchain, lbz_dstream_info *info);
int lbz_decoder_emit_at(lbz_decoder *d, const uint64_t *out_off, uint32_t count, uint8_t *out, size_t out_cap,
                        size_t *out_len, uint32_t *crc);

/* Block-boundary scanner alone (row f3): bit positions of every 48-bit block
   magic 0x314159265359 in the input, ascending.  Returns the number found
   (written up to cap), negative on failure. */
long lbz_scan_blocks(lbz_decoder *d, const uint8_t *in, size_t n, uint64_t *bit_positions, size_t cap);

/* Diagnostics / parity hooks: arrays of the LAST wave of the last call. */
enum lbz_darray {
  LBZ_DA_BLOCK = 0,    /* struct lbz_dblock of a slot                          */
  LBZ_DA_BWT = 1,      /* u8  last column recovered by the prefix decoder      */
  LBZ_DA_TEXT = 2,     /* u8  inverse BWT output (still run-length coded)      */
  LBZ_DA_OUT = 3       /* u8  the wave's decoded bytes (slot = byte offset)    */
};

int lbz_decoder_read(lbz_decoder *d, int array, uint64_t slot, void *dst, size_t bytes);
uint32_t lbz_decoder_last_wave_blocks(const lbz_decoder *d);
uint64_t lbz_decoder_launches(const lbz_decoder *d);
size_t lbz_decoder_device_bytes(const lbz_decoder *d);
/* Device time of the last call (CUDA events on the decoder's stream): whole call and
   {upload, scan, prefix decode + inverse MTF, successor table, splitter walks, run
   expansion + CRC, tail}; stages are those of the last wave. */
double lbz_decoder_last_ms(const lbz_decoder *d);
void lbz_decoder_stage_ms(const lbz_decoder *d, double *out7);

#ifdef __cplusplus
}
#endif
#endif /* LBZIP2_B200_H */
