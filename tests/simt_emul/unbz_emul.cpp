// unbz_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the decompressor's device code (lbzip2_b200/csrc/unbz_kernels.cuh) and host
// orchestration (unbz_engine.inc) for the HOST, running every "launch" as a sequential loop over
// thread indices.  The kernels of that file are written without barriers or warp intrinsics, so a
// sequential schedule is one of the schedules a GPU may produce; this lets the decode logic be
// checked against the oracle in the GPU-less development container (tests/test_unbz_emul.py).
// Nothing in lbzip2_b200/ loads this library; the product has no host path.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define UB_EMUL 1
#include "../../include/lbzip2_b200.h"
#include "../../lbzip2_b200/csrc/unbz_kernels.cuh"

UbEmuIdx ub_emu_idx;

#define UB_NSTAGE 7
struct UbBackend { int unused; };
struct lbz_decoder;

static int ub_dev_alloc(lbz_decoder *, void **p, size_t bytes) { *p = calloc(1, bytes ? bytes : 16); return *p ? 0 : -1; }
static void ub_dev_free(lbz_decoder *, void *p) { free(p); }
static int ub_h2d(lbz_decoder *, void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
static int ub_d2h(lbz_decoder *, void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
static int ub_dev_zero(lbz_decoder *, void *p, size_t bytes) { memset(p, 0, bytes); return 0; }
static int ub_dev_fill32(lbz_decoder *, uint32_t *p, uint32_t v, size_t count) { memset(p, (int)(v & 0xFFu), count * 4); return 0; }
static int ub_dev_move(lbz_decoder *, void *dst, const void *src, size_t bytes) { memmove(dst, src, bytes); return 0; }
static int ub_sync(lbz_decoder *) { return 0; }
static void ub_mark(lbz_decoder *, int) {}
static void ub_timers_collect(lbz_decoder *) {}
static inline void ub_count_launch(lbz_decoder *d);

// order in which the emulated threads run: 0 = ascending, 1 = descending (a different legal schedule)
static int ub_emu_reverse = 0;

#define UB_LAUNCH(d, kern, nthreads, cta, ...)                                                      \
  do {                                                                                              \
    uint64_t nt_ = (nthreads);                                                                      \
    uint64_t grid_ = (nt_ + (cta) - 1) / (cta);                                                     \
    for (uint64_t q_ = 0; q_ < grid_ * (cta); q_++) {                                               \
      uint64_t g_ = ub_emu_reverse ? grid_ * (cta) - 1 - q_ : q_;                                   \
      ub_emu_idx.gid = (unsigned)g_; ub_emu_idx.bid = (unsigned)(g_ / (cta)); ub_emu_idx.tid = (unsigned)(g_ % (cta)); \
      kern(__VA_ARGS__);                                                                            \
    }                                                                                               \
    if (nt_) ub_count_launch(d);                                                                    \
  } while (0)

#define UB_LAUNCH_SMEM(d, kern, nthreads, cta, smem, ...) UB_LAUNCH(d, kern, nthreads, cta, __VA_ARGS__)
static size_t ub_chain_smem() { return 0; }
#define UB_CHAIN_LAUNCH(d, nwords, nblk)                                                             \
  UB_LAUNCH(d, k_ub_chain, (uint64_t)(nblk) * 32u, 32u, (d)->d_words, (nwords), (d)->d_blk, (nblk), (d)->d_sel, (d)->d_tree, \
            (d)->d_l1, (d)->d_ml, (d)->d_mq, (d)->d_gpos, (d)->d_gtree)

#include "../../lbzip2_b200/csrc/unbz_engine.inc"

static inline void ub_count_launch(lbz_decoder *d) { d->launches++; }

extern "C" {

void emu_set_reverse(int r) { ub_emu_reverse = r; }

lbz_decoder *emu_decoder_create(int max_blocks, size_t in_cap, size_t out_cap) {
  lbz_decoder *d = new lbz_decoder();
  const size_t min_out = (size_t)UB_MAXBLK / 5 * 259 + 4096;
  d->max_blocks = (uint32_t)(max_blocks < 1 ? 1 : max_blocks);
  d->in_cap = in_cap < 64 ? 64 : in_cap;
  d->out_cap = out_cap < min_out ? min_out : out_cap;
  d->launches = 0;
  d->loaded_n = ~0ull;
  d->h_blk = (UbBlock *)calloc(d->max_blocks, sizeof(UbBlock));
  if (ub_decoder_alloc(d) != 0) return nullptr;
  return d;
}
void emu_decoder_destroy(lbz_decoder *d) { if (!d) return; ub_decoder_release(d); free(d->h_blk); delete d; }
int emu_decompress(lbz_decoder *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len,
                   lbz_dstream_info *info, unsigned flags) {
  return ub_decompress(d, in, n, out, out_cap, out_len, info, flags);
}
int emu_decoder_open(lbz_decoder *d, const uint8_t *in, size_t n, unsigned flags) { return ub_open(d, in, n, flags); }
int emu_decoder_next(lbz_decoder *d, uint8_t *out, size_t out_cap, size_t *out_len, lbz_dstream_info *info) {
  return ub_next(d, out, out_cap, out_len, info);
}
int emu_decoder_open_stream(lbz_decoder *d, unsigned flags) { return ub_open_stream(d, flags); }
int emu_decoder_feed(lbz_decoder *d, const uint8_t *in, size_t n, int eof, size_t *taken) { return ub_feed(d, in, n, eof, taken); }
int emu_decode_at(lbz_decoder *d, const uint8_t *in, size_t n, const uint64_t *pos, uint32_t count, lbz_dblock *table, unsigned flags) {
  return ub_decode_at(d, in, n, pos, count, table, flags);
}
int emu_emit_at(lbz_decoder *d, const uint64_t *out_off, uint32_t count, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *crc) {
  return ub_emit_at(d, out_off, count, out, out_cap, out_len, crc);
}
int emu_walk_table(const uint8_t *in, size_t n, const lbz_dblock *table, size_t count, uint32_t *chain, uint32_t *chain_crc,
                   size_t *nchain, lbz_dstream_info *info) {
  return ub_walk_table(in, n, table, count, chain, chain_crc, nchain, info);
}
long emu_scan_blocks(lbz_decoder *d, const uint8_t *in, size_t n, uint64_t *pos, size_t cap) {
  if (ub_upload(d, in, n) != 0 || ub_scan(d, (n + 3) / 4) != 0) return -1;
  for (size_t i = 0; i < d->hits.size() && i < cap; i++) pos[i] = d->hits[i];
  return (long)d->hits.size();
}
int emu_decoder_read(lbz_decoder *d, int array, uint64_t slot, void *dst, size_t bytes) {
  const void *src;
  switch (array) {
    case LBZ_DA_BLOCK: src = d->d_blk + slot; break;
    case LBZ_DA_BWT: src = d->d_bwt + slot * UB_STRIDE; break;
    case LBZ_DA_TEXT: src = d->d_txt + slot * UB_STRIDE; break;
    case LBZ_DA_OUT: src = d->d_out + slot; break;
    default: return -1;
  }
  memcpy(dst, src, bytes);
  return 0;
}
uint32_t emu_last_wave_blocks(const lbz_decoder *d) { return d->last_wave_blocks; }
uint64_t emu_launches(const lbz_decoder *d) { return d->launches; }
uint32_t emu_mtf_front(uint32_t *lw, uint32_t r) { return ub_mtf_front(lw, r); }
uint32_t emu_gf_shift(uint32_t a, uint64_t nbytes, const uint32_t *pw) { return ub_gf_shift(a, nbytes, pw); }
}

#ifdef UB_EMUL_ABI
// The product's decoder entry points over the emulation, for the CPU-only check of the expansion
// task graph (oracle/Makefile target _ref/lbzip2_b200_hosttest).  Never part of libbz2b200.so.
extern "C" {
lbz_decoder *lbz_decoder_create(int, int max_blocks, size_t in_cap, size_t out_cap) { return emu_decoder_create(max_blocks, in_cap, out_cap); }
void lbz_decoder_destroy(lbz_decoder *d) { emu_decoder_destroy(d); }
int lbz_decoder_open(lbz_decoder *d, const uint8_t *in, size_t n, unsigned flags) { return ub_open(d, in, n, flags); }
int lbz_decoder_next(lbz_decoder *d, uint8_t *out, size_t out_cap, size_t *out_len, lbz_dstream_info *info) {
  return ub_next(d, out, out_cap, out_len, info);
}
int lbz_decoder_open_stream(lbz_decoder *d, unsigned flags) { return ub_open_stream(d, flags); }
int lbz_decoder_feed(lbz_decoder *d, const uint8_t *in, size_t n, int eof, size_t *taken) {
  size_t dummy = 0;
  return ub_feed(d, in, n, eof, taken ? taken : &dummy);
}
const char *lbz_strerror(int status) {
  static const char *const text[] = {
    "not a valid bzip2 file", "bad block header magic", "empty source alphabet", "bad number of trees",
    "no coding groups", "invalid selector", "invalid delta code", "invalid prefix code",
    "incomplete prefix code", "empty block", "unterminated block", "missing run length",
    "block CRC mismatch", "stream CRC mismatch", "block overflow", "primary index too large",
    "unexpected end of file"};
  if (status == LBZ_OK) return "ok";
  if (status >= LBZ_ERR_MAGIC && status <= LBZ_ERR_EOF) return text[status - LBZ_ERR_MAGIC];
  return "internal error";
}
}
#endif
