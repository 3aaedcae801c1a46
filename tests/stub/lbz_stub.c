/*
 * lbz_stub.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU stand-in for the batch entry points of include/lbzip2_b200.h, backed by
 * the oracle (oracle/bz_oracle.c).  It exists so that the host-side task graph
 * (lbzip2_b200/host/compress_b200.c: staging, dynamic batching, reordering,
 * CRC folding) can be exercised by the `-m "not gpu"` tests in a container
 * without a GPU.  It is linked only into oracle/_ref/lbzip2_b200_hosttest
 * (oracle/Makefile); the product library never contains it.
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/lbzip2_b200.h"
#include "../../oracle/bz_oracle.h"

struct lbz_engine { int level; int max_chunks; };

lbz_engine *lbz_engine_create(int device, int level, int max_chunks) {
  lbz_engine *e = malloc(sizeof *e);
  (void)device;
  e->level = level; e->max_chunks = max_chunks;
  return e;
}
void lbz_engine_destroy(lbz_engine *e) { free(e); }
size_t lbz_bound(size_t n) { return n + n / 32 + 8192 * (n / 100000 + 2) + 64; }
void *lbz_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void lbz_host_free(void *p) { free(p); }

int lbz_compress_chunks(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                        size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs) {
  const size_t mbs = (size_t)e->level * 100000u;
  size_t o = 0, k = 0, pos;
  if (n > (size_t)e->max_chunks * mbs) return -1;
  for (pos = 0; pos < n; pos += mbs) {
    size_t len = n - pos < mbs ? n - pos : mbs, done = 0;
    while (done < len) {                      /* one or two blocks per chunk (src/compress.c:93-110) */
      struct orc_block_info bi;
      size_t w;
      if (out_cap - o < (len - done) + (len - done) / 64 + 2048) return -2;
      w = orc_encode_block(in + pos + done, len - done, (uint32_t)mbs, out + o, &bi);
      if (k < max_recs) {
        memset(&recs[k], 0, sizeof recs[k]);
        recs[k].raw_offset = pos + done; recs[k].raw_len = (uint32_t)bi.consumed;
        recs[k].nblock = bi.nblock; recs[k].crc = bi.block_crc; recs[k].out_len = (uint32_t)w;
      }
      k++; o += w; done += (size_t)bi.consumed;
    }
  }
  *out_len = o; *num_recs = k;
  return 0;
}
