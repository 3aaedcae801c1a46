/*
  expand_b200.c -- batch-aware DEcompression task graph for lbzip2's scheduler.

  Host side of the B200 decoder in the reference's own language (C99).  It is
  compiled INSTEAD of the reference's src/expand.c, src/decode.c and
  src/parse.c, together with the reference's unmodified main.c / process.c /
  signals.c / ... and linked against libbz2b200.so (recipe: oracle/Makefile,
  target _ref/lbzip2_b200).  It provides the one symbol the scheduler needs,
  `const struct process expansion' (reference src/process.h:34-41,93), and
  drives the GPU through the wave interface of the C ABI
  (include/lbzip2_b200.h section 4: lbz_decoder_open / lbz_decoder_next).

  What it replaces, and why: the
  reference task graph (src/expand.c:916-923) lets worker threads scan() for
  block magics, retrieve() one block each out of 256 KiB input buffers with a
  resumable bit-serial automaton, decode() and emit() it, while parse() walks
  the framing and throws away mis-recognised candidates.  On the device the
  same roles exist, but for a whole wave of candidate blocks at a time
  (k_ub_scan; k_ub_header .. k_ub_symbols; k_ub_lf_* / k_ub_walk*; k_ub_rl_*;
  the framing walk in csrc/unbz_engine.inc), so the task graph here only moves
  bytes:

     stage    a worker appends one input buffer to the compressed image of the
              file and gives the buffer back to the reader at once;
     decode   once the image is complete, one worker pushes the next wave of
              blocks through the kernels (lbz_decoder_next) and queues the
              decoded bytes; data errors are reported with the reference's
              texts (src/expand.c:70-94) through failf();
     write    decoded waves go to the writer thread in order.

  The whole compressed file is held in host memory (the scanner and the
  prefix decoder want to see every block of a wave at once); the decoded data
  is streamed, one wave (LBZIP2_B200_DWAVE_MB, default 256 MB) at a time.

  Environment: LBZIP2_B200_DBLOCKS   candidate blocks per wave   (default 320)
               LBZIP2_B200_DWAVE_MB  decoded bytes per wave      (default 256)
               LBZIP2_B200_DEVICE    device ordinal              (default 0)
               LBZIP2_B200_STATS     print statistics to stderr at the end
*/
#include "common.h"

#include <string.h>             /* memcpy() */
#include <stdio.h>              /* fprintf() */
#include <time.h>               /* clock_gettime() */

#include "main.h"               /* bs100k, ispec, xmalloc(), failf() */
#include "process.h"            /* struct process, queues */

#include "lbzip2_b200.h"        /* lbz_decoder, lbz_decoder_open/next() */


/* An input buffer waiting to be appended.  `pos' must stay first: the heap
   helpers of process.c order elements by it (src/process.c:165-218). */
struct in_blk {
  struct position pos;          /* major = sequence number of the buffer */
  void *buffer;
  size_t size;
};

/* A decoded wave waiting for its turn at the writer. */
struct out_blk {
  struct position pos;          /* major = wave number */
  void *buffer;
  size_t size;
  size_t weight;                /* compressed bytes it accounts for */
};

#define MAX_PENDING 4u          /* decoded waves not yet handed to the writer */

static struct pqueue(struct in_blk *) stage_q;
static struct pqueue(struct out_blk *) write_q;
static uint8_t *image;          /* the compressed file, first header included */
static size_t image_len, image_cap;
static uint64_t next_id;        /* next input sequence number */
static uint64_t next_stage;     /* next sequence number to be appended */
static bool staging;            /* a worker is appending */
static lbz_decoder *dec;        /* lives across operands (main.c:935) */
static size_t dec_in_cap, dec_wave_cap;
static unsigned dec_blocks;
static bool session_open, decoding, decode_done;
static uint64_t next_wave, write_wave;
static unsigned pending;
static size_t weight_done;
static bool stats;
static double stat_t0, stat_gpu, stat_setup;
static unsigned long stat_waves, stat_blocks, stat_candidates, stat_false;


static double
now(void)
{
  struct timespec ts;

  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}


static size_t
env_size(const char *name, size_t dflt, size_t lo, size_t hi)
{
  const char *s = getenv(name);
  unsigned long long v;

  if (s == NULL || *s == '\0')
    return dflt;
  v = strtoull(s, NULL, 10);
  if (v < lo)
    v = lo;
  if (v > hi)
    v = hi;
  return (size_t)v;
}


static bool
can_stage(void)
{
  return !staging && !empty(stage_q) && peek(stage_q)->pos.major == next_stage;
}

static void
do_stage(void)
{
  struct in_blk *iblk = dequeue(stage_q);

  next_stage++;
  staging = true;               /* appends are ordered: one at a time */
  sched_unlock();

  if (image_len + iblk->size > image_cap) {
    size_t cap = image_cap;
    uint8_t *bigger;

    while (cap < image_len + iblk->size)
      cap *= 2;
    bigger = xmalloc(cap);
    memcpy(bigger, image, image_len);
    free(image);
    image = bigger;
    image_cap = cap;
  }
  memcpy(image + image_len, iblk->buffer, iblk->size);
  image_len += iblk->size;
  source_release_buffer(iblk->buffer);
  free(iblk);

  sched_lock();
  staging = false;
}


static int deferred_err = LBZ_OK;   /* data error found by the last wave, raised after its valid bytes are written */

static bool
input_complete(void)
{
  return eof && empty(stage_q) && !staging;
}

static bool
can_decode(void)
{
  return input_complete() && !decoding && !decode_done && pending < MAX_PENDING;
}

static void
do_decode(void)
{
  struct out_blk *oblk;
  lbz_dstream_info inf;
  size_t got = 0;
  uint8_t *buf;
  double t0;
  int rv;

  decoding = true;
  sched_unlock();

  if (!session_open) {
    /* the decoder is sized for the largest file seen so far */
    if (dec == NULL || dec_in_cap < image_len) {
      t0 = now();
      if (dec != NULL)
        lbz_decoder_destroy(dec);
      dec_in_cap = image_len + image_len / 4 + 65536;
      dec = lbz_decoder_create((int)env_size("LBZIP2_B200_DEVICE", 0, 0, 63),
                               (int)dec_blocks, dec_in_cap, dec_wave_cap);
      if (dec == NULL)
        failx(0, "cannot set up the GPU decoder");
      stat_setup += now() - t0;
    }
    rv = lbz_decoder_open(dec, image, image_len, 0);
    if (rv < 0)
      failx(0, "GPU decompression failed (lbz_decoder_open returned %d)", rv);
    if (rv != LBZ_OK)           /* cannot happen: main.c checked the header */
      failf(&ispec, "%s", lbz_strerror(rv));
    session_open = true;
  }

  buf = xmalloc(dec_wave_cap);
  t0 = now();
  rv = lbz_decoder_next(dec, buf, dec_wave_cap, &got, &inf);
  stat_gpu += now() - t0;
  if (rv < 0 || rv == LBZ_ERR_OUTCAP)
    failx(0, "GPU decompression failed (lbz_decoder_next returned %d)", rv);
  if (rv != LBZ_OK && rv != LBZ_MORE) {
    /* a data error: the `got` bytes in front of the bad block are valid (lbz_decoder_next
       guarantees it).  Like the reference, which has handed the blocks before the bad one to
       the writer when it fails (src/expand.c:703-741), they are written first; the error is
       raised once the writer has drained (uninit). */
    deferred_err = rv;
    rv = LBZ_OK;
  }

  oblk = XMALLOC(struct out_blk);
  oblk->buffer = buf;
  oblk->size = got;

  sched_lock();
  oblk->pos.major = next_wave++;
  oblk->pos.minor = 0;
  /* progress is accounted in compressed bytes, like src/expand.c:721-723 */
  {
    size_t upto = (size_t)(inf.end_bit / 8);
    if (rv == LBZ_OK || upto > image_len)
      upto = image_len;
    oblk->weight = upto > weight_done ? upto - weight_done : 0;
    weight_done += oblk->weight;
  }
  stat_waves++;
  if (rv == LBZ_OK) {
    decode_done = true;
    session_open = false;
    stat_blocks += inf.num_blocks;
    stat_candidates += inf.candidates;
    stat_false += inf.false_candidates;
  }
  pending++;
  enqueue(write_q, oblk);
  decoding = false;
}


static bool
can_write(void)
{
  return !empty(write_q) && peek(write_q)->pos.major == write_wave &&
    out_slots > 0;
}

static void
do_write(void)
{
  struct out_blk *oblk = dequeue(write_q);

  write_wave++;
  --out_slots;
  --pending;
  sink_write_buffer(oblk->buffer, oblk->size, oblk->weight);
  free(oblk);
}


static bool
can_terminate(void)
{
  return input_complete() && decode_done && !decoding && empty(write_q) &&
    out_slots == total_out_slots;
}


static void
on_input_avail(void *buffer, size_t size)
{
  struct in_blk *iblk = XMALLOC(struct in_blk);

  iblk->pos.major = next_id++;
  iblk->pos.minor = 0u;
  iblk->buffer = buffer;
  iblk->size = size;

  sched_lock();
  enqueue(stage_q, iblk);
  sched_unlock();
}


static void
on_write_complete(void *buffer)
{
  free(buffer);

  sched_lock();
  ++out_slots;
  sched_unlock();
}


static void
init(void)
{
  stat_t0 = now();
  stats = getenv("LBZIP2_B200_STATS") != NULL;
  stat_gpu = stat_setup = 0.0;
  stat_waves = stat_blocks = stat_candidates = stat_false = 0;
  dec_blocks = (unsigned)env_size("LBZIP2_B200_DBLOCKS", 320, 1, 16384);
  dec_wave_cap = env_size("LBZIP2_B200_DWAVE_MB", 256, 48, 65536) << 20;

  pqueue_init(stage_q, total_in_slots);
  pqueue_init(write_q, MAX_PENDING);

  /* The first four bytes were consumed by work() (src/process.c:664-672),
     which also set bs100k from them; the decoder wants the whole file. */
  assert(1 <= bs100k && bs100k <= 9);
  image_cap = 1u << 22;
  image = xmalloc(image_cap);
  image[0] = 0x42;
  image[1] = 0x5A;
  image[2] = 0x68;
  image[3] = 0x30 + bs100k;
  image_len = 4;

  next_id = 0;
  next_stage = 0;
  staging = false;
  session_open = decoding = decode_done = false;
  next_wave = write_wave = 0;
  pending = 0;
  weight_done = 4;
}


static void
uninit(void)
{
  if (deferred_err != LBZ_OK) {
    int e = deferred_err;
    deferred_err = LBZ_OK;
    failf(&ispec, "compressed data error: %s", lbz_strerror(e));
  }
  if (stats) {
    fprintf(stderr, "lbzip2_b200: %zu compressed bytes, %lu blocks in %lu "
            "waves (%lu scanner candidates, %lu rejected); decoder set-up "
            "%.3f s, in lbz_decoder_next %.3f s, wall %.3f s\n", image_len,
            stat_blocks, stat_waves, stat_candidates, stat_false, stat_setup,
            stat_gpu, now() - stat_t0);
    fflush(stderr);             /* main.c:912-916 makes stderr fully buffered */
  }
  free(image);
  image = NULL;
  pqueue_uninit(stage_q);
  pqueue_uninit(write_q);
}


static const struct task task_list[] = {
  { "write",  can_write,  do_write  },
  { "decode", can_decode, do_decode },
  { "stage",  can_stage,  do_stage  },
  { NULL,     NULL,       NULL      },
};

const struct process expansion = {
  task_list,
  init,
  uninit,
  can_terminate,
  on_input_avail,
  on_write_complete,
};
