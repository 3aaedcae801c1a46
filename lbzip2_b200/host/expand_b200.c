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

     setup    the first free worker creates the CUDA context, the decoder and the
              page-locked buffers while the reader is already reading;
     stage    a worker appends one input buffer to the accumulation buffer and gives
              it back to the reader at once;
     decode   one worker at a time hands the accumulated bytes to the decoder's
              window (lbz_decoder_feed), pushes the next wave of completely resident
              blocks through the kernels (lbz_decoder_next) and queues the decoded
              bytes; it runs when the decoder still holds undecoded blocks, when the
              accumulation buffer is full, or when the input has ended -- never for
              a few fresh bytes, a wave has a fixed cost.  Data errors are reported
              with the reference's texts (src/expand.c:70-94) through failf(), after
              the bytes in front of the bad block have been written;
     write    decoded waves go to the writer thread in order.

  Memory is bounded by the window, not by the file: the decoder keeps at most
  LBZIP2_B200_DWINDOW_MB of the compressed file resident (device + host mirror for
  the framing walk) and drops what its waves have consumed; two accumulation
  buffers of half that size let the reader run ahead while a wave is decoded.  The
  decoded data is streamed, one wave (LBZIP2_B200_DWAVE_MB, default 256 MB) at a
  time, through a small pool of buffers that the writer hands back (a fresh buffer
  per wave cost one page fault per 4 KiB of output).  The buffers are ordinary
  memory by default: page-locking them (LBZIP2_B200_PINNED=1, lbz_host_alloc) makes
  the copies faster but was measured at 1.8 s per GB locked on the B200 boxes
  (profiles/r02_run21_rank_ab_streaming_cli.log), which only files of tens of GB
  earn back.  Files of any size and pipes work the same way (the reference's
  expand.c has the same property through its 256 KiB buffers).

  Environment: LBZIP2_B200_DBLOCKS   candidate blocks per wave   (default 320)
               LBZIP2_B200_DWAVE_MB  decoded bytes per wave      (default 256)
               LBZIP2_B200_DWINDOW_MB  compressed bytes resident (default 256; less for smaller files)
               LBZIP2_B200_PINNED    1: page-locked staging buffers (for very large files)
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
/* Compressed bytes on their way to the decoder: the stage task fills acc[fill_i], the decode task
   drains `feeding' (the other one) into the decoder's window.  Both live across operands. */
static uint8_t *acc[2];
static size_t acc_alloc;        /* bytes each of them holds */
static size_t acc_cap;          /* ... and how many this operand uses */
static unsigned fill_i;
static size_t acc_len;          /* bytes in acc[fill_i] */
static uint8_t *feeding;
static size_t feed_len, feed_off;
static size_t total_in;         /* compressed bytes staged so far, first header included */
static uint64_t next_id;        /* next input sequence number */
static uint64_t next_stage;     /* next sequence number to be appended */
static bool staging;            /* a worker is appending */
static lbz_decoder *dec;        /* lives across operands (main.c:935) */
static size_t dec_in_cap, dec_wave_cap;
static unsigned dec_blocks;
static bool session_open, decoding, decode_done;
static bool setup_started, setup_done;
static bool more_ready;         /* the decoder holds complete blocks that no wave has taken yet */
static bool eof_fed;            /* the decoder knows that it has seen the last byte */
#define POOL_MAX (MAX_PENDING + 2u)
static void *pool[POOL_MAX];    /* idle wave buffers (page-locked); live across operands like `dec' */
static unsigned pool_n;
static uint64_t next_wave, write_wave;
static unsigned pending;
static size_t weight_done;
static bool stats;
static bool pinned;             /* staging buffers are page-locked (LBZIP2_B200_PINNED=1) */
static double stat_t0, stat_gpu, stat_setup;
static unsigned long stat_waves, stat_blocks, stat_candidates, stat_false, stat_starved;


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


static void *
stage_alloc(size_t n)
{
  void *p;

  if (!pinned)
    return xmalloc(n);
  p = lbz_host_alloc(n);
  if (p == NULL)
    failx(0, "cannot allocate a page-locked buffer of %zu bytes", n);
  return p;
}

static void
stage_free(void *p)
{
  if (pinned)
    lbz_host_free(p);
  else
    free(p);
}


/* The window for this operand: main.c knows the size of a regular file (ispec.size, 0 = unknown:
   a pipe); small files get small buffers. */
static size_t
window_bytes(void)
{
  size_t w = env_size("LBZIP2_B200_DWINDOW_MB", 256, 2, 65536) << 20;

  if (ispec.size > 0 && (uintmax_t)w > ispec.size + 65536u)
    w = (size_t)ispec.size + 65536u;
  if (w < ((size_t)4 << 20))
    w = (size_t)4 << 20;        /* room for the largest block (1.1 MB) and the largest input buffer */
  return w;
}

static bool
can_setup(void)
{
  return !setup_started;
}

static void
do_setup(void)
{
  size_t want = window_bytes();
  double t0;

  setup_started = true;
  sched_unlock();

  t0 = now();
  if (dec == NULL || dec_in_cap < want) {
    if (dec != NULL)
      lbz_decoder_destroy(dec);
    dec_in_cap = want;
    dec = lbz_decoder_create((int)env_size("LBZIP2_B200_DEVICE", 0, 0, 63),
                             (int)dec_blocks, dec_in_cap, dec_wave_cap);
    if (dec == NULL)
      failx(0, "cannot set up the GPU decoder");
  }
  if (acc_alloc < want / 2) {
    unsigned i;

    for (i = 0; i < 2u; i++) {
      if (acc[i] != NULL)
        stage_free(acc[i]);
      acc[i] = stage_alloc(want / 2);
    }
    acc_alloc = want / 2;
  }
  stat_setup += now() - t0;

  /* The first four bytes were consumed by work() (src/process.c:664-672),
     which also set bs100k from them; the decoder wants the whole file. */
  acc[0][0] = 0x42;
  acc[0][1] = 0x5A;
  acc[0][2] = 0x68;
  acc[0][3] = 0x30 + bs100k;

  sched_lock();
  acc_cap = want / 2;
  fill_i = 0;
  acc_len = 4;
  total_in = 4;
  setup_done = true;
}


static bool
head_is_next(void)
{
  return !empty(stage_q) && peek(stage_q)->pos.major == next_stage;
}

static bool
can_stage(void)
{
  if (!setup_done || staging || !head_is_next())
    return false;
  /* once the decoder has reached its verdict (the end of the last stream with garbage behind it, or
     a data error) the rest of the input is read and dropped, as the reference does (src/expand.c:428-436) */
  return decode_done || acc_len + peek(stage_q)->size <= acc_cap;
}

static void
do_stage(void)
{
  struct in_blk *iblk = dequeue(stage_q);
  uint8_t *dst = acc[fill_i] + acc_len;
  bool drop = decode_done;

  if (!drop && iblk->size > acc_cap)
    failx(0, "an input buffer of %zu bytes does not fit the staging buffer (%zu)", iblk->size, acc_cap);
  next_stage++;
  staging = true;               /* appends are ordered: one at a time; no buffer swap meanwhile */
  sched_unlock();

  if (!drop)
    memcpy(dst, iblk->buffer, iblk->size);
  source_release_buffer(iblk->buffer);

  sched_lock();
  if (!drop)
    acc_len += iblk->size;
  total_in += iblk->size;
  staging = false;
  free(iblk);
}


static int deferred_err = LBZ_OK;   /* data error found by the last wave, raised after its valid bytes are written */

static bool
input_complete(void)
{
  return eof && empty(stage_q) && !staging;
}

/* A wave has a fixed cost of some ten milliseconds: it is started for blocks the decoder already
   holds, for a full accumulation buffer (the reader is then about to stall), or at the end of the
   input -- not for every fresh buffer. */
static bool
can_decode(void)
{
  if (!setup_done || decoding || decode_done || staging || pending >= MAX_PENDING)
    return false;
  if (more_ready || feed_off < feed_len || input_complete())
    return true;
  return head_is_next() && acc_len + peek(stage_q)->size > acc_cap;
}

static void
do_decode(void)
{
  struct out_blk *oblk;
  lbz_dstream_info inf;
  size_t got = 0;
  uint8_t *buf;
  double t0;
  bool last;
  int rv;

  decoding = true;
  if (feed_off == feed_len && acc_len > 0) {
    /* the stage task goes on with the other buffer */
    feeding = acc[fill_i];
    feed_len = acc_len;
    feed_off = 0;
    fill_i ^= 1u;
    acc_len = 0;
  }
  last = input_complete() && acc_len == 0;   /* `feeding' holds everything that is left */
  buf = pool_n > 0 ? pool[--pool_n] : NULL;
  sched_unlock();

  if (!session_open) {
    rv = lbz_decoder_open_stream(dec, 0);
    if (rv != LBZ_OK)
      failx(0, "GPU decompression failed (lbz_decoder_open_stream returned %d)", rv);
    session_open = true;
    eof_fed = false;
  }
  if (!eof_fed && (feed_off < feed_len || last)) {
    size_t taken = 0;

    rv = lbz_decoder_feed(dec, feeding + feed_off, feed_len - feed_off, last, &taken);
    if (rv != 0)
      failx(0, "GPU decompression failed (lbz_decoder_feed returned %d)", rv);
    feed_off += taken;
    if (last && feed_off == feed_len)
      eof_fed = true;
  }

  if (buf == NULL) {
    t0 = now();
    buf = stage_alloc(dec_wave_cap);
    stat_setup += now() - t0;
  }
  t0 = now();
  rv = lbz_decoder_next(dec, buf, dec_wave_cap, &got, &inf);
  stat_gpu += now() - t0;
  if (rv < 0 || rv == LBZ_ERR_OUTCAP)
    failx(0, "GPU decompression failed (lbz_decoder_next returned %d)", rv);
  if (rv == LBZ_NEED_INPUT) {
    /* nothing decodable is resident: wait for the stage task (or for the end of the input) */
    if (eof_fed)
      failx(0, "GPU decompression failed (the decoder wants input after the end of the file)");
    sched_lock();
    if (pool_n < POOL_MAX)
      pool[pool_n++] = buf;
    more_ready = false;
    stat_starved++;
    decoding = false;
    return;
  }
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
    if (rv == LBZ_OK || upto > total_in)
      upto = total_in;
    oblk->weight = upto > weight_done ? upto - weight_done : 0;
    weight_done += oblk->weight;
  }
  stat_waves++;
  more_ready = rv == LBZ_MORE;
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
  sched_lock();
  if (pool_n < POOL_MAX) {
    pool[pool_n++] = buffer;
    buffer = NULL;
  }
  ++out_slots;
  sched_unlock();
  if (buffer != NULL)
    stage_free(buffer);
}


static void
init(void)
{
  stat_t0 = now();
  stats = getenv("LBZIP2_B200_STATS") != NULL;
  if (dec == NULL)              /* fixed for the process: the buffers outlive an operand */
    pinned = env_size("LBZIP2_B200_PINNED", 0, 0, 1) != 0;
  stat_gpu = stat_setup = 0.0;
  stat_waves = stat_blocks = stat_candidates = stat_false = stat_starved = 0;
  dec_blocks = (unsigned)env_size("LBZIP2_B200_DBLOCKS", 320, 1, 16384);
  dec_wave_cap = env_size("LBZIP2_B200_DWAVE_MB", 256, 48, 65536) << 20;

  pqueue_init(stage_q, total_in_slots);
  pqueue_init(write_q, MAX_PENDING);

  assert(1 <= bs100k && bs100k <= 9);
  acc_len = 0;
  acc_cap = 0;
  feeding = NULL;
  feed_len = feed_off = 0;
  total_in = 0;
  next_id = 0;
  next_stage = 0;
  staging = false;
  session_open = decoding = decode_done = false;
  setup_started = setup_done = false;
  more_ready = eof_fed = false;
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
    fprintf(stderr, "lbzip2_b200: %zu compressed bytes through a window of %zu, %lu blocks in %lu "
            "waves (%lu scanner candidates, %lu rejected; the decoder waited for input %lu times); "
            "set-up %.3f s, in lbz_decoder_next %.3f s, wall %.3f s\n", total_in, dec_in_cap,
            stat_blocks, stat_waves, stat_candidates, stat_false, stat_starved, stat_setup,
            stat_gpu, now() - stat_t0);
    fflush(stderr);             /* main.c:912-916 makes stderr fully buffered */
  }
  pqueue_uninit(stage_q);
  pqueue_uninit(write_q);
}


static const struct task task_list[] = {
  { "setup",  can_setup,  do_setup  },
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
