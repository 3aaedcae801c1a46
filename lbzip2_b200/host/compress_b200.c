/*
  compress_b200.c -- batch-aware compression task graph for lbzip2's scheduler.

  Host side of the B200 engine in the reference's own language (C99).  It is
  compiled INSTEAD of the reference's src/compress.c, together with the
  reference's unmodified main.c / process.c / signals.c / ... and linked
  against libbz2b200.so (recipe: oracle/Makefile, target _ref/lbzip2_b200).
  It provides the one symbol the scheduler needs, `const struct process
  compression` (reference src/process.h:34-41,92), and drives the GPU through
  the batch entry point of the C ABI (include/lbzip2_b200.h section 2).

  What it replaces, and why: the reference task
  graph (src/compress.c:353-359) hands ONE block to ONE worker thread
  (do_collect -> collect/encode, do_transmit -> transmit).  One block cannot
  fill a B200, so here the unit of scheduling is a BATCH of consecutive input
  buffers:

     stage    any worker copies one input buffer (in_granul = bs100k*100000
              bytes, src/process.c:631) into its slot of the open batch's pinned
              staging memory and gives the buffer back to the reader at once
              (source_release_buffer) -- many workers stage in parallel;
     launch   one worker pushes a complete batch through the kernels
              (lbz_compress_chunks: RLE1+CRC, BWT, MTF/RLE2, Huffman, bit-pack
              for every chunk of the batch at once) and queues the bytes;
     reorder  batches are handed to the writer thread in stream order and the
              block CRCs are folded into the stream CRC (src/encode.h:38).

  The chunking is the reference's: every input buffer is exactly one chunk of
  bs100k*100000 raw bytes (only the last buffer of a stream may be shorter,
  src/process.c:107-133), a chunk yields one or two blocks
  (src/compress.c:93-110), so the output is bit-identical to `lbzip2` for any -n.

  Batches form dynamically: a batch is launched when it is full, when the
  input is exhausted, or as soon as no other batch is running on the GPU (so a
  small file is not held back and the GPU never idles while input is waiting).

  Environment: LBZIP2_B200_BATCH   chunks per batch        (default 32)
               LBZIP2_B200_ENGINES engines (batches in flight) per device (default 2)
               LBZIP2_B200_GPUS    devices to spread batches over, round-robin (default 1)
               LBZIP2_B200_DEVICE  first device ordinal (default 0)
               LBZIP2_B200_STATS   print batch statistics to stderr at the end

  Not supported: -u (sequential collect, src/compress.c:120-198).
*/
#include "common.h"

#include <assert.h>
#include <string.h>             /* memcpy() */
#include <stdio.h>              /* fprintf() */
#include <time.h>               /* clock_gettime() */

#include "main.h"               /* bs100k, ultra, xmalloc(), failx() */
#include "process.h"            /* struct process, queues */

#include "lbzip2_b200.h"        /* lbz_engine, lbz_compress_chunks() */


/* An input buffer waiting to be staged.  `pos' must stay first: the heap
   helpers of process.c order elements by it (src/process.c:165-218). */
struct in_blk {
  struct position pos;          /* major = sequence number of the buffer */
  void *buffer;
  size_t size;
};

/* A compressed batch waiting for its turn at the writer. */
struct out_blk {
  struct position pos;          /* major = sequence number of its first buffer */
  uint64_t next_seq;            /* sequence number following its last buffer */
  void *buffer;                 /* blocks of the batch, in stream order */
  size_t size;
  size_t weight;                /* raw bytes covered */
  uint32_t *crc;                /* un-inverted block CRCs, in stream order */
  size_t ncrc;
};

enum batch_state { B_NONE, B_CREATING, B_FREE, B_OPEN, B_SEALED, B_RUNNING };

/* One engine = one batch in flight.  Engines are created on demand (task
   `create'): setting up device memory takes longer than compressing a small
   file, so a slot stays B_NONE until input is actually waiting for it, and the
   set-up of engine k+1 overlaps the first batches of engine k. */
struct slot {
  lbz_engine *eng;
  int device;
  uint8_t *h_in;                /* pinned staging: cap chunks */
  uint8_t *h_out;               /* pinned output */
  size_t out_cap;
  lbz_block_rec *recs;
  enum batch_state state;
  uint64_t first_seq;
  unsigned assigned;            /* buffers given a place in h_in */
  unsigned staged;              /* buffers copied */
  size_t in_len;                /* raw bytes assigned */
};

#define MAX_SLOTS 64u
#define MAX_UNSUNK 64u          /* batches opened but not yet handed to the writer */

static struct pqueue(struct in_blk *) stage_q;
static struct pqueue(struct out_blk *) reord_q;
static struct slot slots[MAX_SLOTS];
static unsigned num_slots;
static unsigned batch_cap;      /* chunks per batch */
static size_t chunk_size;       /* bs100k * 100000 */
static int open_slot;           /* slot of the batch being filled, or -1 */
static unsigned running;        /* batches on the GPU */
static unsigned creating;       /* engines being set up */
static bool dev_creating[64];   /* ... one at a time per device */
static unsigned unsunk;
static uint64_t next_id;        /* next input sequence number */
static uint64_t next_stage;     /* next sequence number to be staged */
static uint64_t order;          /* next sequence number the writer expects */
static uint32_t combined_crc;
static bool stats;
static unsigned long stat_batches, stat_chunks, stat_blocks;
static double stat_t0, stat_init, stat_gpu, stat_stage, stat_copy;
static unsigned stat_engines;
static unsigned engines_level;  /* level the engines were created for (0 = none yet) */


static double
now(void)
{
  struct timespec ts;

  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}


static unsigned
env_uint(const char *name, unsigned dflt, unsigned lo, unsigned hi)
{
  const char *s = getenv(name);
  unsigned long v;

  if (s == NULL || *s == '\0')
    return dflt;
  v = strtoul(s, NULL, 10);
  if (v < lo)
    v = lo;
  if (v > hi)
    v = hi;
  return (unsigned)v;
}


static int
find_slot(enum batch_state st)
{
  unsigned i;

  for (i = 0; i < num_slots; i++)
    if (slots[i].state == st)
      return (int)i;
  return -1;
}


/* A sealed batch whose buffers have all been copied, or the open batch when
   nothing more will (or should) be waited for. */
static int
launchable(void)
{
  unsigned i;

  for (i = 0; i < num_slots; i++) {
    const struct slot *s = &slots[i];

    if (s->assigned == 0 || s->staged != s->assigned)
      continue;
    if (s->state == B_SEALED)
      return (int)i;
    if (s->state == B_OPEN && ((eof && empty(stage_q)) || running == 0))
      return (int)i;
  }
  return -1;
}


/* A slot without an engine whose device is not busy setting one up. */
static int
creatable(void)
{
  unsigned i;

  for (i = 0; i < num_slots; i++)
    if (slots[i].state == B_NONE && !dev_creating[slots[i].device])
      return (int)i;
  return -1;
}


/* ---- -u (src/compress.c:120-198: blocks packed across I/O buffers) --------------------------------
   The sequential collector is the reference's own: its task graph (src/compress.c, compiled into this
   binary as `compression_ref`) drives collect()/encode()/transmit() of libbz2b200.so -- collect()
   decides the block split on the calling thread, encode() calls of the worker threads are pooled
   into device batches.  With -u every entry of this process forwards to
   it; the batch tasks below stay idle. */
#ifdef LBZ_WITH_REF_SEQUENTIAL
extern const struct process compression_ref;
#define SEQ (ultra)
#define REF_TASK(i)                                                                      \
  static bool can_ref##i(void) { return ultra && compression_ref.tasks[i].ready(); }      \
  static void do_ref##i(void) { compression_ref.tasks[i].run(); }
REF_TASK(0)
REF_TASK(1)
REF_TASK(2)
REF_TASK(3)
#else
#define SEQ (false)
#endif


static bool
can_create(void)
{
  if (SEQ)
    return false;
  /* input is waiting, no batch is open and no engine is free: set up another one */
  return !empty(stage_q) && open_slot < 0 && find_slot(B_FREE) < 0 &&
    creatable() >= 0;
}


static void
do_create(void)
{
  struct slot *s;
  double t0;

  s = &slots[creatable()];
  s->state = B_CREATING;
  dev_creating[s->device] = true;
  creating++;
  sched_unlock();

  t0 = now();
  s->eng = lbz_engine_create(s->device, (int)bs100k, (int)batch_cap);
  s->out_cap = lbz_bound((size_t)batch_cap * chunk_size);
  s->h_in = lbz_host_alloc((size_t)batch_cap * chunk_size);
  s->h_out = lbz_host_alloc(s->out_cap);
  if (s->eng == NULL || s->h_in == NULL || s->h_out == NULL)
    failx(0, "cannot set up the GPU engine (device %d)", s->device);
  s->recs = XNMALLOC(2u * batch_cap, lbz_block_rec);
  t0 = now() - t0;

  sched_lock();
  stat_init += t0;
  stat_engines++;
  creating--;
  dev_creating[s->device] = false;
  s->state = B_FREE;
}


static bool
can_stage(void)
{
  if (SEQ)
    return false;
  if (empty(stage_q) || peek(stage_q)->pos.major != next_stage)
    return false;
  if (open_slot >= 0)
    return true;                /* an open batch always has room */
  return unsunk < MAX_UNSUNK && find_slot(B_FREE) >= 0;
}


static void
do_stage(void)
{
  struct in_blk *iblk;
  struct slot *s;
  size_t off;
  double t0;

  iblk = dequeue(stage_q);
  next_stage++;

  if (open_slot < 0) {
    open_slot = find_slot(B_FREE);
    s = &slots[open_slot];
    s->state = B_OPEN;
    s->first_seq = iblk->pos.major;
    s->assigned = 0;
    s->staged = 0;
    s->in_len = 0;
    unsunk++;
  }
  s = &slots[open_slot];
  off = (size_t)s->assigned * chunk_size;
  s->assigned++;
  s->in_len += iblk->size;
  /* A full batch is sealed; so is one that received a short buffer, because
     chunk boundaries must stay at multiples of in_granul within a batch. */
  if (s->assigned == batch_cap || iblk->size != chunk_size) {
    s->state = B_SEALED;
    open_slot = -1;
  }
  sched_unlock();

  t0 = stats ? now() : 0.0;
  memcpy(s->h_in + off, iblk->buffer, iblk->size);
  source_release_buffer(iblk->buffer);
  free(iblk);
  t0 = stats ? now() - t0 : 0.0;

  sched_lock();
  stat_stage += t0;
  s->staged++;
}


static bool
can_launch(void)
{
  if (SEQ)
    return false;
  return launchable() >= 0;
}


static void
do_launch(void)
{
  struct slot *s;
  struct out_blk *oblk;
  size_t out_len, nrec, k;
  int i, rc;
  double t0, t1, t2;

  i = launchable();
  s = &slots[i];
  if (open_slot == i)
    open_slot = -1;
  s->state = B_RUNNING;
  running++;
  sched_unlock();

  out_len = 0;
  nrec = 0;
  t0 = stats ? now() : 0.0;
  rc = lbz_compress_chunks(s->eng, s->h_in, s->in_len, s->h_out, s->out_cap,
                           &out_len, s->recs, 2u * batch_cap, &nrec);
  if (rc != 0)
    failx(0, "GPU compression failed (lbz_compress_chunks returned %d)", rc);
  t1 = stats ? now() : 0.0;

  oblk = XMALLOC(struct out_blk);
  oblk->pos.major = s->first_seq;
  oblk->pos.minor = 0;
  oblk->next_seq = s->first_seq + s->assigned;
  oblk->buffer = xmalloc(out_len > 0 ? out_len : 1);
  memcpy(oblk->buffer, s->h_out, out_len);
  oblk->size = out_len;
  oblk->weight = s->in_len;
  oblk->crc = XNMALLOC(nrec > 0 ? nrec : 1, uint32_t);
  oblk->ncrc = nrec;
  for (k = 0; k < nrec; k++)
    oblk->crc[k] = s->recs[k].crc;

  t2 = stats ? now() : 0.0;

  sched_lock();
  if (stats) {
    stat_gpu += t1 - t0;
    stat_copy += t2 - t1;
    stat_batches++;
    stat_chunks += s->assigned;
    stat_blocks += nrec;
  }
  running--;
  s->state = B_FREE;
  enqueue(reord_q, oblk);
}


static bool
can_reorder(void)
{
  if (SEQ)
    return false;
  return !empty(reord_q) && peek(reord_q)->pos.major == order && out_slots > 0;
}


static void
do_reorder(void)
{
  struct out_blk *oblk;
  size_t k;

  oblk = dequeue(reord_q);
  order = oblk->next_seq;
  --out_slots;
  --unsunk;

  for (k = 0; k < oblk->ncrc; k++)
    combined_crc = combine_crc(combined_crc, oblk->crc[k]);
  sink_write_buffer(oblk->buffer, oblk->size, oblk->weight);

  free(oblk->crc);
  free(oblk);
}


static bool
can_terminate(void)
{
#ifdef LBZ_WITH_REF_SEQUENTIAL
  if (ultra)
    return compression_ref.finished();
#endif
  return eof && empty(stage_q) && empty(reord_q) && unsunk == 0 &&
    creating == 0 && out_slots == total_out_slots;
}


static void
on_input_avail(void *buffer, size_t size)
{
  struct in_blk *iblk;

#ifdef LBZ_WITH_REF_SEQUENTIAL
  if (ultra) {
    compression_ref.on_block(buffer, size);
    return;
  }
#endif
  iblk = XMALLOC(struct in_blk);

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
#ifdef LBZ_WITH_REF_SEQUENTIAL
  if (ultra) {
    compression_ref.on_written(buffer);
    return;
  }
#endif
  free(buffer);

  sched_lock();
  ++out_slots;
  sched_unlock();
}


static void
init(void)
{
  uint8_t header[HEADER_SIZE];
  unsigned per_dev, ngpu, dev0, i;

  stat_t0 = now();
  stat_batches = stat_chunks = stat_blocks = 0;
  stat_gpu = stat_stage = stat_copy = 0.0;
#ifdef LBZ_WITH_REF_SEQUENTIAL
  if (ultra) {
    assert(compression_ref.tasks[4].name == NULL);
    compression_ref.init();
    return;
  }
#else
  if (ultra)
    failx(0, "-u is not supported by the GPU task graph");
#endif
  assert(1 <= bs100k && bs100k <= 9);

  batch_cap = env_uint("LBZIP2_B200_BATCH", 32u, 1u, 1024u);
  per_dev = env_uint("LBZIP2_B200_ENGINES", 2u, 1u, 8u);
  ngpu = env_uint("LBZIP2_B200_GPUS", 1u, 1u, 8u);
  dev0 = env_uint("LBZIP2_B200_DEVICE", 0u, 0u, 63u);
  stats = getenv("LBZIP2_B200_STATS") != NULL;
  num_slots = per_dev * ngpu;
  if (num_slots > MAX_SLOTS)
    num_slots = MAX_SLOTS;
  chunk_size = bs100k * 100000u;

  /* The engines outlive one call of work(): main() runs work() once per operand
     (src/main.c:935).  They are never torn down; process exit releases the device. */
  if (engines_level != bs100k) {
    for (i = 0; i < MAX_SLOTS; i++) {
      struct slot *s = &slots[i];

      if (s->eng != NULL) {
        lbz_engine_destroy(s->eng);
        lbz_host_free(s->h_in);
        lbz_host_free(s->h_out);
        free(s->recs);
        s->eng = NULL;
      }
    }
    engines_level = bs100k;
  }
  for (i = 0; i < num_slots; i++) {
    slots[i].device = (int)(dev0 + i % ngpu);
    slots[i].state = slots[i].eng != NULL ? B_FREE : B_NONE;
  }
  creating = 0;
  memset(dev_creating, 0, sizeof(dev_creating));
  stat_init = 0.0;
  stat_engines = 0;

  pqueue_init(stage_q, total_in_slots);
  pqueue_init(reord_q, MAX_UNSUNK);
  open_slot = -1;
  running = 0;
  unsunk = 0;
  next_id = 0;
  next_stage = 0;
  order = 0;
  combined_crc = 0;

  header[0] = 0x42;             /* "BZh" + level, src/compress.c:290-301 */
  header[1] = 0x5A;
  header[2] = 0x68;
  header[3] = 0x30 + bs100k;
  xwrite(header, HEADER_SIZE);
}


static void
uninit(void)
{
#ifdef LBZ_WITH_REF_SEQUENTIAL
  if (ultra) {
    compression_ref.uninit();
    return;
  }
#endif
  uint8_t trailer[TRAILER_SIZE];

  trailer[0] = 0x17;            /* end-of-stream magic + stream CRC, */
  trailer[1] = 0x72;            /* src/compress.c:304-321            */
  trailer[2] = 0x45;
  trailer[3] = 0x38;
  trailer[4] = 0x50;
  trailer[5] = 0x90;
  trailer[6] = combined_crc >> 24;
  trailer[7] = (combined_crc >> 16) & 0xFF;
  trailer[8] = (combined_crc >> 8) & 0xFF;
  trailer[9] = combined_crc & 0xFF;
  xwrite(trailer, TRAILER_SIZE);

  if (stats) {
    fprintf(stderr, "lbzip2_b200: %lu batches, %lu chunks (%.1f per batch), "
            "%lu blocks, %u engines x %u chunks (%u set up in %.3f s); wall "
            "%.3f s; in lbz_compress_chunks %.3f s, staging copies %.3f s, "
            "output copies %.3f s (summed over threads)\n", stat_batches,
            stat_chunks, stat_batches ? (double)stat_chunks / stat_batches : 0.0,
            stat_blocks, num_slots, batch_cap, stat_engines, stat_init,
            now() - stat_t0, stat_gpu, stat_stage, stat_copy);
    fflush(stderr);             /* main.c:912-916 makes stderr fully buffered */
  }

  pqueue_uninit(stage_q);
  pqueue_uninit(reord_q);
}


static const struct task task_list[] = {
#ifdef LBZ_WITH_REF_SEQUENTIAL
  { "seq0",    can_ref0,    do_ref0    },
  { "seq1",    can_ref1,    do_ref1    },
  { "seq2",    can_ref2,    do_ref2    },
  { "seq3",    can_ref3,    do_ref3    },
#endif
  { "reorder", can_reorder, do_reorder },
  { "launch",  can_launch,  do_launch  },
  { "stage",   can_stage,   do_stage   },
  { "create",  can_create,  do_create  },
  { NULL,      NULL,        NULL       },
};

const struct process compression = {
  task_list,
  init,
  uninit,
  can_terminate,
  on_input_avail,
  on_write_complete,
};
