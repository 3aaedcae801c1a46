// engine.cu -- host side of the B200 bzip2 block-compression engine and its C ABI.
//
// Owns the device slab (layout: lbz_common.cuh), sequences the stage kernels
//   rle1.cu -> bwt.cu -> mtf.cu -> huffman.cu -> pack.cu
// on one CUDA stream per engine, and exposes
//   * the batch API (lbz_compress_chunks / _device / _stream),
//   * the reference-shaped per-block API (encoder_alloc_size .. transmit,
//     reference src/encode.h:29-36) on top of a pool of single-chunk engines,
//   * the stage hooks used by tests/.
// There is no CPU implementation of any stage in this library.
#include "lbz_common.cuh"
#include "../../include/lbzip2_b200.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <thread>
#include <future>
#include <chrono>
#include <deque>
#include <map>

struct BwtBuffers {           // must match bwt.cu
  const uint8_t *T;
  uint32_t *sa, *sa2, *rank;
  uint8_t *head;
  uint64_t *key, *key2;
  uint32_t *val, *val2, *pos, *pos2, *gs, *gs2;
  uint32_t *tstat, *gbase, *khist;
  void *agg;
  uint32_t *counters;
  uint32_t *epoch;
  int hints;
  void (*on_sorted)(void *);
  void *on_sorted_arg;
  uint8_t *bwt;
  uint32_t K;
  uint32_t *hbits, *cbits;
  uint32_t *tickets;
  uint32_t *wl, *wl_count;
  uint32_t wl_list_off;
};

extern "C" {
int lbz_launch_rle1(const LbzGeom *g, const uint8_t *d_in, const uint32_t *d_chunk_len, uint8_t *d_T,
                    LbzBlockMeta *d_meta, cudaStream_t st);
size_t lbz_rle1_scratch_words(const LbzGeom *g, uint32_t max_chunks);
int lbz_launch_rle1_tiles(const LbzGeom *g, const uint8_t *d_in, const uint32_t *d_chunk_len, uint8_t *d_T,
                          LbzBlockMeta *d_meta, uint32_t *d_scratch, cudaStream_t st);
int lbz_run_bwt(const LbzGeom *gp, LbzBlockMeta *d_meta, BwtBuffers B, uint32_t *h_counters,
                uint32_t *rounds_out, uint64_t *launches, const LbzTimers *tm, cudaStream_t st);
int lbz_launch_mtf(const LbzGeom *g, LbzBlockMeta *d_meta, const uint8_t *d_bwt, uint8_t *d_mtfrank,
                   uint16_t *d_mtfv, uint32_t *d_freq, int *d_parttab, uint32_t *d_emitcnt, cudaStream_t st);
uint32_t lbz_mtf_parts();
int lbz_launch_huffman(const LbzGeom *g, LbzBlockMeta *d_meta, uint16_t *d_mtfv, const uint32_t *d_freq,
                       void *d_coding, uint32_t cluster_factor, cudaStream_t st);
int lbz_launch_pack(const LbzGeom *g, LbzBlockMeta *d_meta, const uint16_t *d_mtfv, const void *d_coding,
                    uint8_t *d_out, uint32_t *d_out_off, uint32_t *d_total, uint8_t *d_packed, cudaStream_t st);
}

struct lbz_engine {
  int device = 0;
  int level = 9;
  uint32_t max_chunks = 0;
  LbzGeom g{};                 // g.nchunks = chunks of the current batch
  cudaStream_t st = nullptr;
  size_t dev_bytes = 0;
  uint64_t launches = 0;
  uint32_t last_rounds = 0;
  LbzTimers tm{};
  cudaEvent_t ev_call[2] = {nullptr, nullptr};
  double last_ms = 0.0;                 // device time of the last compress call (all batches)
  double stage_ms[LBZ_NSTAGE] = {};     // accumulated over the last call
  double k0_ms = 0.0;                   // dominant kernel: summed launch time, last call
  uint32_t k0_launches = 0;
  uint64_t k0_elements = 0;             // elements sorted per launch (sum over blocks), last batch
  double k0_elem_launches = 0.0;        // sum over the call's batches of elements x timed launches
  // device
  uint8_t *d_in = nullptr;     // max_chunks * mbs raw bytes
  uint32_t *d_chunk_len = nullptr;
  uint32_t *d_rle_scratch = nullptr;   // tile tables of the RLE1 stage (rle1.cu)
  uint8_t *d_T = nullptr, *d_bwt = nullptr, *d_mtfrank = nullptr, *d_head = nullptr;
  uint32_t *d_sa = nullptr, *d_sa2 = nullptr, *d_rank = nullptr;
  uint64_t *d_key = nullptr, *d_key2 = nullptr;
  uint32_t *d_val = nullptr, *d_val2 = nullptr, *d_pos = nullptr, *d_pos2 = nullptr, *d_gs = nullptr, *d_gs2 = nullptr;
  uint32_t *d_tstat = nullptr, *d_gbase = nullptr, *d_khist = nullptr, *d_counters = nullptr;
  uint32_t *d_tickets = nullptr;       // work-item counters of the persistent pass kernel (bwt.cu)
  uint32_t *d_wl = nullptr;            // its work lists: [0, wl_half) text passes, [wl_half, 2 wl_half) list passes, then 2 counts
  uint32_t wl_half = 0;
  uint64_t *d_agg = nullptr;
  uint32_t epoch = 0;
  uint16_t *d_mtfv = nullptr;
  uint32_t *d_freq = nullptr;
  int *d_parttab = nullptr;
  uint32_t *d_emitcnt = nullptr;
  LbzCoding *d_coding = nullptr;
  LbzBlockMeta *d_meta = nullptr;
  uint8_t *d_out = nullptr, *d_packed = nullptr, *d_packed2 = nullptr;   // packed2: second sub-batch of a lane
  cudaStream_t st_copy = nullptr;      // copy-out of finished sub-batches (two-lane mode)
  uint32_t *d_out_off = nullptr;
  // pinned host
  uint32_t *h_chunk_len = nullptr;
  LbzBlockMeta *h_meta = nullptr;
  uint32_t *h_counters = nullptr;   // [0..1] sort counters, [2] packed total
  std::vector<void *> allocs;
  struct PendingAlloc { void **pp; size_t bytes; };
  std::vector<PendingAlloc> pending;
  // Two-lane mode: large engines are split into two half-capacity lanes (this
  // object + `sib`) that work on disjoint chunk ranges of a batch on their own
  // streams, driven by two host threads, so that the latency-bound kernels of one
  // lane (rle1, huffman, round bookkeeping, host round trips) overlap with the
  // bandwidth-bound sort passes of the other, and H2D/D2H overlap with compute.
  int hints = 0;
  uint32_t bwt_k = 8;                  // bytes covered by the initial radix sort (LBZ_BWT_K, 5..8)
  void (*on_sorted)(void *) = nullptr;  // set per call by the two-lane driver
  void *on_sorted_arg = nullptr;
  lbz_engine *sib = nullptr;
  // blocks that enter the pipeline after the RLE1 stage (collected on the host, per-block API only):
  // block bytes + record are copied into their slot once the RLE1 kernels have run
  struct Inject { uint32_t slot; const uint8_t *bytes; LbzBlockMeta meta; };
  std::vector<Inject> inject;
  uint32_t total_chunks = 0;           // capacity of the whole engine (both lanes)
  cudaEvent_t ev_done = nullptr;
};

#define ENG_CHECK(x)                                                              \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      fprintf(stderr, "lbzip2_b200: CUDA error %s at %s:%d: %s\n",                \
              cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_));  \
      return -1;                                                                  \
    }                                                                             \
  } while (0)

// Device arrays are carved out of ONE allocation per engine (one cudaMalloc instead of forty: engine
// set-up is what a short CLI run pays first): dev_alloc() only records the request, dev_commit()
// allocates the slab and hands out 256-byte aligned pieces with 256 bytes of slack behind each.
template <class T>
static int dev_alloc(lbz_engine *e, T **p, size_t count) {
  const size_t bytes = (count * sizeof(T) + 256 + 255) & ~(size_t)255;
  e->pending.push_back({reinterpret_cast<void **>(p), bytes});
  return 0;
}
static int dev_commit(lbz_engine *e) {
  size_t total = 0;
  for (const auto &r : e->pending) total += r.bytes;
  void *q = nullptr;
  cudaError_t err = cudaMalloc(&q, total ? total : 256);
  if (err != cudaSuccess) {
    fprintf(stderr, "lbzip2_b200: cudaMalloc(%zu) failed: %s\n", total, cudaGetErrorString(err));
    return -1;
  }
  e->allocs.push_back(q);
  e->dev_bytes += total;
  uint8_t *base = reinterpret_cast<uint8_t *>(q);
  for (const auto &r : e->pending) { *r.pp = base; base += r.bytes; }
  e->pending.clear();
  return 0;
}

static uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

extern "C" const char *lbz_version(void) { return "lbzip2_b200 0.1 (sm_100a)"; }

extern "C" size_t lbz_bound(size_t n) {
  // incompressible data costs ~1.005 n (reference man page :63-64); per block <= 20 bits/symbol worst case
  return n + n / 32 + 8192 * (n / 100000 + 2) + 64;
}

// `mbs` = maximal block size = raw chunk size: level * 100000 for the batch API, any value in
// 1..900000 for the reference-shaped API (src/encode.c:121-122).
static lbz_engine *engine_create_mbs(int device, uint32_t mbs, int max_chunks) {
  const int level = (int)((mbs + 99999u) / 100000u);
  if (mbs < 1 || mbs > 900000 || max_chunks < 1 || max_chunks > 16384) {
    fprintf(stderr, "lbzip2_b200: bad engine parameters (block size %u, max_chunks %d)\n", mbs, max_chunks);
    return nullptr;
  }
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    fprintf(stderr, "lbzip2_b200: no CUDA device available (%s); this library has no CPU path\n",
            cudaGetErrorString(err));
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    fprintf(stderr, "lbzip2_b200: device %d out of range (%d devices)\n", device, ndev);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  lbz_engine *e = new lbz_engine();
  e->device = device;
  e->level = level;
  e->max_chunks = (uint32_t)max_chunks;
  e->total_chunks = (uint32_t)max_chunks;
  LbzGeom &g = e->g;
  g.mbs = mbs;
  g.S1 = round_up(g.mbs + 64, LBZ_TILE);
  g.S2 = round_up(g.mbs / 4 + 64, LBZ_TILE);
  g.stride = g.S1 + g.S2;
  g.tiles1 = g.S1 / LBZ_TILE;
  g.nchunks = 0;
  g.out_cap = round_up(g.S1 * 5 / 2 + 8192, 256);
  const size_t E = (size_t)e->max_chunks * g.stride;
  const size_t NB = 2 * (size_t)e->max_chunks;
  int rc = 0;
  rc |= cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess;
  for (int i = 0; i <= LBZ_NSTAGE; i++) rc |= cudaEventCreate(&e->tm.stage[i]) != cudaSuccess;
  for (int i = 0; i < 2 * LBZ_NK0; i++) rc |= cudaEventCreate(&e->tm.k0[i]) != cudaSuccess;
  rc |= cudaEventCreate(&e->ev_done) != cudaSuccess;
  rc |= cudaEventCreate(&e->ev_call[0]) != cudaSuccess;
  rc |= cudaEventCreate(&e->ev_call[1]) != cudaSuccess;
  e->tm.enabled = 1;
  { const char *hv = getenv("LBZ_CACHEHINT"); e->hints = hv ? atoi(hv) : 0; }
  { const char *kv = getenv("LBZ_BWT_K"); const int k = kv ? atoi(kv) : 8; e->bwt_k = (uint32_t)(k < 5 ? 5 : (k > 8 ? 8 : k)); }
  rc |= dev_alloc(e, &e->d_in, (size_t)e->max_chunks * g.mbs);
  rc |= dev_alloc(e, &e->d_chunk_len, e->max_chunks);
  rc |= dev_alloc(e, &e->d_rle_scratch, lbz_rle1_scratch_words(&g, e->max_chunks));
  rc |= dev_alloc(e, &e->d_T, E);
  rc |= dev_alloc(e, &e->d_bwt, E);
  rc |= dev_alloc(e, &e->d_mtfrank, E);
  rc |= dev_alloc(e, &e->d_head, E);
  rc |= dev_alloc(e, &e->d_sa, E);
  rc |= dev_alloc(e, &e->d_sa2, E);
  rc |= dev_alloc(e, &e->d_rank, E);
  rc |= dev_alloc(e, &e->d_key, E);
  rc |= dev_alloc(e, &e->d_key2, E);
  rc |= dev_alloc(e, &e->d_val, E);
  rc |= dev_alloc(e, &e->d_val2, E);
  rc |= dev_alloc(e, &e->d_pos, E);
  rc |= dev_alloc(e, &e->d_pos2, E);
  rc |= dev_alloc(e, &e->d_gs, E);
  rc |= dev_alloc(e, &e->d_gs2, E);
  rc |= dev_alloc(e, &e->d_tstat, NB * (g.S1 / 2048u) * 256);
  rc |= dev_alloc(e, &e->d_gbase, NB * 256 * 6);
  rc |= dev_alloc(e, &e->d_khist, NB * 256 * 5);
  rc |= dev_alloc(e, &e->d_agg, 2 * NB * g.tiles1 * 2);   // TileAgg (16 B) for two lists
  rc |= dev_alloc(e, &e->d_counters, 8);
  rc |= dev_alloc(e, &e->d_tickets, 1024);
  e->wl_half = (uint32_t)(E / LBZ_TILE) + 64u;
  rc |= dev_alloc(e, &e->d_wl, 2 * (size_t)e->wl_half + 8);
  rc |= dev_alloc(e, &e->d_mtfv, E);
  rc |= dev_alloc(e, &e->d_freq, NB * 260);
  rc |= dev_alloc(e, &e->d_parttab, NB * lbz_mtf_parts() * 256);
  rc |= dev_alloc(e, &e->d_emitcnt, NB * 16);
  rc |= dev_alloc(e, &e->d_coding, NB);
  rc |= dev_alloc(e, &e->d_meta, NB);
  rc |= dev_alloc(e, &e->d_out, NB * (size_t)g.out_cap);
  rc |= dev_alloc(e, &e->d_packed, lbz_bound((size_t)e->max_chunks * g.mbs));
  rc |= dev_alloc(e, &e->d_packed2, lbz_bound((size_t)e->max_chunks * g.mbs));
  rc |= cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking) != cudaSuccess;
  rc |= dev_alloc(e, &e->d_out_off, NB);
  rc |= dev_commit(e);
  rc |= cudaHostAlloc((void **)&e->h_chunk_len, e->max_chunks * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess;
  rc |= cudaHostAlloc((void **)&e->h_meta, NB * sizeof(LbzBlockMeta), cudaHostAllocDefault) != cudaSuccess;
  rc |= cudaHostAlloc((void **)&e->h_counters, 8 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess;
  if (rc) {
    fprintf(stderr, "lbzip2_b200: engine allocation failed\n");
    lbz_engine_destroy(e);
    return nullptr;
  }
  cudaMemsetAsync(e->d_meta, 0, NB * sizeof(LbzBlockMeta), e->st);
  cudaMemsetAsync(e->d_tstat, 0, NB * (g.S1 / 2048u) * 256 * sizeof(uint32_t), e->st);
  cudaMemsetAsync(e->d_counters, 0, 8 * sizeof(uint32_t), e->st);
  cudaMemsetAsync(e->d_tickets, 0, 1024 * sizeof(uint32_t), e->st);
  cudaStreamSynchronize(e->st);
  return e;
}

static lbz_engine *engine_create_one(int device, int level, int max_chunks) {
  if (level < 1 || level > 9) {
    fprintf(stderr, "lbzip2_b200: bad engine parameters (level %d)\n", level);
    return nullptr;
  }
  return engine_create_mbs(device, (uint32_t)level * 100000u, max_chunks);
}

extern "C" lbz_engine *lbz_engine_create(int device, int level, int max_chunks) {
  const char *ev = getenv("LBZ_LANES");
  const int lanes = (max_chunks >= 32 && !(ev && atoi(ev) == 1)) ? 2 : 1;
  const int per = (max_chunks + lanes - 1) / lanes;
  lbz_engine *e = engine_create_one(device, level, per);
  if (!e) return nullptr;
  e->total_chunks = (uint32_t)(per * lanes);
  if (lanes == 2) {
    e->sib = engine_create_one(device, level, per);
    if (!e->sib) { lbz_engine_destroy(e); return nullptr; }
    e->sib->total_chunks = (uint32_t)per;
  }
  return e;
}

extern "C" void lbz_engine_destroy(lbz_engine *e) {
  if (!e) return;
  if (e->sib) { lbz_engine_destroy(e->sib); e->sib = nullptr; }
  if (e->ev_done) cudaEventDestroy(e->ev_done);
  cudaSetDevice(e->device);
  if (e->st) { cudaStreamSynchronize(e->st); cudaStreamDestroy(e->st); }
  if (e->st_copy) { cudaStreamSynchronize(e->st_copy); cudaStreamDestroy(e->st_copy); }
  for (void *p : e->allocs) cudaFree(p);
  for (int i = 0; i <= LBZ_NSTAGE; i++) if (e->tm.stage[i]) cudaEventDestroy(e->tm.stage[i]);
  for (int i = 0; i < 2 * LBZ_NK0; i++) if (e->tm.k0[i]) cudaEventDestroy(e->tm.k0[i]);
  for (int i = 0; i < 2; i++) if (e->ev_call[i]) cudaEventDestroy(e->ev_call[i]);
  if (e->h_chunk_len) cudaFreeHost(e->h_chunk_len);
  if (e->h_meta) cudaFreeHost(e->h_meta);
  if (e->h_counters) cudaFreeHost(e->h_counters);
  delete e;
}

extern "C" void *lbz_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void lbz_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" uint64_t lbz_engine_launches(const lbz_engine *e) { return e->launches + (e->sib ? e->sib->launches : 0); }
extern "C" uint32_t lbz_engine_last_rounds(const lbz_engine *e) {
  return (e->sib && e->sib->last_rounds > e->last_rounds) ? e->sib->last_rounds : e->last_rounds;
}
extern "C" double lbz_engine_last_ms(const lbz_engine *e) { return e->last_ms; }
// stage times are summed over the lanes (the lanes overlap, so the sum exceeds the call time)
extern "C" void lbz_engine_stage_ms(const lbz_engine *e, double *out7) {
  for (int i = 0; i < 7; i++) out7[i] = e->stage_ms[i] + (e->sib ? e->sib->stage_ms[i] : 0.0);
}
// dominant kernel: summed launch time, launches, average elements per launch
extern "C" void lbz_engine_k0_stats(const lbz_engine *e, double *sum_ms, uint32_t *launches, uint64_t *elements) {
  double ms = e->k0_ms; uint64_t nl = e->k0_launches; double el = e->k0_elem_launches;
  if (e->sib) { ms += e->sib->k0_ms; nl += e->sib->k0_launches; el += e->sib->k0_elem_launches; }
  *sum_ms = ms; *launches = (uint32_t)nl; *elements = nl ? (uint64_t)(el / (double)nl) : 0;
}
extern "C" size_t lbz_engine_device_bytes(const lbz_engine *e) { return e->dev_bytes + (e->sib ? e->sib->dev_bytes : 0); }
extern "C" uint32_t lbz_dbg_num_slots(const lbz_engine *e) { return 2 * e->g.nchunks; }
extern "C" int lbz_dbg_set_chunks(lbz_engine *e, uint32_t nchunks) {
  if (nchunks > e->max_chunks) return -1;
  e->g.nchunks = nchunks;
  return 0;
}

static BwtBuffers bwt_buffers(lbz_engine *e) {
  BwtBuffers B;
  B.T = e->d_T; B.sa = e->d_sa; B.sa2 = e->d_sa2; B.rank = e->d_rank; B.head = e->d_head;
  B.key = e->d_key; B.key2 = e->d_key2; B.val = e->d_val; B.val2 = e->d_val2;
  B.pos = e->d_pos; B.pos2 = e->d_pos2; B.gs = e->d_gs; B.gs2 = e->d_gs2;
  B.tstat = e->d_tstat; B.gbase = e->d_gbase; B.khist = e->d_khist; B.agg = e->d_agg;
  B.counters = e->d_counters; B.epoch = &e->epoch; B.bwt = e->d_bwt;
  B.hbits = reinterpret_cast<uint32_t *>(e->d_head);
  B.cbits = B.hbits + ((size_t)e->max_chunks * e->g.stride) / 32 + 64;   // the byte array holds both bitmaps with room to spare
  B.tickets = e->d_tickets;
  B.wl = e->d_wl; B.wl_list_off = e->wl_half; B.wl_count = e->d_wl + 2 * (size_t)e->wl_half;
  B.K = e->bwt_k; B.hints = e->hints; B.on_sorted = e->on_sorted; B.on_sorted_arg = e->on_sorted_arg;
  return B;
}

// Set up the chunk table for `n` raw bytes (<= max_chunks chunks).
static int set_chunks(lbz_engine *e, size_t n) {
  const size_t mbs = e->g.mbs;
  const size_t nc = (n + mbs - 1) / mbs;
  if (nc > e->max_chunks) {
    fprintf(stderr, "lbzip2_b200: batch of %zu chunks exceeds engine capacity %u\n", nc, e->max_chunks);
    return -1;
  }
  e->g.nchunks = (uint32_t)nc;
  for (size_t c = 0; c < nc; c++) e->h_chunk_len[c] = (uint32_t)((c + 1) * mbs <= n ? mbs : n - c * mbs);
  if (nc) ENG_CHECK(cudaMemcpyAsync(e->d_chunk_len, e->h_chunk_len, nc * sizeof(uint32_t), cudaMemcpyHostToDevice, e->st));
  return 0;
}

static int run_stage(lbz_engine *e, int stage, const uint8_t *d_in, uint8_t *d_packed) {
  const LbzGeom *g = &e->g;
  const uint32_t nb = 2 * g->nchunks;
  switch (stage) {
    case LBZ_ST_RLE1: {
      static int v1 = -1;
      if (v1 < 0) { const char *ev = getenv("LBZ_RLE_V1"); v1 = (ev && atoi(ev) != 0) ? 1 : 0; }
      if (v1) { e->launches += 1; return lbz_launch_rle1(g, d_in, e->d_chunk_len, e->d_T, e->d_meta, e->st); }
      e->launches += 10;
      return lbz_launch_rle1_tiles(g, d_in, e->d_chunk_len, e->d_T, e->d_meta, e->d_rle_scratch, e->st);
    }
    case LBZ_ST_BWT:
      return lbz_run_bwt(g, e->d_meta, bwt_buffers(e), e->h_counters, &e->last_rounds, &e->launches, &e->tm, e->st);
    case LBZ_ST_MTF:
      e->launches += 4;
      return lbz_launch_mtf(g, e->d_meta, e->d_bwt, e->d_mtfrank, e->d_mtfv, e->d_freq, e->d_parttab, e->d_emitcnt, e->st);
    case LBZ_ST_HUFFMAN:
      e->launches += 1;
      return lbz_launch_huffman(g, e->d_meta, e->d_mtfv, e->d_freq, e->d_coding, CLUSTER_FACTOR, e->st);
    case LBZ_ST_PACK:
      e->launches += 3;
      (void)nb;
      return lbz_launch_pack(g, e->d_meta, e->d_mtfv, e->d_coding, e->d_out, e->d_out_off, e->d_counters + 2,
                             d_packed, e->st);
  }
  return -1;
}

// Run the whole pipeline for the current chunk table; results: h_meta (host),
// packed bytes at d_packed (device), *total = packed byte count.
static int run_pipeline(lbz_engine *e, const uint8_t *d_in, uint8_t *d_packed, size_t *total) {
  const uint32_t nb = 2 * e->g.nchunks;
  *total = 0;
  if (nb == 0) return 0;
  // stage boundaries: [0] start, [1] after rle1, [2] after initial sort, [3] after refinement,
  // [4] after last-column gather, [5] after mtf, [6] after huffman, [7] after pack+gather
  static const int after[5] = {1, 4, 5, 6, 7};
  cudaEventRecord(e->tm.stage[0], e->st);
  for (int s = LBZ_ST_RLE1; s <= LBZ_ST_PACK; s++) {
    if (run_stage(e, s, d_in, d_packed)) return -1;
    if (s == LBZ_ST_RLE1) {
      for (const lbz_engine::Inject &in : e->inject) {
        ENG_CHECK(cudaMemcpyAsync(e->d_T + lbz_slot_off(e->g, in.slot), in.bytes, in.meta.n, cudaMemcpyHostToDevice, e->st));
        ENG_CHECK(cudaMemcpyAsync(e->d_meta + in.slot, &in.meta, sizeof(LbzBlockMeta), cudaMemcpyHostToDevice, e->st));
      }
    }
    cudaEventRecord(e->tm.stage[after[s]], e->st);
  }
  ENG_CHECK(cudaMemcpyAsync(e->h_meta, e->d_meta, nb * sizeof(LbzBlockMeta), cudaMemcpyDeviceToHost, e->st));
  ENG_CHECK(cudaMemcpyAsync(e->h_counters + 2, e->d_counters + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->st));
  ENG_CHECK(cudaStreamSynchronize(e->st));
  for (int i = 0; i < 7; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->tm.stage[i], e->tm.stage[i + 1]) == cudaSuccess) e->stage_ms[i] += ms;
  }
  e->k0_elements = 0;
  for (uint32_t b = 0; b < nb; b++) e->k0_elements += e->h_meta[b].n;
  for (int i = 0; i < (int)e->bwt_k && i < LBZ_NK0; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->tm.k0[2 * i], e->tm.k0[2 * i + 1]) == cudaSuccess) {
      e->k0_ms += ms; e->k0_launches++; e->k0_elem_launches += (double)e->k0_elements;
    }
  }
  size_t sum = 0;
  for (uint32_t b = 0; b < nb; b++) {
    const LbzBlockMeta &m = e->h_meta[b];
    if (m.n == 0) continue;
    if (m.pad_[0] != 8u * m.out_len) {
      fprintf(stderr, "lbzip2_b200: internal error: block slot %u packed %u bits, expected %u\n", b, m.pad_[0],
              8u * m.out_len);
      return -1;
    }
    sum += m.out_len;
  }
  if (sum != e->h_counters[2]) {
    fprintf(stderr, "lbzip2_b200: internal error: gathered %u bytes, expected %zu\n", e->h_counters[2], sum);
    return -1;
  }
  *total = sum;
  return 0;
}

static void call_end(lbz_engine *e) {
  cudaEventRecord(e->ev_call[1], e->st);
  cudaEventSynchronize(e->ev_call[1]);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e->ev_call[0], e->ev_call[1]);
  e->last_ms = ms;
}

static size_t fill_recs(lbz_engine *e, uint64_t raw_base, lbz_block_rec *recs, size_t max_recs, size_t have) {
  const uint32_t nb = 2 * e->g.nchunks;
  size_t k = have;
  for (uint32_t b = 0; b < nb; b++) {
    const LbzBlockMeta &m = e->h_meta[b];
    if (m.n == 0) continue;
    if (recs && k < max_recs) {
      lbz_block_rec &r = recs[k];
      const uint64_t chunk_off = (uint64_t)(b >> 1) * e->g.mbs;
      r.raw_offset = raw_base + chunk_off + ((b & 1) ? e->h_meta[b - 1].raw_len : 0u);
      r.raw_len = m.raw_len; r.nblock = m.n; r.crc = m.crc; r.bwt_idx = m.bwt_idx; r.tie_count = m.tie_count;
      r.nmtf = m.nmtf; r.num_trees = m.num_trees; r.num_selectors = m.num_selectors; r.out_len = m.out_len;
      r.reserved = 0;
    }
    k++;
  }
  return k;
}

static void reset_call_stats(lbz_engine *e) {
  for (int i = 0; i < LBZ_NSTAGE; i++) e->stage_ms[i] = 0.0;
  e->k0_ms = 0.0; e->k0_launches = 0; e->k0_elem_launches = 0.0;
}

// One lane: chunk table, (H2D,) all stages.  Leaves the packed blocks in
// dst_dev (or e->d_packed) and the block records in e->h_meta.
static int lane_run(lbz_engine *e, const uint8_t *src, bool src_on_device, size_t len, uint8_t *dst_dev, size_t *total) {
  if (cudaSetDevice(e->device) != cudaSuccess) return -1;
  if (set_chunks(e, len)) return -1;
  const uint8_t *d_in = src;
  if (!src_on_device) {
    ENG_CHECK(cudaMemcpyAsync(e->d_in, src, len, cudaMemcpyHostToDevice, e->st));
    d_in = e->d_in;
  }
  return run_pipeline(e, d_in, dst_dev ? dst_dev : e->d_packed, total);
}

static size_t fill_recs_from(const LbzGeom &g, const LbzBlockMeta *metas, uint32_t nb, uint64_t raw_base,
                             lbz_block_rec *recs, size_t max_recs, size_t have) {
  size_t k = have;
  for (uint32_t b = 0; b < nb; b++) {
    const LbzBlockMeta &m = metas[b];
    if (m.n == 0) continue;
    if (recs && k < max_recs) {
      lbz_block_rec &r = recs[k];
      const uint64_t chunk_off = (uint64_t)(b >> 1) * g.mbs;
      r.raw_offset = raw_base + chunk_off + ((b & 1) ? metas[b - 1].raw_len : 0u);
      r.raw_len = m.raw_len; r.nblock = m.n; r.crc = m.crc; r.bwt_idx = m.bwt_idx; r.tie_count = m.tie_count;
      r.nmtf = m.nmtf; r.num_trees = m.num_trees; r.num_selectors = m.num_selectors; r.out_len = m.out_len;
      r.reserved = 0;
    }
    k++;
  }
  return k;
}

// Compress up to total_chunks chunks (one "super batch") from `src` into `dst`
// (host or device memory, `on_device`).
//
// Two-lane engines cut the batch into sub-batches in stream order, alternating between
// the lanes (two host threads, two streams); the copy-out of a finished sub-batch and
// the H2D of the other lane overlap the kernels.  Default: one sub-batch per lane.
// LBZ_SPLIT4=1 cuts four unequal ones,
//     sb0 (lane A, large)  sb1 (lane B, small)  sb2 (lane A, small)  sb3 (lane B, large),
// meant to let the lanes drift out of phase (one lane's latency-bound RLE1/Huffman CTAs
// next to the other's sort passes); on the benchmark batch the smaller launches cost
// more than the overlap gains.
struct SubBatch {
  lbz_engine *lane = nullptr;
  size_t in_off = 0, in_len = 0;
  uint8_t *packed = nullptr;
  size_t total = 0;
  int rc = 0;
  bool done = false;
  std::vector<LbzBlockMeta> metas;
};

static int super_batch(lbz_engine *e, const uint8_t *src, size_t len, uint8_t *dst, size_t dst_cap, bool on_device,
                       bool dst_on_device, uint64_t raw_base, size_t *written, lbz_block_rec *recs, size_t max_recs,
                       size_t *nrec) {
  const size_t mbs = e->g.mbs;
  const size_t nc = (len + mbs - 1) / mbs;
  const cudaMemcpyKind kind = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (!e->sib || nc < 2) {
    size_t total;
    if (lane_run(e, src, on_device, len, nullptr, &total)) return -1;
    if (total > dst_cap) { fprintf(stderr, "lbzip2_b200: output buffer too small\n"); return -2; }
    ENG_CHECK(cudaMemcpyAsync(dst, e->d_packed, total, kind, e->st));
    ENG_CHECK(cudaStreamSynchronize(e->st));
    *nrec = fill_recs(e, raw_base, recs, max_recs, *nrec);
    *written = total;
    return 0;
  }
  lbz_engine *A = e, *B = e->sib;
  const size_t nA = (nc + 1) / 2, nB = nc - nA;
  static int split4 = -1;
  // the four-way split measured slower on the 100 MB text batch (5.08 vs 5.90 GB/s): off by default
  if (split4 < 0) { const char *ev = getenv("LBZ_SPLIT4"); split4 = ev ? atoi(ev) : 0; }
  size_t cnt[4];
  int nsub;
  if (split4 && nc >= 16) {
    const size_t a0 = (nA * 72 + 99) / 100, b0 = (nB * 28) / 100 ? (nB * 28) / 100 : 1;
    cnt[0] = a0; cnt[1] = b0; cnt[2] = nA - a0; cnt[3] = nB - b0;
    nsub = 4;
  } else {
    cnt[0] = nA; cnt[1] = nB;
    nsub = 2;
  }
  SubBatch sub[4];
  {
    size_t c0 = 0;
    for (int k = 0; k < nsub; k++) {
      sub[k].lane = (k & 1) ? B : A;
      sub[k].in_off = c0 * mbs;
      sub[k].in_len = (c0 + cnt[k]) * mbs <= len ? cnt[k] * mbs : len - c0 * mbs;
      sub[k].packed = (k < 2) ? sub[k].lane->d_packed : sub[k].lane->d_packed2;
      c0 += cnt[k];
    }
  }
  // copy-out in stream order as soon as every earlier sub-batch's size is known
  std::mutex mu;
  int next_copy = 0;
  size_t out_off = 0;
  int copy_rc = 0;
  auto advance_copies = [&]() {                    // called with `mu` held
    while (next_copy < nsub && sub[next_copy].done && !copy_rc) {
      SubBatch &sb = sub[next_copy];
      if (sb.rc) { copy_rc = sb.rc; break; }
      if (out_off + sb.total > dst_cap) { copy_rc = -2; break; }
      if (sb.total && cudaMemcpyAsync(dst + out_off, sb.packed, sb.total, kind, A->st_copy) != cudaSuccess) { copy_rc = -1; break; }
      out_off += sb.total;
      next_copy++;
    }
  };
  auto run_lane = [&](int first) {
    for (int k = first; k < nsub; k += 2) {
      SubBatch &sb = sub[k];
      size_t total = 0;
      int rc = 0;
      if (sb.in_len) {
        rc = lane_run(sb.lane, src + sb.in_off, on_device, sb.in_len, sb.packed, &total);
        if (rc == 0) sb.metas.assign(sb.lane->h_meta, sb.lane->h_meta + 2 * sb.lane->g.nchunks);
      }
      std::lock_guard<std::mutex> lk(mu);
      sb.rc = rc; sb.total = total; sb.done = true;
      advance_copies();
      if (rc) break;
    }
    std::lock_guard<std::mutex> lk(mu);             // a failed lane must not leave the other one waiting
    for (int k = first; k < nsub; k += 2) if (!sub[k].done) { sub[k].rc = -1; sub[k].done = true; }
    advance_copies();
  };
  std::thread tb(run_lane, 1);
  run_lane(0);
  tb.join();
  if (cudaSetDevice(A->device) != cudaSuccess) return -1;
  if (copy_rc == 0 && next_copy == nsub && cudaStreamSynchronize(A->st_copy) != cudaSuccess) copy_rc = -1;
  if (copy_rc || next_copy != nsub) {
    if (copy_rc == -2) fprintf(stderr, "lbzip2_b200: output buffer too small\n");
    cudaStreamSynchronize(A->st_copy);
    return copy_rc ? copy_rc : -1;
  }
  for (int k = 0; k < nsub; k++)
    *nrec = fill_recs_from(e->g, sub[k].metas.data(), (uint32_t)sub[k].metas.size(), raw_base + sub[k].in_off, recs, max_recs, *nrec);
  *written = out_off;
  return 0;
}

static int compress_any(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, bool on_device,
                        bool dst_on_device, size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs) {
  if (!e) return -1;
  ENG_CHECK(cudaSetDevice(e->device));
  const size_t batch_bytes = (size_t)e->total_chunks * e->g.mbs;
  // inputs larger than the engine are cut into consecutive batches of total_chunks chunks
  // (host and device legs alike: `in + pos` / `out + o` are plain pointer arithmetic)
  size_t o = 0, nrec = 0;
  reset_call_stats(e);
  if (e->sib) reset_call_stats(e->sib);
  cudaEventRecord(e->ev_call[0], e->st);
  for (size_t pos = 0; pos < n; pos += batch_bytes) {
    const size_t len = (n - pos < batch_bytes) ? n - pos : batch_bytes;
    size_t written = 0;
    const int rc = super_batch(e, in + pos, len, out + o, out_cap - o, on_device, dst_on_device, pos, &written, recs, max_recs, &nrec);
    if (rc) return rc;
    o += written;
  }
  call_end(e);
  if (out_len) *out_len = o;
  if (num_recs) *num_recs = nrec;
  return 0;
}

extern "C" int lbz_compress_chunks(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                   size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs) {
  return compress_any(e, in, n, out, out_cap, false, false, out_len, recs, max_recs, num_recs);
}

// Host input (pinned memory gives the full link rate), DEVICE output: what a multi-process writer
// uses -- the blocks stay in HBM until the stream offsets of all ranks' blocks are known, then
// lbz_scatter_to_host puts every block at its place in the one output stream.
extern "C" int lbz_compress_chunks_h2d(lbz_engine *e, const uint8_t *in, size_t n, void *d_out, size_t out_cap,
                                       size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs) {
  if (e && out_cap < lbz_bound(n)) { fprintf(stderr, "lbzip2_b200: device output buffer too small\n"); return -2; }
  return compress_any(e, in, n, reinterpret_cast<uint8_t *>(d_out), out_cap, false, true, out_len, recs, max_recs, num_recs);
}

// Optional (LBZ_SCATTER_KERNEL=1; measured SLOWER than the per-block copies on B200 -- SM stores over
// PCIe reach a fraction of the copy engines' rate: 30.2 vs 18.3 ms per 100 MB step at N=2,
// profiles/r02_n2_scatter_kernel_vs_memcpy.log -- so the copy engines are the default).
// One launch instead of one copy per block: every CTA row moves one block from HBM to its place in
// the (CUDA-registered, hence device-visible) host buffer, 4-byte words assembled from the two
// source words they straddle (source and destination are byte-aligned independently).
__global__ void __launch_bounds__(256)
k_scatter_blocks(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, uint8_t *__restrict__ dst,
                 const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ len) {
  const uint32_t i = blockIdx.y;
  const uint64_t n = len[i];
  const uint8_t *s = src + src_off[i];
  uint8_t *d = dst + dst_off[i];
  // head bytes up to the first 4-byte boundary of the destination
  const uint64_t head0 = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 3u)) & 3u;
  const uint64_t head = n < head0 ? n : head0;
  const uint64_t words = (n - head) / 4u;
  const uint64_t tail0 = head + 4u * words;
  if (blockIdx.x == 0) {
    if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
    if (threadIdx.x < n - tail0) d[tail0 + threadIdx.x] = s[tail0 + threadIdx.x];
  }
  const uint8_t *sb = s + head;
  const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(sb) & 3u);
  const uint32_t *sw = reinterpret_cast<const uint32_t *>(sb - (sh >> 3));     // aligned word that holds sb[0]
  uint32_t *dw = reinterpret_cast<uint32_t *>(d + head);
  for (uint64_t w = (uint64_t)blockIdx.x * 256u + threadIdx.x; w < words; w += (uint64_t)gridDim.x * 256u) {
    const uint32_t lo = sw[w];
    const uint32_t hi = sh ? sw[w + 1] : 0u;            // within the source allocation: the block has >= 4 more bytes or slack follows
    dw[w] = sh ? __funnelshift_r(lo, hi, sh) : lo;
  }
}

extern "C" int lbz_scatter_to_host(lbz_engine *e, const void *d_src, const uint64_t *src_off, void *h_dst,
                                   const uint64_t *dst_off, const uint64_t *len, size_t count) {
  if (!e) return -1;
  if (count == 0) return 0;
  ENG_CHECK(cudaSetDevice(e->device));
  const uint8_t *s = reinterpret_cast<const uint8_t *>(d_src);
  uint8_t *d = reinterpret_cast<uint8_t *>(h_dst);
  // a pinned / registered destination is visible from the device: one kernel writes all blocks
  cudaPointerAttributes at;
  void *dev_view = nullptr;
  if (cudaPointerGetAttributes(&at, h_dst) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
    dev_view = at.devicePointer;
  else
    cudaGetLastError();
  static int use_kernel = -1;
  if (use_kernel < 0) { const char *ev = getenv("LBZ_SCATTER_KERNEL"); use_kernel = ev ? atoi(ev) : 0; }
  if (dev_view && use_kernel && count <= 65535) {
    uint64_t *tab = nullptr;                           // [src_off | dst_off | len] on the device
    ENG_CHECK(cudaMallocAsync((void **)&tab, 3 * count * sizeof(uint64_t), e->st));
    ENG_CHECK(cudaMemcpyAsync(tab, src_off, count * sizeof(uint64_t), cudaMemcpyHostToDevice, e->st));
    ENG_CHECK(cudaMemcpyAsync(tab + count, dst_off, count * sizeof(uint64_t), cudaMemcpyHostToDevice, e->st));
    ENG_CHECK(cudaMemcpyAsync(tab + 2 * count, len, count * sizeof(uint64_t), cudaMemcpyHostToDevice, e->st));
    k_scatter_blocks<<<dim3(8, (unsigned)count), 256, 0, e->st>>>(s, tab, reinterpret_cast<uint8_t *>(dev_view), tab + count,
                                                                 tab + 2 * count);
    ENG_CHECK(cudaGetLastError());
    ENG_CHECK(cudaFreeAsync(tab, e->st));
    ENG_CHECK(cudaStreamSynchronize(e->st));
    return 0;
  }
  for (size_t i = 0; i < count; i++)
    if (len[i]) ENG_CHECK(cudaMemcpyAsync(d + dst_off[i], s + src_off[i], len[i], cudaMemcpyDeviceToHost, (i & 1) ? e->st_copy : e->st));
  ENG_CHECK(cudaStreamSynchronize(e->st));
  ENG_CHECK(cudaStreamSynchronize(e->st_copy));
  return 0;
}

extern "C" int lbz_compress_chunks_device(lbz_engine *e, const void *d_in, size_t n, void *d_out, size_t out_cap,
                                          size_t *out_len, lbz_block_rec *recs, size_t max_recs, size_t *num_recs) {
  if (e && out_cap < lbz_bound(n)) { fprintf(stderr, "lbzip2_b200: device output buffer too small\n"); return -2; }
  return compress_any(e, reinterpret_cast<const uint8_t *>(d_in), n, reinterpret_cast<uint8_t *>(d_out), out_cap, true, true,
                      out_len, recs, max_recs, num_recs);
}

extern "C" int lbz_compress_stream(lbz_engine *e, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                   size_t *out_len) {
  if (!e || out_cap < 14) return -1;
  const size_t mbs = e->g.mbs;
  const size_t maxrec = 2 * ((n + mbs - 1) / mbs) + 2;
  std::vector<lbz_block_rec> recs(maxrec);
  size_t blen = 0, nrec = 0;
  out[0] = 'B'; out[1] = 'Z'; out[2] = 'h'; out[3] = (uint8_t)('0' + e->level);     // compress.c:290-301
  if (lbz_compress_chunks(e, in, n, out + 4, out_cap - 14, &blen, recs.data(), maxrec, &nrec)) return -1;
  uint32_t cc = 0;
  for (size_t i = 0; i < nrec; i++) cc = ((cc << 1) ^ (cc >> 31)) ^ ~recs[i].crc;     // combine_crc, encode.h:38
  uint8_t *p = out + 4 + blen;
  static const uint8_t eos[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};               // compress.c:304-321
  memcpy(p, eos, 6);
  p[6] = (uint8_t)(cc >> 24); p[7] = (uint8_t)(cc >> 16); p[8] = (uint8_t)(cc >> 8); p[9] = (uint8_t)cc;
  if (out_len) *out_len = 4 + blen + 10;
  return 0;
}

// ---------------------------------------------------------------------------
// Stage hooks (tests only)
extern "C" int lbz_dbg_load(lbz_engine *e, const uint8_t *in, size_t n) {
  ENG_CHECK(cudaSetDevice(e->device));
  if (set_chunks(e, n)) return -1;
  if (n) ENG_CHECK(cudaMemcpyAsync(e->d_in, in, n, cudaMemcpyHostToDevice, e->st));
  ENG_CHECK(cudaStreamSynchronize(e->st));
  return 0;
}
extern "C" int lbz_dbg_run(lbz_engine *e, int stage) {
  ENG_CHECK(cudaSetDevice(e->device));
  if (run_stage(e, stage, e->d_in, e->d_packed)) return -1;
  ENG_CHECK(cudaStreamSynchronize(e->st));
  ENG_CHECK(cudaGetLastError());
  return 0;
}
static int dbg_locate(lbz_engine *e, int array, uint32_t slot, void **p, size_t *cap) {
  const LbzGeom &g = e->g;
  if (slot >= 2 * e->max_chunks) return -1;
  const size_t off = lbz_slot_off(g, slot), sc = lbz_slot_cap(g, slot);
  switch (array) {
    case LBZ_AR_TEXT: *p = e->d_T + off; *cap = sc; return 0;
    case LBZ_AR_BWT: *p = e->d_bwt + off; *cap = sc; return 0;
    case LBZ_AR_MTFV: *p = e->d_mtfv + off; *cap = sc * 2; return 0;
    case LBZ_AR_FREQ: *p = e->d_freq + (size_t)slot * 260; *cap = 260 * 4; return 0;
    case LBZ_AR_CODING: *p = e->d_coding + slot; *cap = sizeof(LbzCoding); return 0;
    case LBZ_AR_OUT: *p = e->d_out + (size_t)slot * g.out_cap; *cap = g.out_cap; return 0;
    case LBZ_AR_META: *p = e->d_meta + slot; *cap = sizeof(LbzBlockMeta); return 0;
    case LBZ_AR_SA: *p = e->d_sa + off; *cap = sc * 4; return 0;
  }
  return -1;
}
extern "C" int lbz_dbg_read(lbz_engine *e, int array, uint32_t slot, void *dst, size_t bytes) {
  void *p; size_t cap;
  ENG_CHECK(cudaSetDevice(e->device));
  if (dbg_locate(e, array, slot, &p, &cap) || bytes > cap) return -1;
  ENG_CHECK(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int lbz_dbg_write(lbz_engine *e, int array, uint32_t slot, const void *src, size_t bytes) {
  void *p; size_t cap;
  ENG_CHECK(cudaSetDevice(e->device));
  if (dbg_locate(e, array, slot, &p, &cap) || bytes > cap) return -1;
  ENG_CHECK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------
// Reference-shaped per-block API (reference src/encode.h:29-36).
//
// The caller owns `struct encoder_state` (plain malloc/free, no destructor:
// src/compress.c:89,223), so it only holds a handle; the device context is a
// pooled single-chunk engine acquired by the first collect() and released by
// transmit().  Calls on distinct states may come from different threads
// concurrently (src/compress.c:81,103,113,217,222); every pooled engine has its
// own stream, so concurrent blocks overlap on the GPU.
struct encoder_state {
  uint32_t magic;
  uint32_t max_block_size;
  uint32_t cluster_factor;
  int32_t pool_slot;         // -1: no device context yet
  uint32_t raw_len;          // raw bytes staged so far
  uint32_t done;             // encode() has run
  uint32_t out_len;
  uint32_t crc;
  int32_t rle_state;         // mirrors the reference's collect() state for resumed calls
  uint32_t rle_char;
  uint32_t staged_cap;
  uint32_t tree_cost;        // bits: prefix-code transmission + payload cost reported by the Huffman kernel
  uint32_t nblock;           // RLE1 output bytes so far (host-side count, src/encode.c:333)
  uint32_t mode;             // 0: raw bytes staged, RLE1 + CRC run on the device; 1: block collected on the host (see collect())
  uint32_t hcrc;             // mode 1: running CRC of the consumed bytes (un-inverted, src/encode.c:542)
  uint32_t used[8];          // mode 1: used-byte map
  uint8_t *staged;           // staging of the raw bytes of this block (lives after the struct)
};

extern "C" void failx(int x, const char *fmt, ...) __attribute__((weak));
static void (*g_fatal)(const char *) = nullptr;
extern "C" void lbz_set_fatal_handler(void (*fn)(const char *msg)) { g_fatal = fn; }

// CRC-32/BZIP2 table (reference src/crctab.c:6, declared in src/decode.h:70), generated at load
// time from the polynomial (build-aux/make-crctab.pl:29-33): poly 0x04C11DB7, MSB first.
extern "C" { uint32_t crc_table[256]; }
__attribute__((constructor)) static void lbz_init_crc_table() {
  // (32 hardware work queues instead of the default 8 -- CUDA_DEVICE_MAX_CONNECTIONS -- were tried for the
  // per-block shims, which keep one stream per worker thread busy: lbzip2_gpu -d fell from 35-44 to 18 MB/s,
  // profiles/r02_run20_shim_queues.log; the default stays.)
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t r = i << 24;
    for (int k = 0; k < 8; k++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04C11DB7u : (r << 1);
    crc_table[i] = r;
  }
}

namespace {
#define POOL_MAX 256
struct Pool {
  std::mutex mu;
  std::condition_variable cv;
  lbz_engine *engines[POOL_MAX] = {};   // fixed array: slots are read without the lock by their owner
  uint32_t sizes[POOL_MAX] = {};        // maximal block size of every context
  bool busy[POOL_MAX] = {};
  int count = 0;
  int max_engines = 0;
  int device = 0;
};
Pool &g_pool = *new Pool;     // leaked on purpose (see Batcher)

// The reference API has no error channel on the encode side: unrecoverable errors
// go through the host program's own failx() (src/main.h:81-88, src/main.c:60-74: it removes the
// partial output file before exiting) when the program linked against this library defines it,
// or through a handler installed with lbz_set_fatal_handler(); only then abort().
[[noreturn]] void die(const char *msg) {
  if (g_fatal) g_fatal(msg);
  if (failx) failx(0, "lbzip2_b200: %s", msg);
  fprintf(stderr, "lbzip2_b200: fatal: %s\n", msg);
  abort();
}

int pool_acquire(uint32_t mbs) {
  std::unique_lock<std::mutex> lk(g_pool.mu);
  if (g_pool.max_engines == 0) {
    const char *s = getenv("LBZIP2_B200_CONTEXTS");
    g_pool.max_engines = s ? atoi(s) : 32;
    if (g_pool.max_engines < 1) g_pool.max_engines = 1;
    if (g_pool.max_engines > POOL_MAX) g_pool.max_engines = POOL_MAX;
    const char *d = getenv("LBZIP2_B200_DEVICE");
    g_pool.device = d ? atoi(d) : 0;
  }
  for (;;) {
    for (int i = 0; i < g_pool.count; i++)
      if (!g_pool.busy[i] && g_pool.engines[i] && g_pool.sizes[i] == mbs) { g_pool.busy[i] = true; return i; }
    int slot = -1;
    lbz_engine *old = nullptr;
    if (g_pool.count < g_pool.max_engines) {
      slot = g_pool.count++;
    } else {
      for (int i = 0; i < g_pool.count; i++)       // recycle an idle context of another block size
        if (!g_pool.busy[i] && g_pool.engines[i]) { slot = i; old = g_pool.engines[i]; break; }
    }
    if (slot >= 0) {
      g_pool.busy[slot] = true;
      g_pool.engines[slot] = nullptr;
      g_pool.sizes[slot] = mbs;
      lk.unlock();
      if (old) lbz_engine_destroy(old);
      lbz_engine *e = engine_create_mbs(g_pool.device, mbs, 1);
      if (!e) die("cannot create a device context (no usable GPU?)");
      lk.lock();
      g_pool.engines[slot] = e;
      return slot;
    }
    g_pool.cv.wait(lk);
  }
}
void pool_release(int i) {
  std::lock_guard<std::mutex> lk(g_pool.mu);
  g_pool.busy[i] = false;
  g_pool.cv.notify_one();
}
}  // namespace

namespace { int batch_limit(); }

#define ENC_MAGIC 0xB2005A42u

// ---- the block-split automaton of collect() on the host ------------------------------------------
// collect() must say how many of the offered bytes the block takes and whether it is full
// (src/encode.c:135-336) when it returns, and the scheduler may offer one block's input in
// several pieces (the -u mode, src/compress.c:160-187).  Asking the device costs a host<->device
// round trip per call, so the split is decided here: `rle_feed` walks the input once and tracks
// only (n' so far, length and byte of the pending run).  Restated from the rules, not the code:
//   * a run of r equal bytes (r <= 259, longer runs restart) adds min(r, 4) bytes, plus a count
//     byte r - 4 once it has 4 or more and ends;
//   * the block is full when n' reaches the capacity after a literal or a count byte, or when the
//     third literal of a run leaves one slot and the run continues (the fourth literal is only
//     written together with room for its count byte, src/encode.c:218,233).
// EMIT = true additionally produces the block bytes, the CRC of the consumed bytes and the
// used-byte map (for blocks whose raw bytes do not fit the device's chunk buffer: long runs in -u
// mode) -- exactly what the reference's collect() does on its calling thread.
struct RleSt { uint32_t nblock; int32_t state; uint32_t ch; };    // state: -1 full, 0 no pending run, else its length (< 259)

template <bool EMIT>
static size_t rle_feed(RleSt &r, const uint8_t *p, size_t avail, uint32_t cap, uint8_t *out, uint32_t *crcp, uint32_t *used) {
  const uint8_t *q = p, *const end = p + avail;
  uint32_t crc = EMIT ? *crcp : 0u;
  auto lit = [&](uint32_t c) {                       // consume one input byte that becomes a block byte
    if (EMIT) { out[r.nblock] = (uint8_t)c; used[c >> 5] |= 1u << (c & 31u); crc = (crc << 8) ^ crc_table[(crc >> 24) ^ c]; }
    r.nblock++;
  };
  auto count_byte = [&](uint32_t v) {
    if (EMIT) { out[r.nblock] = (uint8_t)v; used[v >> 5] |= 1u << (v & 31u); }
    r.nblock++;
  };
  while (q < end && r.state >= 0) {
    if (r.state == 0) {
      if (r.nblock == cap) { r.state = -1; break; }
      r.ch = *q++;
      lit(r.ch);
      r.state = 1;
      if (r.nblock == cap) { r.state = -1; break; }
    } else if (r.state < 3) {
      if (*q != r.ch) { r.state = 0; continue; }
      q++;
      lit(r.ch);
      r.state++;
      if (r.nblock == cap) { r.state = -1; break; }
    } else if (r.state == 3) {
      if (*q != r.ch) { r.state = 0; continue; }
      if (r.nblock + 1u >= cap) { r.state = -1; break; }        // one slot left and the run goes on: cut before the 4th byte
      q++;
      lit(r.ch);
      r.state = 4;
    } else {
      while (q < end && *q == r.ch && r.state < 259) {
        if (EMIT) crc = (crc << 8) ^ crc_table[(crc >> 24) ^ r.ch];
        q++;
        r.state++;
      }
      if (r.state == 259) { count_byte(255u); r.state = 0; }
      else if (q < end) { count_byte((uint32_t)r.state - 4u); r.state = 0; }
    }
  }
  if (r.state == 0 && r.nblock == cap) r.state = -1;             // src/encode.c:162: full is noticed before "input exhausted"
  if (EMIT) *crcp = crc;
  return (size_t)(q - p);
}

// Host bytes behind the handle: the raw bytes of the block until encode(), the packed block
// afterwards (also the internal buffer of transmit(s, NULL), src/encode.c:1177-1182).  A packed
// block never exceeds 5/2 n' + 8 KiB (lbz_common.cuh out_cap); raw input never exceeds mbs.
static size_t enc_staged_cap(unsigned long mbs) { return (size_t)mbs * 5 / 2 + 8192 + 64; }

extern "C" size_t encoder_alloc_size(unsigned long max_block_size) {
  // handle + host staging for the raw bytes of one block.  A block of n' <=
  // mbs RLE1 bytes can cover up to mbs*259/5 raw bytes in theory; the
  // scheduler never offers more than in_granul = mbs bytes per state in the
  // default mode (src/process.c:631, src/compress.c:93-110).  collect()
  // handles longer inputs by stopping at the staging capacity when needed.
  return sizeof(struct encoder_state) + 64 + enc_staged_cap(max_block_size) + 64;
}

extern "C" void encoder_init(struct encoder_state *s, unsigned long max_block_size, unsigned cluster_factor) {
  if (!s || max_block_size == 0 || max_block_size > 900000 || cluster_factor == 0 || cluster_factor > 65535)
    die("encoder_init: bad arguments (src/encode.c:121-123)");
  memset(s, 0, sizeof(*s));
  s->magic = ENC_MAGIC;
  s->max_block_size = (uint32_t)max_block_size;
  s->cluster_factor = cluster_factor;
  s->pool_slot = -1;
  s->staged = reinterpret_cast<uint8_t *>(s) + ((sizeof(struct encoder_state) + 63) / 64) * 64;
  s->staged_cap = (uint32_t)enc_staged_cap(max_block_size);
}

// collect(): the block split is decided on the host (rle_feed above); the raw bytes the block
// takes are staged behind the handle and RLE1 + CRC run on the device with the rest of the block
// (mode 0).  Several calls per state are fine (the -u mode, src/compress.c:160-187).  Only when one
// block takes more raw bytes than the device's chunk buffer holds (max_block_size; long runs packed
// across I/O buffers in -u mode) the block is collected on the host like in the reference -- RLE1
// output, CRC and used-byte map (mode 1) -- and enters the device at the BWT stage.
static inline uint8_t *enc_block_area(struct encoder_state *s) { return s->staged + ((s->max_block_size + 127u) & ~63u); }

extern "C" int collect(struct encoder_state *s, const uint8_t *buf, size_t *buf_sz) {
  if (!s || s->magic != ENC_MAGIC) die("collect: state not initialised");
  if (s->done) die("collect: called after encode()");
  if (s->rle_state < 0) return 1;                       // already full: nothing is consumed
  const size_t avail = *buf_sz;
  if (avail == 0) return 0;
  const uint32_t mbs = s->max_block_size;
  RleSt r{s->nblock, s->rle_state, s->rle_char};
  size_t consumed = 0;
  if (s->mode == 0) {
    const size_t room = (size_t)mbs - s->raw_len;
    const size_t offer = avail < room ? avail : room;
    consumed = rle_feed<false>(r, buf, offer, mbs, nullptr, nullptr, nullptr);
    memcpy(s->staged + s->raw_len, buf, consumed);
    s->raw_len += (uint32_t)consumed;
    if (r.state >= 0 && consumed == offer && offer < avail) {
      // the chunk buffer is exhausted and the block is not full: collect it on the host from here on
      RleSt r2{0u, 0, 0u};
      s->hcrc = 0xFFFFFFFFu;
      memset(s->used, 0, sizeof s->used);
      if (rle_feed<true>(r2, s->staged, s->raw_len, mbs, enc_block_area(s), &s->hcrc, s->used) != s->raw_len ||
          r2.nblock != r.nblock || r2.state != r.state)
        die("collect: internal error: host block split is not reproducible");
      s->mode = 1;
    }
  }
  if (s->mode == 1 && r.state >= 0 && consumed < avail) {
    const size_t c2 = rle_feed<true>(r, buf + consumed, avail - consumed, mbs, enc_block_area(s), &s->hcrc, s->used);
    consumed += c2;
    s->raw_len += (uint32_t)c2;
  }
  s->nblock = r.nblock; s->rle_state = r.state; s->rle_char = r.ch;
  *buf_sz = avail - consumed;
  return r.state < 0;
}

// ---- cross-thread batching of encode() ---------------------------------------
// The scheduler calls encode() from many worker threads, one block each
// (src/compress.c:113).  One block cannot fill a B200, so the calls are pooled:
// every encode() queues its staged raw bytes and sleeps; a dispatcher thread
// collects what arrives within a short window (or until the batch engine is
// full), pushes the whole batch through the kernels at once and wakes the
// callers.  `LBZIP2_B200_BATCH` = chunks per batch (default 32, 0 = no batching:
// every block runs alone on its pooled context).
namespace {
struct BReq { encoder_state *s; int rc; bool done; };
#define BATCH_DISPATCHERS_MAX 4
struct Batcher {
  std::mutex mu;
  std::condition_variable cv_req, cv_done;
  std::deque<BReq *> q;
  bool running = false;
  bool forming = false;                      // one dispatcher at a time gathers a batch; the others run theirs
  int max_batch = -1;
  int dispatchers = 2;
  std::map<uint32_t, lbz_engine *> eng[BATCH_DISPATCHERS_MAX];   // per dispatcher: one batch engine per block size in use
};
// Deliberately leaked: the dispatcher thread sleeps on these condition variables
// for the life of the process, and destroying a condition variable that has a
// waiter (static destruction at exit) blocks forever in glibc.
Batcher &g_batch = *new Batcher;

// Push a batch of collected blocks (one per chunk slot) through the kernels on engine `e` and
// leave every block's bytes in its handle.
int encode_blocks(lbz_engine *e, std::vector<BReq *> &batch) {
  if (cudaSetDevice(e->device) != cudaSuccess) return -1;
  const uint32_t k = (uint32_t)batch.size();
  const size_t mbs = e->g.mbs;
  e->g.nchunks = k;
  e->inject.clear();
  for (uint32_t c = 0; c < k; c++) {
    encoder_state *s = batch[c]->s;
    // a pending run of four or more still owes its count byte (src/encode.c:443-447)
    const bool flush = s->rle_state >= 4;
    if (s->mode == 0) {
      e->h_chunk_len[c] = s->raw_len;
      ENG_CHECK(cudaMemcpyAsync(e->d_in + (size_t)c * mbs, s->staged, s->raw_len, cudaMemcpyHostToDevice, e->st));
    } else {
      e->h_chunk_len[c] = 0;                         // nothing for the RLE1 kernels in this slot
      uint8_t *blk = enc_block_area(s);
      if (flush) {
        const uint32_t v = (uint32_t)s->rle_state - 4u;
        blk[s->nblock++] = (uint8_t)v;
        s->used[v >> 5] |= 1u << (v & 31u);
        s->rle_state = 0;
      }
      lbz_engine::Inject in;
      in.slot = 2 * c; in.bytes = blk;
      memset(&in.meta, 0, sizeof in.meta);
      in.meta.n = s->nblock; in.meta.raw_len = s->raw_len; in.meta.crc = s->hcrc;
      memcpy(in.meta.used, s->used, sizeof s->used);
      e->inject.push_back(in);
    }
  }
  ENG_CHECK(cudaMemcpyAsync(e->d_chunk_len, e->h_chunk_len, k * sizeof(uint32_t), cudaMemcpyHostToDevice, e->st));
  size_t total = 0;
  const int prc = run_pipeline(e, e->d_in, e->d_packed, &total);
  e->inject.clear();
  if (prc) return -1;
  size_t off = 0;
  for (uint32_t c = 0; c < k; c++) {
    encoder_state *s = batch[c]->s;
    const LbzBlockMeta &m0 = e->h_meta[2 * c], &m1 = e->h_meta[2 * c + 1];
    // the device must cut the block exactly where collect() said it would
    const uint32_t want_n = s->nblock + ((s->mode == 0 && s->rle_state >= 4) ? 1u : 0u);
    if (m0.raw_len != s->raw_len || m0.n != want_n || m1.n != 0) {
      batch[c]->rc = -1; off += m0.out_len + (m1.n ? m1.out_len : 0); continue;
    }
    // the raw bytes are no longer needed: the block's bytes take their place in the handle
    ENG_CHECK(cudaMemcpyAsync(s->staged, e->d_packed + off, m0.out_len, cudaMemcpyDeviceToHost, e->st));
    s->out_len = m0.out_len;
    s->crc = m0.crc;
    s->tree_cost = m0.tree_cost;
    off += m0.out_len;
  }
  ENG_CHECK(cudaStreamSynchronize(e->st));
  return 0;
}

int run_batch(int who, uint32_t mbs_key, std::vector<BReq *> &batch) {
  lbz_engine *&e = g_batch.eng[who][mbs_key];
  if (!e) {
    e = engine_create_mbs(g_pool.device, mbs_key, g_batch.max_batch);
    if (!e) return -1;
    e->total_chunks = (uint32_t)g_batch.max_batch;
  }
  return encode_blocks(e, batch);
}

// Dispatcher `who` (LBZIP2_B200_DISPATCHERS of them, default 2, each with its own engines): while one
// batch runs on the device the other dispatcher gathers the next one, so the host work of the
// scheduler's threads (reading, collect(), transmit(), writing) overlaps the kernels.  For that the
// scheduler needs more worker threads than one batch holds: -n 64 with the default batch of 32.
void batch_worker(int who) {
  std::unique_lock<std::mutex> lk(g_batch.mu);
  for (;;) {
    g_batch.cv_req.wait(lk, [] { return !g_batch.q.empty() && !g_batch.forming; });
    g_batch.forming = true;
    // gathering window: wait a little for more blocks unless the batch is already full
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(400);
    while ((int)g_batch.q.size() < g_batch.max_batch &&
           g_batch.cv_req.wait_until(lk, deadline) != std::cv_status::timeout) {}
    const uint32_t level = g_batch.q.front()->s->max_block_size;       // batches are uniform in block size
    std::vector<BReq *> batch;
    for (auto it = g_batch.q.begin(); it != g_batch.q.end() && (int)batch.size() < g_batch.max_batch;) {
      if ((*it)->s->max_block_size == level) { batch.push_back(*it); it = g_batch.q.erase(it); }
      else ++it;
    }
    g_batch.forming = false;
    g_batch.cv_req.notify_all();                     // another dispatcher may start gathering
    lk.unlock();
    const int rc = run_batch(who, level, batch);
    lk.lock();
    for (BReq *r : batch) { if (rc) r->rc = rc; r->done = true; }
    g_batch.cv_done.notify_all();
  }
}

int batch_limit() {
  std::lock_guard<std::mutex> lk(g_batch.mu);
  if (g_batch.max_batch < 0) {
    const char *ev = getenv("LBZIP2_B200_BATCH");
    g_batch.max_batch = ev ? atoi(ev) : 32;
    if (g_batch.max_batch < 0) g_batch.max_batch = 0;
    if (g_batch.max_batch > 1024) g_batch.max_batch = 1024;
    const char *dv = getenv("LBZIP2_B200_DISPATCHERS");
    g_batch.dispatchers = dv ? atoi(dv) : 2;
    if (g_batch.dispatchers < 1) g_batch.dispatchers = 1;
    if (g_batch.dispatchers > BATCH_DISPATCHERS_MAX) g_batch.dispatchers = BATCH_DISPATCHERS_MAX;
  }
  return g_batch.max_batch;
}
}  // namespace

extern "C" size_t encode(struct encoder_state *s, uint32_t *crc) {
  if (!s || s->magic != ENC_MAGIC) die("encode: state not initialised");
  if (s->raw_len == 0) die("encode: empty block (src/encode.c:448)");
  if (batch_limit() > 0) {
    BReq r{s, 0, false};
    {
      std::unique_lock<std::mutex> lk(g_batch.mu);
      if (!g_batch.running) {
        g_batch.running = true;
        for (int w = 0; w < g_batch.dispatchers; w++) std::thread(batch_worker, w).detach();
      }
      g_batch.q.push_back(&r);
      g_batch.cv_req.notify_all();
      g_batch.cv_done.wait(lk, [&] { return r.done; });
    }
    if (r.rc) die("encode: batched kernel pipeline failed");
    s->done = 2;                             // block bytes live in the handle
    if (crc) *crc = s->crc;
    return s->out_len;
  }
  // no batching: the block runs alone on a pooled single-chunk context
  const int slot = pool_acquire(s->max_block_size);
  BReq r{s, 0, false};
  std::vector<BReq *> one{&r};
  const int rc = encode_blocks(g_pool.engines[slot], one);
  pool_release(slot);
  if (rc || r.rc) die("encode: kernel pipeline failed");
  s->done = 2;                               // block bytes live in the handle
  if (crc) *crc = s->crc;
  return s->out_len;
}

extern "C" unsigned generate_prefix_code(struct encoder_state *s) {
  if (!s || s->magic != ENC_MAGIC) die("generate_prefix_code: state not initialised");
  if (!s->done) die("generate_prefix_code: only valid after encode() in this build (the block is coded on the device as a whole)");
  return s->tree_cost;
}

extern "C" void *transmit(struct encoder_state *s, void *buf) {
  if (!s || s->magic != ENC_MAGIC || !s->done) die("transmit: encode() has not run");
  const size_t bytes = ((size_t)s->out_len + 3) / 4 * 4;
  if (bytes > s->staged_cap) die("transmit: internal size error");
  // no external buffer: the block is handed out in the handle's own buffer (src/encode.c:1177-1182)
  memset(s->staged + s->out_len, 0, bytes - s->out_len);
  if (!buf) return s->staged;
  memcpy(buf, s->staged, bytes);
  return buf;
}

extern "C" int32_t divbwt(uint8_t *T, int32_t *SA, int32_t *bucket, int32_t n) {
  (void)bucket;
  if (n <= 0 || n > 900000) die("divbwt: bad length");
  const int slot = pool_acquire(900000u);
  lbz_engine *e = g_pool.engines[slot];
  // one chunk, one block: inject the text and the block record, run the BWT stage
  e->g.nchunks = 1;
  LbzBlockMeta m;
  memset(&m, 0, sizeof m);
  m.n = (uint32_t)n;
  LbzBlockMeta z;
  memset(&z, 0, sizeof z);
  if (lbz_dbg_write(e, LBZ_AR_TEXT, 0, T, (size_t)n) || lbz_dbg_write(e, LBZ_AR_META, 0, &m, sizeof m) ||
      lbz_dbg_write(e, LBZ_AR_META, 1, &z, sizeof z) || lbz_dbg_run(e, LBZ_ST_BWT))
    die("divbwt: device error");
  std::vector<uint8_t> bw((size_t)n);
  if (lbz_dbg_read(e, LBZ_AR_BWT, 0, bw.data(), (size_t)n) || lbz_dbg_read(e, LBZ_AR_META, 0, &m, sizeof m))
    die("divbwt: readback failed");
  for (int32_t i = 0; i < n; i++) SA[i] = bw[i];
  pool_release(slot);
  return (int32_t)m.bwt_idx;
}
