// unbz.cu -- CUDA backend + C ABI of the batch decompressor (include/lbzip2_b200.h section 4).
// Device code: unbz_kernels.cuh.  Host orchestration and the stream walk: unbz_engine.inc.
// There is no CPU path: creation fails loudly without a usable GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/lbzip2_b200.h"
#include "unbz_kernels.cuh"

#define UB_NSTAGE 7
#define UB_NEVENT (UB_NSTAGE + 1)

struct UbBackend {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev[UB_NEVENT];
};

struct lbz_decoder;
static int ub_cuda_fail(cudaError_t e, const char *what, int line) {
  fprintf(stderr, "lbzip2_b200: CUDA error %s in %s (unbz.cu:%d): %s\n", cudaGetErrorName(e), what, line,
          cudaGetErrorString(e));
  return -1;
}
#define UB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ub_cuda_fail(e_, #x, __LINE__); } while (0)

static inline cudaStream_t ub_stream(lbz_decoder *d);
static inline UbBackend *ub_backend(lbz_decoder *d);
static inline void ub_count_launch(lbz_decoder *d);

static int ub_dev_alloc(lbz_decoder *, void **p, size_t bytes) { UB_CUDA(cudaMalloc(p, bytes ? bytes : 16)); return 0; }
static void ub_dev_free(lbz_decoder *, void *p) { cudaFree(p); }
static int ub_h2d(lbz_decoder *d, void *dst, const void *src, size_t bytes) {
  UB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ub_stream(d)));
  return 0;
}
static int ub_d2h(lbz_decoder *d, void *dst, const void *src, size_t bytes) {
  UB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ub_stream(d)));
  return 0;
}
static int ub_dev_zero(lbz_decoder *d, void *p, size_t bytes) { UB_CUDA(cudaMemsetAsync(p, 0, bytes, ub_stream(d))); return 0; }
static int ub_dev_fill32(lbz_decoder *d, uint32_t *p, uint32_t v, size_t count) {
  // every use fills with a value whose four bytes are equal
  UB_CUDA(cudaMemsetAsync(p, (int)(v & 0xFFu), count * 4, ub_stream(d)));
  return 0;
}
static int ub_sync(lbz_decoder *d) { UB_CUDA(cudaStreamSynchronize(ub_stream(d))); return 0; }
static void ub_mark(lbz_decoder *d, int i) {
  UbBackend *b = ub_backend(d);
  if (i < 0) { for (int k = 0; k < UB_NEVENT; k++) cudaEventRecord(b->ev[k], b->stream); return; }
  cudaEventRecord(b->ev[i + 1], b->stream);
}

#define UB_LAUNCH(d, kern, nthreads, cta, ...)                                                      \
  do {                                                                                              \
    uint64_t nt_ = (nthreads);                                                                      \
    if (nt_) {                                                                                      \
      unsigned grid_ = (unsigned)((nt_ + (cta) - 1) / (cta));                                       \
      kern<<<grid_, (cta), 0, ub_stream(d)>>>(__VA_ARGS__);                                         \
      cudaError_t le_ = cudaGetLastError();                                                         \
      if (le_ != cudaSuccess) return ub_cuda_fail(le_, #kern, __LINE__);                            \
      ub_count_launch(d);                                                                           \
    }                                                                                               \
  } while (0)

// Same, with dynamic shared memory the kernel does not touch: a way to bound how many CTAs of a
// latency-bound one-thread-per-CTA kernel share an SM's L1 (LBZ_CHAIN_SMEM_KB, 0 = no bound).
#define UB_LAUNCH_SMEM(d, kern, nthreads, cta, smem, ...)                                           \
  do {                                                                                              \
    uint64_t nt_ = (nthreads);                                                                      \
    if (nt_) {                                                                                      \
      unsigned grid_ = (unsigned)((nt_ + (cta) - 1) / (cta));                                       \
      size_t sm_ = (smem);                                                                          \
      if (sm_ > 48u * 1024u) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_); \
      kern<<<grid_, (cta), sm_, ub_stream(d)>>>(__VA_ARGS__);                                       \
      cudaError_t le_ = cudaGetLastError();                                                         \
      if (le_ != cudaSuccess) return ub_cuda_fail(le_, #kern, __LINE__);                            \
      ub_count_launch(d);                                                                           \
    }                                                                                               \
  } while (0)
static size_t ub_chain_smem() {
  static long kb = -1;
  if (kb < 0) { const char *s = getenv("LBZ_CHAIN_SMEM_KB"); kb = s ? atol(s) : 0; if (kb < 0 || kb > 200) kb = 0; }
  return (size_t)kb * 1024u;
}

static void ub_timers_collect(lbz_decoder *d);

#include "unbz_engine.inc"

static_assert(sizeof(lbz_dblock) == sizeof(UbBlock), "lbz_dblock mirrors UbBlock");
static_assert((int)LBZ_ERR_EOF == (int)UB_ERR_EOF, "status numbering");

static inline cudaStream_t ub_stream(lbz_decoder *d) { return d->be.stream; }
static inline UbBackend *ub_backend(lbz_decoder *d) { return &d->be; }
static inline void ub_count_launch(lbz_decoder *d) { d->launches++; }

static void ub_timers_collect(lbz_decoder *d) {
  UbBackend *b = &d->be;
  if (cudaEventSynchronize(b->ev[UB_NEVENT - 1]) != cudaSuccess) return;
  float ms = 0;
  for (int k = 0; k < UB_NSTAGE; k++) {
    d->stage_ms[k] = cudaEventElapsedTime(&ms, b->ev[k], b->ev[k + 1]) == cudaSuccess ? ms : 0.0;
  }
  d->last_ms = cudaEventElapsedTime(&ms, b->ev[0], b->ev[UB_NEVENT - 1]) == cudaSuccess ? ms : 0.0;
}

extern "C" lbz_decoder *lbz_decoder_create(int device, int max_blocks, size_t in_cap, size_t out_cap) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    fprintf(stderr, "lbzip2_b200: no usable CUDA device %d (%s); this library has no CPU path\n", device,
            e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return nullptr;
  }
  if (max_blocks < 1) max_blocks = 1;
  const size_t min_out = (size_t)UB_MAXBLK / 5 * 259 + 4096;      // the largest single block (a "zip bomb" block)
  if (out_cap < min_out) out_cap = min_out;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  lbz_decoder *d = new lbz_decoder();
  d->be.device = device;
  d->max_blocks = (uint32_t)max_blocks;
  d->in_cap = in_cap < 64 ? 64 : in_cap;
  d->out_cap = out_cap;
  d->launches = 0;
  d->loaded_n = ~0ull;
  bool ok = cudaStreamCreateWithFlags(&d->be.stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int k = 0; ok && k < UB_NEVENT; k++) ok = cudaEventCreate(&d->be.ev[k]) == cudaSuccess;
  ok = ok && cudaHostAlloc((void **)&d->h_blk, sizeof(UbBlock) * d->max_blocks, cudaHostAllocDefault) == cudaSuccess;
  if (!ok || ub_decoder_alloc(d) != 0) {
    fprintf(stderr, "lbzip2_b200: decoder set-up failed (%d blocks, %zu in, %zu out)\n", max_blocks, in_cap, out_cap);
    lbz_decoder_destroy(d);
    return nullptr;
  }
  return d;
}

extern "C" void lbz_decoder_destroy(lbz_decoder *d) {
  if (!d) return;
  cudaSetDevice(d->be.device);
  if (d->be.stream) cudaStreamSynchronize(d->be.stream);
  ub_decoder_release(d);
  if (d->h_blk) cudaFreeHost(d->h_blk);
  for (int k = 0; k < UB_NEVENT; k++) if (d->be.ev[k]) cudaEventDestroy(d->be.ev[k]);
  if (d->be.stream) cudaStreamDestroy(d->be.stream);
  delete d;
}

extern "C" int lbz_decoder_load(lbz_decoder *d, const uint8_t *in, size_t n) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  if (ub_upload(d, in, n) != 0) return -1;
  return ub_sync(d);
}

extern "C" int lbz_decompress_ex(lbz_decoder *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                 size_t *out_len, lbz_dstream_info *info, unsigned flags) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  size_t dummy = 0;
  return ub_decompress(d, in, n, out, out_cap, out_len ? out_len : &dummy, info, flags);
}

extern "C" int lbz_decompress_stream(lbz_decoder *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                     size_t *out_len, lbz_dstream_info *info) {
  return lbz_decompress_ex(d, in, n, out, out_cap, out_len, info, 0);
}

extern "C" int lbz_decoder_open(lbz_decoder *d, const uint8_t *in, size_t n, unsigned flags) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  return ub_open(d, in, n, flags);
}

extern "C" int lbz_decoder_next(lbz_decoder *d, uint8_t *out, size_t out_cap, size_t *out_len, lbz_dstream_info *info) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  size_t dummy = 0;
  return ub_next(d, out, out_cap, out_len ? out_len : &dummy, info);
}

extern "C" int lbz_decoder_decode_at(lbz_decoder *d, const uint8_t *in, size_t n, const uint64_t *magic_bits, uint32_t count,
                                     lbz_dblock *table, unsigned flags) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  return ub_decode_at(d, in, n, magic_bits, count, table, flags);
}

extern "C" int lbz_decoder_emit_at(lbz_decoder *d, const uint64_t *out_off, uint32_t count, uint8_t *out, size_t out_cap,
                                   size_t *out_len, uint32_t *crc) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  size_t dummy = 0;
  return ub_emit_at(d, out_off, count, out, out_cap, out_len ? out_len : &dummy, crc);
}

extern "C" int lbz_walk_table(const uint8_t *in, size_t n, const lbz_dblock *table, size_t count, uint32_t *chain,
                              uint32_t *chain_crc, size_t *nchain, lbz_dstream_info *info) {
  return ub_walk_table(in, n, table, count, chain, chain_crc, nchain, info);
}

extern "C" long lbz_scan_blocks(lbz_decoder *d, const uint8_t *in, size_t n, uint64_t *bit_positions, size_t cap) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  if (ub_upload(d, in, n) != 0) return -1;
  if (ub_scan(d, (n + 3) / 4) != 0) return -1;
  for (size_t i = 0; i < d->hits.size() && i < cap; i++) bit_positions[i] = d->hits[i];
  return (long)d->hits.size();
}

extern "C" int lbz_decoder_read(lbz_decoder *d, int array, uint64_t slot, void *dst, size_t bytes) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  const void *src = nullptr;
  switch (array) {
    case LBZ_DA_BLOCK:
      if (slot >= d->max_blocks || bytes > sizeof(UbBlock)) return -1;
      src = d->d_blk + slot;
      break;
    case LBZ_DA_BWT:
      if (slot >= d->max_blocks || bytes > UB_STRIDE) return -1;
      src = d->d_bwt + slot * UB_STRIDE;
      break;
    case LBZ_DA_TEXT:
      if (slot >= d->max_blocks || bytes > UB_STRIDE) return -1;
      src = d->d_txt + slot * UB_STRIDE;
      break;
    case LBZ_DA_OUT:
      if (slot + bytes > d->out_cap) return -1;
      src = d->d_out + slot;
      break;
    default:
      return -1;
  }
  if (ub_d2h(d, dst, src, bytes) != 0) return -1;
  return ub_sync(d);
}

extern "C" uint32_t lbz_decoder_last_wave_blocks(const lbz_decoder *d) { return d->last_wave_blocks; }
extern "C" uint64_t lbz_decoder_launches(const lbz_decoder *d) { return d->launches; }
extern "C" size_t lbz_decoder_device_bytes(const lbz_decoder *d) { return d->device_bytes; }
extern "C" double lbz_decoder_last_ms(const lbz_decoder *d) { return d->last_ms; }
extern "C" void lbz_decoder_stage_ms(const lbz_decoder *d, double *out7) {
  for (int k = 0; k < UB_NSTAGE; k++) out7[k] = d->stage_ms[k];
}

extern "C" const char *lbz_strerror(int status) {
  // the reference's texts: src/expand.c:70-94 (err2str), src/process.c:680 (first header)
  static const char *const text[] = {
    "not a valid bzip2 file", "bad block header magic", "empty source alphabet", "bad number of trees",
    "no coding groups", "invalid selector", "invalid delta code", "invalid prefix code",
    "incomplete prefix code", "empty block", "unterminated block", "missing run length",
    "block CRC mismatch", "stream CRC mismatch", "block overflow", "primary index too large",
    "unexpected end of file"};
  if (status == LBZ_OK) return "ok";
  if (status >= LBZ_ERR_MAGIC && status <= LBZ_ERR_EOF) return text[status - LBZ_ERR_MAGIC];
  if (status == LBZ_ERR_OUTCAP) return "output buffer too small";
  return "internal error";
}
