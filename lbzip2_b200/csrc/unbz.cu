// unbz.cu -- CUDA backend + C ABI of the batch decompressor (include/lbzip2_b200.h section 4).
// Device code: unbz_kernels.cuh.  Host orchestration and the stream walk: unbz_engine.inc.
// There is no CPU path: creation fails loudly without a usable GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

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
// dst < src inside one allocation: pieces no longer than the distance moved never overlap, and the
// stream keeps them in order
static int ub_dev_move(lbz_decoder *d, void *dst, const void *src, size_t bytes) {
  const size_t gap = (size_t)((const uint8_t *)src - (uint8_t *)dst);
  for (size_t done = 0; done < bytes; done += gap) {
    const size_t c = bytes - done < gap ? bytes - done : gap;
    UB_CUDA(cudaMemcpyAsync((uint8_t *)dst + done, (const uint8_t *)src + done, c, cudaMemcpyDeviceToDevice, ub_stream(d)));
  }
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

// ---------------------------------------------------------------------------------------------
// k_ub_chain2 -- the walk over the code lengths of a block (src/decode.c:605-791: where code k
// starts depends on the lengths of codes 0..k-1), device-only rewrite of k_ub_chain for latency:
//   * one CTA of four warps per block; the CTA copies the block's multi-code byte tables (4 KB
//     per tree) into shared memory, then ONE warp walks.  Which warp is decided per SM (an atomic
//     counter per %smid), so that two blocks that share an SM walk on different sub-partitions
//     -- single-warp CTAs all sit on sub-partition 0 and two of them ran at half speed each.
//   * the bit window is a 64-bit register refilled one word ahead; a table step is: shift, one
//     shared-memory byte load, and/shift/add -- about a third of the instructions of k_ub_chain;
//     the values are kept out of the uniform datapath (every lane computes the same walk).
//   * the selector list is a register of 4-bit fields (no local-memory array).
// Results are those of k_ub_chain (same tables, same rules); tests/test_gpu_unbz.py checks both.
#define CH_THREADS 128u
#define CH_DBG_BLOCKS 1024u
__device__ __forceinline__ uint32_t ch_bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint32_t ch_lds_u8(uint32_t addr) {       // addr: 32-bit shared-window address
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ch_lds_u16(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

#define CH_SMEM_BYTES (6u * UB_WSIZE + 6u * UB_WSIZE * 2u)      // byte tables + one-code tables of up to six trees: 72 KB
__global__ void __launch_bounds__(CH_THREADS)
k_ub_chain2(const uint32_t *__restrict__ words, uint64_t nwords, UbBlock *blk, uint32_t nblk,
            const uint8_t *__restrict__ sel_all, const UbTreeG *__restrict__ tree_all,
            const uint16_t *__restrict__ l1_all, const uint32_t *__restrict__ ml_all, const uint8_t *__restrict__ mq_all,
            uint64_t *__restrict__ gpos_all, uint8_t *__restrict__ gtree_all, uint32_t *sm_slots) {
  extern __shared__ __align__(16) uint8_t ch_smem[];
  uint8_t *smq = ch_smem;                                          // [6][4096] multi-code byte entries
  uint16_t *sl1 = reinterpret_cast<uint16_t *>(ch_smem + 6u * UB_WSIZE);   // [6][4096] (symbol << 5 | length) of the first code
  __shared__ uint32_t s_slot;
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  if (b >= nblk) return;
  UbBlock &B = blk[b];
  if (B.status != UB_PENDING) return;
  const uint32_t ntrees = B.num_trees > 6u ? 6u : B.num_trees;
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(mq_all + (size_t)b * 6u * UB_WSIZE);
    uint4 *dst = reinterpret_cast<uint4 *>(smq);
    for (uint32_t i = tid; i < ntrees * (UB_WSIZE / 16u); i += CH_THREADS) dst[i] = src[i];
    const uint4 *src1 = reinterpret_cast<const uint4 *>(l1_all + (size_t)b * 6u * UB_WSIZE);
    uint4 *dst1 = reinterpret_cast<uint4 *>(sl1);
    for (uint32_t i = tid; i < ntrees * (UB_WSIZE / 8u); i += CH_THREADS) dst1[i] = src1[i];
  }
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (tid == 0) s_slot = atomicAdd(&sm_slots[smid & 255u], 1u);
  __syncthreads();
  if ((tid >> 5) != (s_slot & 3u)) return;
  const long long t_start = clock64();
  // every lane of the walking warp computes the same walk; `zero` is 0 but opaque to the
  // compiler, which keeps the chain in vector registers
  uint32_t zero;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(zero));
  const bool writer = zero == 0u;
  zero >>= 5;

  const uint8_t *sel = sel_all + (size_t)b * UB_SELCAP;
  uint64_t *gpos = gpos_all + (size_t)b * (UB_MAXGRP + 1u);
  uint8_t *gtree = gtree_all + (size_t)b * (UB_MAXGRP + 1u);
  const uint32_t eob = B.alpha_size - 1u;
  const uint32_t smq_base = (uint32_t)__cvta_generic_to_shared(smq);
  const uint32_t sl1_base = (uint32_t)__cvta_generic_to_shared(sl1);
  const uint32_t nsel = B.num_selectors > UB_MAXGRP ? UB_MAXGRP : B.num_selectors;   // src/decode.c:631-632
  uint32_t slp = 0;                                  // selector list, 4 bits per entry: tree number or error code
  for (uint32_t t = 0; t < 6u; t++) slp |= (t < B.num_trees ? ((uint32_t)tree_all[(size_t)b * 6u + t].status & 15u) : 0u) << (4u * t);
  uint64_t pos = B.sym_bit + zero;
  uint32_t status = UB_ERR_UNTERM, nsym = 0, g = 0;
  UbBits br;
  br.words = words; br.nwords = nwords;
  // window of the lean reader, carried from group to group: v = the next `avail` bits of the block,
  // left-justified (at least 32 of them whenever a table is consulted), wi = index of the word that
  // is appended next, nxt = that word (raw).  Position = 32 wi - avail.
  // (the next word is addressed by a pointer that lives in registers: indexing the parameter made the
  // compiler re-load the base from the constant bank, ~40 cycles in front of every refill test)
  uint64_t v = 0;
  uint32_t avail = 0, nxt = 0;
  const uint32_t *wp = words;                        // address of the word that is appended next (= nxt)
#define wi ((uint32_t)(wp - words))
  const uint32_t pf_lim = nwords > 64u ? (uint32_t)(nwords - 64u < 0xFFFFFFFFull ? nwords - 64u : 0xFFFFFFFFull) : 0u;   // prefetch stays inside the input
  // wi = (pos >> 5) + 2 or + 1 (+2 right after a refill, the window then holds 33..64 bits): wi + 34 <= nwords implies the lean test
  const uint32_t lean_lim = nwords > 34u ? (uint32_t)(nwords - 34u < 0xFFFFFFFFull ? nwords - 34u : 0xFFFFFFFFull) : 0u;
  bool window = false;                               // the registers above describe `pos`
#define CH_REFILL()                                                                              \
  do {                                                                                           \
    v |= (uint64_t)ch_bswap(nxt) << (32u - avail);                                               \
    avail += 32u;                                                                                \
    wp++;                                                                                        \
    nxt = *wp;                                                                                   \
  } while (0)

  for (; g < nsel; g++) {
    const uint32_t r4 = 4u * min((uint32_t)sel[g], 7u);      // the header kernel admits only indices below num_trees
    const uint32_t t = (slp >> r4) & 15u;
    if (t >= 6u) { status = t; break; }              // a bad tree is selected (src/decode.c:640-642)
    {                                                // move entry r to the front
      const uint32_t below = slp & ((1u << r4) - 1u);
      const uint32_t above = (r4 >= 28u) ? 0u : (slp >> (r4 + 4u)) << (r4 + 4u);
      slp = above | (below << 4) | t;
    }
    if (window) pos = ((uint64_t)wi << 5) - avail;
    if (writer) { gpos[g] = pos; gtree[g] = (uint8_t)t; }
    const UbTreeG &T = tree_all[(size_t)b * 6u + t];
    bool done = false;

    // lean window reader while a whole group (50 codes of at most 20 bits = 32 words) lies inside the input
    if (window ? (wi < lean_lim) : ((pos >> 5) + 36u <= nwords)) {
      if (!window) {
        const uint32_t w0 = (uint32_t)(pos >> 5), bp = (uint32_t)(pos & 31u);
        v = (((uint64_t)ch_bswap(words[w0]) << 32) | ch_bswap(words[w0 + 1])) << bp;
        avail = 64u - bp;
        wp = words + w0 + 2u;
        nxt = *wp;
        window = true;
      }
      if (wi < pf_lim) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + 64));          // the stream two cache lines ahead
      const uint32_t tb = smq_base + t * UB_WSIZE;
      const uint32_t tl = sl1_base + t * (UB_WSIZE * 2u);
      // Table steps while even four-code entries cannot overshoot the group.  Only the bit window is
      // carried from step to step (shift, one shared-memory byte, shift); the code count rides along.
      // Two steps per refill test: a step consumes at most 12 bits and 32 are guaranteed.  A flagged
      // entry (a code longer than the window -- every other group of a text block has one -- or the
      // end-of-block symbol) is rare per step: it is decoded as ONE code from the one-code table (or
      // the canonical tables) and the walk goes on.
      uint32_t cnt = 0;                                // codes of this group walked so far
#define CH_ONE_CODE()                                                                            \
      do {                                                                                       \
        if (avail < 32u) CH_REFILL();                                                            \
        const uint32_t x_ = ch_lds_u16(tl + 2u * (uint32_t)(v >> (64u - UB_WBITS)));             \
        uint32_t len_, s_;                                                                       \
        if (x_) { s_ = x_ >> 5; len_ = x_ & 31u; }                                               \
        else s_ = ub_canon_decode(T, (uint32_t)(v >> 44), &len_);                                \
        v <<= len_;                                                                              \
        avail -= len_;                                                                           \
        cnt += 1u;                                                                               \
        if (s_ == eob) done = true;                                                              \
      } while (0)
      while (cnt <= 42u && !done) {
        if (avail < 32u) CH_REFILL();
        const uint64_t vb = v;
        const uint32_t q1 = ch_lds_u8(tb + (uint32_t)(v >> (64u - UB_WBITS)));
        v <<= (q1 & 15u);
        const uint32_t q2 = ch_lds_u8(tb + (uint32_t)(v >> (64u - UB_WBITS)));
        v <<= (q2 & 15u);
        if (!((q1 | q2) & 0x80u)) {
          avail -= (q1 & 15u) + (q2 & 15u);
          cnt += (q1 >> 4) + (q2 >> 4);
        } else {
          v = vb;                                      // take the first entry if it is plain, then one code
          if (!(q1 & 0x80u)) { v <<= (q1 & 15u); avail -= q1 & 15u; cnt += q1 >> 4; }
          CH_ONE_CODE();
        }
      }
      while (cnt <= 46u && !done) {
        if (avail < 32u) CH_REFILL();
        const uint32_t q = ch_lds_u8(tb + (uint32_t)(v >> (64u - UB_WBITS)));
        if (!(q & 0x80u)) { v <<= (q & 15u); avail -= q & 15u; cnt += q >> 4; }
        else CH_ONE_CODE();
      }
      while (cnt < 50u && !done) CH_ONE_CODE();        // the last codes of the group, one at a time
#undef CH_ONE_CODE
      nsym += cnt;
      if (done) pos = ((uint64_t)wi << 5) - avail;
    } else {
      window = false;
      const uint16_t *l1 = l1_all + ((size_t)b * 6u + t) * UB_WSIZE;
      bool eof = !ub_bits_seek(br, pos);
      for (uint32_t j = 0; j < 50u && !eof && !done; j++) {
        if (!ub_bits_need(br)) { eof = true; break; }  // NEED(), src/decode.c:387-407
        uint32_t c20 = ub_bits_peek(br, 20);
        uint32_t x = l1[c20 >> (20u - UB_WBITS)];
        uint32_t s, k;
        if (x) { s = x >> 5; k = x & 31u; } else s = ub_canon_decode(T, c20, &k);
        ub_bits_dump(br, k);
        nsym++;
        if (s == eob) done = true;
      }
      if (!eof || done) pos = ub_bits_pos(br);
      if (eof && !done) { pos = ub_bits_pos(br); status = UB_ERR_EOF; g++; break; }
    }
    if (done) { status = UB_OK; g++; break; }
  }
  if (window && status != UB_OK) pos = ((uint64_t)wi << 5) - avail;              // ran out of selectors inside the window reader
#undef CH_REFILL
#undef wi
  if (writer) {
    B.nsym = nsym;
    B.ngrp = g;
    B.end_bit = pos;
    B.status = status;
    if (b < CH_DBG_BLOCKS) {                         // per-block walk statistics (LBZ_CHAIN_DEBUG=1 prints them)
      uint32_t *dbg = sm_slots + 256u + 4u * b;
      dbg[0] = smid; dbg[1] = s_slot; dbg[2] = (uint32_t)((clock64() - t_start) >> 10); dbg[3] = g;
    }
  }
}

static uint32_t *ub_sm_slots(int device) {
  static uint32_t *slots[64] = {};
  if (device < 0 || device >= 64) return nullptr;
  if (!slots[device] && cudaMalloc((void **)&slots[device], (256 + 4 * CH_DBG_BLOCKS) * sizeof(uint32_t)) != cudaSuccess) return nullptr;
  return slots[device];
}
static void ub_chain_debug_print(uint32_t *d_slots, uint32_t nblk, cudaStream_t st) {
  static int on = -1;
  if (on < 0) on = getenv("LBZ_CHAIN_DEBUG") != nullptr;
  if (!on) return;
  const uint32_t n = nblk < CH_DBG_BLOCKS ? nblk : CH_DBG_BLOCKS;
  static uint32_t h[256 + 4 * CH_DBG_BLOCKS];
  cudaStreamSynchronize(st);
  if (cudaMemcpy(h, d_slots, (256 + 4 * n) * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  // blocks alone on their SM vs blocks that shared it: kilo-cycles per group of 50 codes
  double alone = 0, shared = 0; uint32_t na = 0, ns = 0, mx = 0;
  for (uint32_t b = 0; b < n; b++) {
    const uint32_t *e = h + 256 + 4 * b;
    if (!e[3]) continue;
    const double per = (double)e[2] * 1024.0 / (double)e[3];
    if (h[e[0] & 255u] > 1) { shared += per; ns++; } else { alone += per; na++; }
    if (e[2] > mx) mx = e[2];
  }
  fprintf(stderr, "lbzip2_b200: k_ub_chain2: %u blocks alone on an SM: %.0f cycles per group; %u blocks sharing an SM: %.0f cycles per group; "
          "longest walk %.2f M cycles\n", na, na ? alone / na : 0.0, ns, ns ? shared / ns : 0.0, mx * 1024.0 / 1e6);
}
static int ub_chain_version() {
  static int v = -1;
  if (v < 0) { const char *s = getenv("LBZ_CHAIN_VER"); v = s ? atoi(s) : 2; if (v != 1) v = 2; }
  return v;
}
#define UB_CHAIN_LAUNCH(d, nwords, nblk)                                                             \
  do {                                                                                              \
    uint32_t *slots_ = ub_chain_version() == 2 ? ub_sm_slots(ub_backend(d)->device) : nullptr;      \
    if (slots_) {                                                                                   \
      UB_CUDA(cudaMemsetAsync(slots_, 0, 256 * sizeof(uint32_t), ub_stream(d)));                    \
      static bool ch_attr_ = false;                                                                 \
      if (!ch_attr_) { cudaFuncSetAttribute(k_ub_chain2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM_BYTES); ch_attr_ = true; } \
      k_ub_chain2<<<(nblk), CH_THREADS, CH_SMEM_BYTES, ub_stream(d)>>>((d)->d_words, (nwords), (d)->d_blk, (nblk), (d)->d_sel, (d)->d_tree, \
                                                            (d)->d_l1, (d)->d_ml, (d)->d_mq, (d)->d_gpos, (d)->d_gtree, slots_);  \
      cudaError_t le_ = cudaGetLastError();                                                         \
      if (le_ != cudaSuccess) return ub_cuda_fail(le_, "k_ub_chain2", __LINE__);                    \
      ub_count_launch(d);                                                                           \
      ub_chain_debug_print(slots_, (nblk), ub_stream(d));                                           \
    } else {                                                                                        \
      UB_LAUNCH_SMEM(d, k_ub_chain, (uint64_t)(nblk) * 32u, 32u, ub_chain_smem(), (d)->d_words, (nwords), (d)->d_blk, (nblk), \
                     (d)->d_sel, (d)->d_tree, (d)->d_l1, (d)->d_ml, (d)->d_mq, (d)->d_gpos, (d)->d_gtree);           \
    }                                                                                               \
  } while (0)

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

extern "C" int lbz_decoder_open_stream(lbz_decoder *d, unsigned flags) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  return ub_open_stream(d, flags);
}

extern "C" int lbz_decoder_feed(lbz_decoder *d, const uint8_t *in, size_t n, int eof, size_t *taken) {
  if (cudaSetDevice(d->be.device) != cudaSuccess) return -1;
  size_t dummy = 0;
  return ub_feed(d, in, n, eof, taken ? taken : &dummy);
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
      if (bytes > d->out_cap || slot > d->out_cap - bytes) return -1;
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

// ---------------------------------------------------------------------------------------------
// 5. The reference's per-block decoder API (src/decode.h:72-81): decoder_init / decoder_free /
//    retrieve / decode / emit, so that the UNMODIFIED src/expand.c (reference scheduler, parser
//    and scanner) drives the GPU decoder.  oracle/Makefile links _ref/lbzip2_gpu without
//    src/decode.c for exactly this.
//
// The impedance: retrieve() is fed one 256 KiB I/O buffer at a time
// (src/expand.c:553-565), must return MORE at the end of a buffer and OK with the bit cursor ON
// the block's last bit -- the caller turns the cursor into the stream position the parser
// resumes at and releases the buffers behind it (src/expand.c:305-343, :585-596).  Where a block
// ends is only known after its codes have been decoded, so every call decodes what it has: the
// bits offered so far are kept behind the state, the block is pushed through the kernels on a
// pooled single-block decoder, and
//   * "input exhausted" (and more to come)  -> MORE, everything offered is consumed;
//   * block complete                        -> its bytes and CRC are fetched at once, the cursor is
//     put on the end bit -- which lies in the CURRENT buffer, because a call that had seen the end
//     would have returned OK itself -- and decode() has nothing left to do; emit() hands the
//     bytes out in the caller's 900 000-byte pieces.
// A block that spans k buffers is decoded k times (k <= 2 for text at -9, <= 5 for incompressible
// data); the scheduler's worker threads (-n) keep that many single-block decodes in flight.  The
// throughput path is the batch decompressor (section 4); this is the drop-in boundary.
struct in_blk;
struct bitstream {              // src/decode.h:39-46
  unsigned live;
  uint64_t buff;
  struct in_blk *block;
  const uint32_t *data;
  const uint32_t *limit;
  bool eof;
};
struct retriever_internal_state {
  std::vector<uint8_t> acc;     // [24-byte lead-in with a stand-in block header][the block's bits, byte-aligned to the input words]
  uint64_t start_bit;           // the stand-in header's first bit in acc
  bool started;
  std::vector<uint8_t> out;     // decoded bytes of the block
  size_t out_pos;
  uint32_t crc, rl_state;
};
struct decoder_state {          // src/decode.h:49-66 (public: the caller reads block_size and crc, src/expand.c:666,682)
  struct retriever_internal_state *internal_state;
  bool rand;
  unsigned bwt_idx;
  unsigned block_size;
  uint32_t crc;
  uint32_t ftab[256];
  uint32_t *tt;
  int rle_state;
  uint32_t rle_crc, rle_index, rle_avail;
  uint8_t rle_char, rle_prev;
};

namespace {
struct DecPool {
  std::mutex mu;
  std::vector<lbz_decoder *> idle;
};
DecPool &g_decpool = *new DecPool;     // leaked on purpose: worker threads may outlive static destruction
const size_t kShimInCap = 4u << 20;    // > the longest block: 900 000 symbols of 20 bits + tables

lbz_decoder *decpool_acquire() {
  {
    std::lock_guard<std::mutex> lk(g_decpool.mu);
    if (!g_decpool.idle.empty()) { lbz_decoder *d = g_decpool.idle.back(); g_decpool.idle.pop_back(); return d; }
  }
  const char *dv = getenv("LBZIP2_B200_DEVICE");
  return lbz_decoder_create(dv ? atoi(dv) : 0, 1, kShimInCap, 0);
}
void decpool_release(lbz_decoder *d) {
  std::lock_guard<std::mutex> lk(g_decpool.mu);
  g_decpool.idle.push_back(d);
}
[[noreturn]] void shim_die(const char *msg) {
  fprintf(stderr, "lbzip2_b200: fatal: %s\n", msg);
  abort();
}
}  // namespace

extern "C" void decoder_init(struct decoder_state *ds) {
  ds->internal_state = new retriever_internal_state();
  ds->internal_state->started = false;
  ds->internal_state->out_pos = 0;
  ds->tt = nullptr;
  ds->block_size = 0;
}

extern "C" void decoder_free(struct decoder_state *ds) {
  delete ds->internal_state;
  ds->internal_state = nullptr;
}

extern "C" int retrieve(struct decoder_state *ds, struct bitstream *bs) {
  retriever_internal_state *rs = ds->internal_state;
  if (!rs) shim_die("retrieve: state not initialised");
  if (!rs->started) {
    // the cursor's pending bits (left-justified in buff, src/decode.c:367-372) come right before the
    // first word; an 80-bit stand-in for the block header the parser has consumed (magic + CRC,
    // src/parse.c:213-243) comes right before them, because the decoder addresses blocks by their magic
    const unsigned w = bs->live;
    if (w > 63u) shim_die("retrieve: bad bit cursor");
    rs->acc.assign(24, 0);
    const uint64_t x = w ? (bs->buff >> (64u - w)) : 0ull;
    for (int k = 0; k < 8; k++) rs->acc[16 + k] = (uint8_t)(x >> (56 - 8 * k));
    const uint64_t body = 192u - w;
    rs->start_bit = body - 80u;
    const uint64_t magic = 0x314159265359ull;
    for (unsigned k = 0; k < 48u; k++)
      if ((magic >> (47u - k)) & 1u) { const uint64_t p = rs->start_bit + k; rs->acc[p >> 3] |= (uint8_t)(0x80u >> (p & 7u)); }
    rs->started = true;
  }
  const uint32_t *const data0 = bs->data;
  const size_t call_base_bits = rs->acc.size() * 8;
  if (bs->limit > bs->data && rs->acc.size() < kShimInCap - (1u << 19))
    rs->acc.insert(rs->acc.end(), reinterpret_cast<const uint8_t *>(bs->data), reinterpret_cast<const uint8_t *>(bs->limit));
  lbz_decoder *dec = decpool_acquire();
  if (!dec) shim_die("retrieve: cannot create a device context (no usable GPU?)");
  lbz_dblock tb;
  uint64_t pos = rs->start_bit;
  if (ub_decode_at(dec, rs->acc.data(), rs->acc.size(), &pos, 1, &tb, 0) != 0) shim_die("retrieve: device error");
  int status = (int)tb.status;
  if (status == UB_OK && tb.rl_state == 4u) { /* reported by emit(), like the reference */ }
  if (status != UB_OK) {
    decpool_release(dec);
    bs->data = bs->limit; bs->live = 0; bs->buff = 0;                  // everything offered is consumed
    if (status == UB_ERR_EOF && !bs->eof) return UB_MORE;              // src/decode.c:387-396
    return status;
  }
  ds->rand = tb.rand != 0;
  ds->bwt_idx = tb.bwt_idx;
  ds->block_size = tb.block_size;
  rs->out.resize((size_t)tb.out_len);
  rs->out_pos = 0;
  const uint64_t off0 = 0;
  size_t olen = 0;
  uint32_t crc = 0;
  if (tb.out_len > dec->out_cap ||
      ub_emit_at(dec, &off0, 1, rs->out.empty() ? nullptr : rs->out.data(), rs->out.size(), &olen, &crc) != 0)
    shim_die("retrieve: device error while writing the block");
  rs->crc = crc;
  rs->rl_state = dec->h_blk[0].rl_state;
  decpool_release(dec);
  // the bit cursor goes on the first bit after the block
  const uint64_t E = tb.end_bit;
  if (E >= call_base_bits) {
    const uint64_t rel = E - call_base_bits;                           // bits of this call's words that belong to the block
    const uint64_t nw = (rel + 31u) / 32u;
    const unsigned live = (unsigned)(nw * 32u - rel);
    bs->data = data0 + nw;
    bs->live = live;
    bs->buff = live ? ((uint64_t)(__builtin_bswap32(data0[nw - 1]) & ((1u << live) - 1u)) << (64u - live)) : 0ull;
  } else {
    // the block ended inside the bits the cursor held when the first call came in
    const uint64_t left = call_base_bits - E;                          // pending bits that stay pending
    if (left > 63u || call_base_bits != 192u) shim_die("retrieve: internal cursor error");
    const uint64_t x = bs->live ? (bs->buff >> (64u - bs->live)) : 0ull;
    bs->live = (unsigned)left;
    bs->buff = left ? (x << (64u - left)) : 0ull;
  }
  return UB_OK;
}

extern "C" void decode(struct decoder_state *ds) { (void)ds; }       // the inverse BWT ran with the rest of the block

extern "C" int emit(struct decoder_state *ds, void *buf, size_t *buf_sz) {
  retriever_internal_state *rs = ds->internal_state;
  const size_t cap = *buf_sz, left = rs->out.size() - rs->out_pos;
  const size_t n = cap < left ? cap : left;
  if (n) memcpy(buf, rs->out.data() + rs->out_pos, n);
  rs->out_pos += n;
  *buf_sz = cap - n;
  if (rs->out_pos < rs->out.size()) return UB_MORE;
  if (rs->rl_state == 4u) return UB_ERR_RUNLEN;                        // src/decode.c:936-942
  ds->crc = rs->crc;
  return UB_OK;
}
