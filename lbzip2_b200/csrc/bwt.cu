// bwt.cu -- Burrows-Wheeler transform of many independent blocks under
// cyclic-rotation order.
//
// Replaces divbwt() (reference src/divbwt.c:1707-1726; contract: last column
// of the sorted cyclic rotations + position of rotation 0).  The reference is
// a serial divsufsort; this is a batched, bandwidth-oriented design:
//
//   1. order rotations by their first BWT_K bytes with BWT_K stable counting-
//      sort passes (LSD radix, 8-bit digits read straight from the text that
//      stays L2-resident; only 32-bit rotation indices move through HBM);
//   2. prefix doubling (Larsson-Sadakane style with group filtering): every
//      round sorts only the rotations that still sit in tied groups by the key
//      (group start, rank of rotation i+h), refines the groups, drops the ones
//      that became singletons, and doubles h.  A block is finished when no
//      tied group is left or when h >= n (then the remaining ties are exactly
//      equal rotations, i.e. the block is periodic);
//   3. gather the last column.
//
// All blocks of a batch are processed by the same launches; the block is the
// sort segment (blockIdx.y), tiles of LBZ_TILE elements are blockIdx.x.
#include "lbz_common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 16               // SORT_THREADS * SORT_ITEMS == LBZ_TILE
#define SORT_WARPS (SORT_THREADS / 32)
#define BWT_K 8u                    // bytes covered by the initial radix sort
#define STREAM_THREADS 1024

struct BwtBuffers {
  const uint8_t *T;       // text, slot layout
  uint32_t *sa;           // rotation order (also radix ping)
  uint32_t *sa2;          // radix pong
  uint32_t *rank;         // rank[i] = first position of i's group
  uint8_t *head;          // group-head flags after the initial sort
  uint64_t *key, *key2;   // round keys (ping/pong)
  uint32_t *val, *val2;   // round payload = rotation index (ping/pong)
  uint32_t *pos, *pos2;   // SA position of each tied element (ping/pong)
  uint32_t *gs, *gs2;     // group start of each tied element (ping/pong)
  uint32_t *tstat;        // chained-scan status words [(b*tiles1 + tile)*256 + digit]
  uint32_t *gbase;        // digit bases: [b*256 + d] text passes, then [nb*256 + (b*5 + pass)*256 + d] key passes
  uint32_t *khist;        // key-pass digit histograms [(b*5 + pass)*256 + d]
  void *agg;              // per-tile aggregates of the rank/refine passes
  uint32_t *counters;     // [0] = max unsorted over blocks, [1] = total unsorted, [3] = error flag
  uint32_t *epoch;        // host-side pass counter (status-word epoch)
  int hints;              // use L2 residency hints for the rank scatter/gather
  void (*on_sorted)(void *);  // host callback after the initial sort has been enqueued (lane staggering)
  void *on_sorted_arg;
  uint8_t *bwt;           // output last column
  uint32_t K;             // bytes covered by the initial radix sort (5..8)
  uint32_t *hbits, *cbits; // group-head / size-class bitmaps (1 bit per order position), alias of `head`
  uint32_t *tickets;       // [1024] work-item counters of the persistent pass kernel, indexed by status epoch (zero before use)
  uint32_t *wl;            // work lists of the persistent pass kernel: text passes at [0], list passes at [wl_list_off]
  uint32_t *wl_count;      // [2] number of work items: text passes, list passes
  uint32_t wl_list_off;
};

__device__ __forceinline__ uint32_t wrap_add(uint32_t v, uint32_t d, uint32_t n) {
  uint32_t j = v + d;
  if (j >= n) { j -= n; if (j >= n) j %= n; }
  return j;
}

// ---------------------------------------------------------------------------
// Identity order plus the key word of the first four passes: text bytes 4..7 of
// every rotation (contiguous reads -- the order is still the identity).
__global__ void k_sa_init(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
                          uint32_t *__restrict__ sa, uint32_t *__restrict__ key32) {
  const uint32_t b = blockIdx.y;
  const uint32_t n = meta[b].n;
  const uint32_t base = blockIdx.x * LBZ_TILE;
  if (base >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *Tb = T + off;
  for (uint32_t i = base + threadIdx.x; i < min(base + LBZ_TILE, n); i += blockDim.x) {
    sa[off + i] = i;
    uint32_t k = 0;
    if (i + 8 <= n) {
      k = ((uint32_t)Tb[i + 4] << 24) | ((uint32_t)Tb[i + 5] << 16) | ((uint32_t)Tb[i + 6] << 8) | Tb[i + 7];
    } else {
      uint32_t j = (i + 4) % n;
#pragma unroll
      for (int d = 0; d < 4; d++) { k = (k << 8) | Tb[j]; if (++j >= n) j = 0; }
    }
    key32[off + i] = k;
  }
}

// ---------------------------------------------------------------------------
// Counting sort pass, segmented by block.  MODE 0: digit = text byte
// T[(v + d) mod n], payload = rotation index only.  MODE 1: digit = byte
// `shift/8` of a 64-bit key that travels with the payload.

// One counting-sort pass in ONE launch ("onesweep"): every tile ranks its items
// per digit, publishes its per-digit counts and obtains its per-digit offset by
// a decoupled look-back over the preceding tiles of the same block (chained
// scan).  A status word carries flag (2 bits: 1 = tile aggregate, 2 = inclusive
// prefix), a 10-bit epoch (pass counter, so the array never needs clearing
// between passes) and a 20-bit count.  Tiles are dispatched in blockIdx.x order,
// so a tile only ever waits for tiles that are already resident or finished.
#define TS_FLAG_AGG 0x40000000u
#define TS_FLAG_PREFIX 0x80000000u
#define TS_EPOCH_MASK 0x3FF00000u
#define TS_VALUE_MASK 0x000FFFFFu
#define TS_SPIN_LIMIT (1u << 27)
#define TS_WINDOW 8                // status words fetched per look-back step (k_text_pass2)

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// K = key type carried with the rotation index (u32: four text bytes, u64: round
// key).  COUNT_U: segment size = tied-set size (round passes) instead of n.
// GATHER: the key is not read from src_key but fetched from the text: the four
// leading bytes of rotation `val` (used once, by the 5th pass of the initial sort).
template <typename K, int THREADS, int ITEMS>
struct RadixSmem {
  K skey[THREADS * ITEMS];
  uint32_t sval[THREADS * ITEMS];
  uint32_t wcnt[THREADS / 32][256];
  uint32_t dstart[256];
  uint32_t delta[256];
  uint32_t ws[40];
};

__device__ __forceinline__ uint32_t text_key4(const uint8_t *__restrict__ Tb, uint32_t v, uint32_t n) {
  if (v + 4 <= n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(Tb + (v & ~3u));
    const uint32_t x = __funnelshift_r(w[0], w[1], 8u * (v & 3u));   // bytes v..v+3, little endian
    return __byte_perm(x, 0, 0x0123);                                // first byte most significant
  }
  uint32_t k = 0, j = v;
#pragma unroll
  for (int d = 0; d < 4; d++) { k = (k << 8) | Tb[j]; if (++j >= n) j = 0; }
  return k;
}

template <typename K, int COUNT_U, int GATHER, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, (THREADS * ITEMS <= 2048 ? 4 : 2))
k_radix_pass(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint32_t *__restrict__ src_val, uint32_t *__restrict__ dst_val,
             const K *__restrict__ src_key, K *__restrict__ dst_key,
             uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase, uint32_t gstride,
             uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err) {
  extern __shared__ __align__(16) unsigned char radix_smem_raw[];
  RadixSmem<K, THREADS, ITEMS> &S = *reinterpret_cast<RadixSmem<K, THREADS, ITEMS> *>(radix_smem_raw);
  constexpr uint32_t RTILE = THREADS * ITEMS;
  constexpr int NW = THREADS / 32;
  const uint32_t rtiles = g.S1 / RTILE;            // status rows per block slot
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t cnt = COUNT_U ? meta[b].ul : meta[b].n;     // round passes sort list L only
  const uint32_t tbase = tile * RTILE;
  if (tbase >= cnt) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b) + (COUNT_U ? meta[b].lbase : 0u);
  const uint32_t tile_cnt = min(RTILE, cnt - tbase);
  for (uint32_t i = threadIdx.x; i < NW * 256; i += THREADS) (&S.wcnt[0][0])[i] = 0;
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t lt = lanemask_lt();
  uint32_t val[ITEMS];
  uint32_t rd[ITEMS];     // rank within warp strip | digit << 16 | valid << 31
  K key[ITEMS];

#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const uint32_t idx = tbase + warp * (32 * ITEMS) + it * 32 + lane;
    const bool valid = idx < cnt;
    val[it] = valid ? src_val[off + idx] : 0u;
    if (!GATHER) key[it] = valid ? src_key[off + idx] : (K)0;
    rd[it] = valid ? 0x80000000u : 0u;
  }
  if (GATHER) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = rd[it] ? (K)text_key4(T + off, val[it], n) : (K)0;
  }
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const bool valid = rd[it] != 0;
    const uint32_t digit = valid ? ((uint32_t)(key[it] >> shift) & 0xFFu) : 0x100u;   // invalid lanes: own match group
    const uint32_t mask = __match_any_sync(0xffffffffu, digit);
    uint32_t base = 0;
    if (valid) base = S.wcnt[warp][digit];
    __syncwarp();
    if (valid && (mask & lt) == 0) S.wcnt[warp][digit] = base + __popc(mask);   // group leader
    __syncwarp();
    rd[it] |= (digit << 16) | (base + __popc(mask & lt));
  }
  __syncthreads();
  // per-digit totals of the tile; publish them at once so that successors can proceed
  const uint32_t d = threadIdx.x;
  uint32_t total = 0;
  uint32_t *mine = tstat + ((size_t)b * rtiles + tile) * 256 + (d & 255u);
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  if (d < 256) {
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const uint32_t c = S.wcnt[w][d];
      S.wcnt[w][d] = total;
      total += c;
    }
    st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total);
  }
  uint32_t tsum;
  const uint32_t dst0 = cta_excl_sum(total, S.ws, &tsum);      // start of digit d inside the tile
  if (d < 256) S.dstart[d] = dst0;
  __syncthreads();
  // stage the tile in digit order (needs only tile-local offsets) ...
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    if (rd[it] & 0x80000000u) {
      const uint32_t digit = (rd[it] >> 16) & 0xFFu;
      const uint32_t slot = S.dstart[digit] + S.wcnt[warp][digit] + (rd[it] & 0xFFFFu);
      S.skey[slot] = key[it];
      S.sval[slot] = val[it];
    }
  }
  // ... then resolve the global offset of every digit by looking back over the
  // preceding tiles of this block (by now they have usually published)
  if (d < 256) {
    uint32_t excl = 0;
    if (tile != 0) {
      const uint32_t *look = mine - 256;
      uint32_t spins = 0;
      for (;;) {
        const uint32_t sw = ld_volatile_u32(look);
        if ((sw & TS_EPOCH_MASK) != ep || (sw >> 30) == 0u) {
          if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }     // never expected: fail loudly on the host
          __nanosleep(40);
          continue;
        }
        excl += sw & TS_VALUE_MASK;
        if (sw & TS_FLAG_PREFIX) break;
        look -= 256;
      }
      st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total) & TS_VALUE_MASK));
    }
    S.delta[d] = gbase[(size_t)b * gstride + d] + excl - dst0;   // global index = delta[digit] + tile-local slot
  }
  __syncthreads();
  // write out with consecutive threads on consecutive slots (runs of equal
  // digits are contiguous in global memory)
  for (uint32_t i = threadIdx.x; i < tile_cnt; i += THREADS) {
    const K k = S.skey[i];
    const uint32_t dst = S.delta[(uint32_t)(k >> shift) & 0xFFu] + i;
    dst_key[off + dst] = k;
    dst_val[off + dst] = S.sval[i];
  }
}

// The eight passes of the initial sort, specialised: (key, index) travel as one
// 8-byte pair (one 64-bit load/store per element instead of two 32-bit ones).
//   MODE 0: pair read from `src`
//   MODE 1: index read from `src`, key = the four leading text bytes of that rotation
//           (pass 5: bytes 0..3 replace bytes 4..7 as the carried key)
//   MODE 2: first pass: index = position, key = text bytes 4..7 (no input arrays)
//   LAST  : last pass: only the index is written (to `sa`), the key is dropped
struct TextSmem {
  uint2 spair[512 * 8];
  uint32_t wcnt[16][256];
  uint32_t dstart[256];
  uint32_t delta[256];
  uint32_t ws[40];
};

template <int MODE, int LAST, int MINB>
__global__ void __launch_bounds__(512, MINB)
k_text_pass(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
            const uint2 *__restrict__ src, uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
            uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
            uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff) {
  extern __shared__ __align__(16) unsigned char radix_smem_raw[];
  TextSmem &S = *reinterpret_cast<TextSmem *>(radix_smem_raw);
  constexpr int THREADS = 512, ITEMS = 8, NW = 16;
  constexpr uint32_t RTILE = THREADS * ITEMS;
  const uint32_t rtiles = g.S1 / RTILE;
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t cnt = meta[b].n;
  const uint32_t tbase = tile * RTILE;
  if (tbase >= cnt) return;
  const uint32_t n = cnt;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t tile_cnt = min(RTILE, cnt - tbase);
  for (uint32_t i = threadIdx.x; i < NW * 256; i += THREADS) (&S.wcnt[0][0])[i] = 0;
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t lt = lanemask_lt();
  uint32_t val[ITEMS], key[ITEMS], rd[ITEMS];

#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const uint32_t idx = tbase + warp * (32 * ITEMS) + it * 32 + lane;
    const bool valid = idx < cnt;
    rd[it] = valid ? 0x80000000u : 0u;
    if (MODE == 2) {
      val[it] = idx; key[it] = 0;
    } else {
      const uint2 pr = valid ? src[off + idx] : make_uint2(0u, 0u);
      key[it] = pr.x; val[it] = pr.y;
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = rd[it] ? text_key4(T + off, val[it], n) : 0u;
  }
  if (MODE == 2) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = rd[it] ? text_key4(T + off, wrap_add(val[it], koff, n), n) : 0u;
  }
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const bool valid = rd[it] != 0;
    const uint32_t digit = valid ? ((key[it] >> shift) & 0xFFu) : 0x100u;
    const uint32_t mask = __match_any_sync(0xffffffffu, digit);
    uint32_t base = 0;
    if (valid) base = S.wcnt[warp][digit];
    __syncwarp();
    if (valid && (mask & lt) == 0) S.wcnt[warp][digit] = base + __popc(mask);
    __syncwarp();
    rd[it] |= (digit << 16) | (base + __popc(mask & lt));
  }
  __syncthreads();
  const uint32_t d = threadIdx.x;
  uint32_t total = 0;
  uint32_t *mine = tstat + ((size_t)b * rtiles + tile) * 256 + (d & 255u);
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  if (d < 256) {
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const uint32_t c = S.wcnt[w][d];
      S.wcnt[w][d] = total;
      total += c;
    }
    st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total);
  }
  uint32_t tsum;
  const uint32_t dst0 = cta_excl_sum(total, S.ws, &tsum);
  if (d < 256) S.dstart[d] = dst0;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    if (rd[it] & 0x80000000u) {
      const uint32_t digit = (rd[it] >> 16) & 0xFFu;
      const uint32_t slot = S.dstart[digit] + S.wcnt[warp][digit] + (rd[it] & 0xFFFFu);
      S.spair[slot] = make_uint2(key[it], val[it]);
    }
  }
  if (d < 256) {
    uint32_t excl = 0;
    if (tile != 0) {
      const uint32_t *look = mine - 256;
      uint32_t spins = 0;
      for (;;) {
        const uint32_t sw = ld_volatile_u32(look);
        if ((sw & TS_EPOCH_MASK) != ep || (sw >> 30) == 0u) {
          if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }
          __nanosleep(40);
          continue;
        }
        excl += sw & TS_VALUE_MASK;
        if (sw & TS_FLAG_PREFIX) break;
        look -= 256;
      }
      st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total) & TS_VALUE_MASK));
    }
    S.delta[d] = gbase[(size_t)b * 256 + d] + excl - dst0;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tile_cnt; i += THREADS) {
    const uint2 pr = S.spair[i];
    const uint32_t o = S.delta[(pr.x >> shift) & 0xFFu] + i;
    if (LAST) sa_out[off + o] = pr.y; else dst[off + o] = pr;
  }
}


// Second implementation of the same pass, written for instruction count: loads are
// issued before the counters are cleared, tile-bound checks are one compare against a
// per-thread limit, in-warp ranks are packed four to a register, the per-digit
// totals are summed by all 512 threads (two halves of eight warps each) and scanned
// with two barriers.  Same shared-memory protocol and status words as k_text_pass.
struct TextSmem2 {
  uint2 spair[512 * 8];
  uint32_t wcnt[16][256];
  uint32_t hsum[2][256];
  uint32_t dstart[2][256];
  uint32_t delta[256];
  uint32_t ws[8];
};

// LIST = 1: the same pass over the (32-bit key, index) pairs of a block's round list L
// (segment = meta.ul elements at list offset meta.lbase; digit bases per block and pass).
template <int MODE, int LAST, int MINB, int LIST = 0>
__global__ void __launch_bounds__(512, MINB)
k_text_pass2(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint2 *__restrict__ src, uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
             uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
             uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff, uint32_t gstride,
             uint32_t xpose, uint32_t nb) {
  extern __shared__ __align__(16) unsigned char radix_smem_raw[];
  TextSmem2 &S = *reinterpret_cast<TextSmem2 *>(radix_smem_raw);
  constexpr int THREADS = 512, ITEMS = 8;
  constexpr uint32_t RTILE = THREADS * ITEMS;
  const uint32_t rtiles = g.S1 / RTILE;
  // xpose: grid = (blocks of a group, tiles, groups) -- CTAs are dispatched x-fastest, so consecutive
  // CTAs then work on different blocks and a tile's predecessor has usually published its inclusive
  // prefix.  The group is the whole batch except for the pass that gathers from the text (MODE 1):
  // there 32 blocks share the resident CTAs, so that their text (29 MB) stays in L2 -- spread over
  // all blocks of the batch every gathered sector came from DRAM (5.2 GB per launch instead of 0.8).
  const uint32_t b = xpose ? blockIdx.z * gridDim.x + blockIdx.x : blockIdx.y, tile = xpose ? blockIdx.y : blockIdx.x;
  if (b >= nb) return;
  const uint32_t n = meta[b].n;
  const uint32_t cnt = LIST ? meta[b].ul : n;
  const uint32_t tbase = tile * RTILE;
  if (tbase >= cnt) return;
  const uint32_t off = lbz_slot_off(g, b) + (LIST ? meta[b].lbase : 0u);
  const uint32_t tile_cnt = min(RTILE, cnt - tbase);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  if (MODE == 0 && xpose > 1u && tid == 0) {
    // TMA bulk prefetch into L2 of the tile that the CTA dispatched `xpose` rows later will load
    // (block-fastest dispatch: that CTA starts about when the CTAs resident now retire)
    const uint32_t t2 = tbase + xpose * RTILE;
    if (t2 < cnt) {
      const uint32_t e0 = (off + t2) & ~1u;
      const uint32_t bytes = (min(RTILE, cnt - t2) * 8u) & ~15u;
      if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + e0), "r"(bytes) : "memory");
    }
  }
  const uint32_t wbase = warp * (32 * ITEMS) + lane;              // tile index of this thread's item 0
  const uint32_t lim = tile_cnt > wbase ? tile_cnt - wbase : 0u;  // item `it` exists iff it * 32 < lim

  uint32_t val[ITEMS], key[ITEMS];
  if (MODE == 2) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) val[it] = tbase + wbase + it * 32;
  } else {
    const uint2 *sp = src + off + tbase + wbase;
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const uint2 pr = (it * 32u < lim) ? sp[it * 32] : make_uint2(0u, 0u);
      key[it] = pr.x; val[it] = pr.y;
    }
  }
  {
    uint4 *z = reinterpret_cast<uint4 *>(&S.wcnt[0][0]);
    z[tid] = make_uint4(0u, 0u, 0u, 0u);
    z[tid + THREADS] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (MODE == 1) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = (it * 32u < lim) ? text_key4(T + off, val[it], n) : 0u;
  }
  if (MODE == 2) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = (it * 32u < lim) ? text_key4(T + off, wrap_add(val[it], koff, n), n) : 0u;
  }
  __syncthreads();

  const uint32_t lt = lanemask_lt();
  uint32_t *wrow = &S.wcnt[warp][0];
  uint32_t rkp[ITEMS / 4];                                         // in-warp ranks (< 256), four per register
#pragma unroll
  for (int q = 0; q < ITEMS / 4; q++) rkp[q] = 0;
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const bool valid = it * 32u < lim;
    const uint32_t digit = valid ? ((key[it] >> shift) & 0xFFu) : 0x100u;   // missing items: a group of their own
    const uint32_t mask = __match_any_sync(0xffffffffu, digit);
    uint32_t base = 0;
    if (valid) base = wrow[digit];
    __syncwarp();
    if (valid && (mask & lt) == 0) wrow[digit] = base + __popc(mask);       // group leader
    __syncwarp();
    rkp[it >> 2] |= (base + __popc(mask & lt)) << (8 * (it & 3));
  }
  __syncthreads();

  // per-digit totals: thread (d, half) sums eight warps, the halves meet in hsum
  const uint32_t d = tid & 255u, half = tid >> 8;
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const uint32_t c = S.wcnt[half * 8 + w][d]; S.wcnt[half * 8 + w][d] = run; run += c; }
    S.hsum[half][d] = run;
  }
  __syncthreads();
  uint32_t *mine = tstat + ((size_t)b * rtiles + tile) * 256 + d;
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  uint32_t inc = 0, total = 0, h0 = 0;
  if (half == 0) {
    h0 = S.hsum[0][d];
    total = h0 + S.hsum[1][d];
    st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total);
    inc = warp_incl_sum(total);
    if (lane == 31) S.ws[warp] = inc;
  }
  __syncthreads();
  uint32_t dst0 = 0;
  if (half == 0) {
    uint32_t wb = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) wb += (w < warp) ? S.ws[w] : 0u;
    dst0 = wb + inc - total;                                       // start of digit d inside the tile
    S.dstart[0][d] = dst0;                                         // warps 0..7
    S.dstart[1][d] = dst0 + h0;                                    // warps 8..15 come after the first half's items
  }
  __syncthreads();
  {
    const uint32_t *drow = &S.dstart[half][0];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      if (it * 32u < lim) {
        const uint32_t digit = (key[it] >> shift) & 0xFFu;
        const uint32_t slot = drow[digit] + wrow[digit] + ((rkp[it >> 2] >> (8 * (it & 3))) & 0xFFu);
        S.spair[slot] = make_uint2(key[it], val[it]);
      }
    }
  }
  if (half == 0) {
    uint32_t excl = 0;
    if (tile != 0) {
      // Look-back over the preceding tiles of this block, a window of TS_WINDOW status
      // words at a time: the loads of a window are independent and issued together, so
      // the walk costs one memory round trip per window instead of one per tile (the
      // predecessors are typically still in flight and only offer their aggregates).
      const uint32_t *row0 = mine - (size_t)tile * 256;            // status word of tile 0, this digit
      int t = (int)tile - 1;
      uint32_t spins = 0;
      bool done = false;
      while (!done) {
        uint32_t sw[TS_WINDOW];
#pragma unroll
        for (int k = 0; k < TS_WINDOW; k++)
          sw[k] = (t - k >= 0) ? ld_volatile_u32(row0 + (size_t)(t - k) * 256) : (TS_FLAG_PREFIX | ep);
        int used = 0;
#pragma unroll
        for (int k = 0; k < TS_WINDOW; k++) {
          if (!done && used == k) {
            const uint32_t w = sw[k];
            if ((w & TS_EPOCH_MASK) == ep && (w >> 30) != 0u) {
              excl += w & TS_VALUE_MASK;
              used = k + 1;
              if (w & TS_FLAG_PREFIX) done = true;
            }
          }
        }
        t -= used;
        if (!done && used < TS_WINDOW) {                             // tile t has not published yet
          if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }
          __nanosleep(40);
        }
      }
      st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total) & TS_VALUE_MASK));
    }
    S.delta[d] = gbase[(size_t)b * gstride + d] + excl - dst0;
  }
  __syncthreads();
  if (LAST) {
    uint32_t *o = sa_out + off;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
      const uint32_t i = tid + k * THREADS;
      if (i < tile_cnt) { const uint2 pr = S.spair[i]; o[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr.y; }
    }
  } else {
    uint2 *o = dst + off;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
      const uint32_t i = tid + k * THREADS;
      if (i < tile_cnt) { const uint2 pr = S.spair[i]; o[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr; }
    }
  }
}



// ---- k_text_pass2s: the same pass, specialised (LBZ_TP_SPEC=1).  26 % fewer instructions, measured 5 % SLOWER
// than k_text_pass2 (0.535 against 0.507 ms per launch; with the atomic ranking 0.557 ms): the pass is bound by the
// latency of its shared-memory round trips and barriers, not by issue slots (profiles/r02_pass_variants.md).
// SH >= 0: the digit's bit offset is a compile-time constant (byte extraction becomes one PRMT, no
// reload of the launch parameter); FULL: a tile of exactly RTILE elements, no per-item bound checks
// (all tiles of a block but its last).  The digit start table `delta` includes the slot offset, so
// the write-out index is one 32-bit add.
template <int MODE, int LAST, int LIST, int SH, bool FULL, int RANKV>
__device__ __forceinline__ void text_pass2s_tile(TextSmem2 &S, const uint8_t *__restrict__ T, const uint2 *__restrict__ src,
                                                uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
                                                uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
                                                uint32_t shift_rt, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff,
                                                uint32_t gstride, uint32_t b, uint32_t tile, uint32_t n, uint32_t off,
                                                uint32_t tbase, uint32_t tile_cnt_rt, uint32_t rtiles) {
  constexpr int THREADS = 512, ITEMS = 8;
  constexpr uint32_t RTILE = THREADS * ITEMS;
  const uint32_t shift = SH >= 0 ? (uint32_t)SH : shift_rt;
  const uint32_t tile_cnt = FULL ? RTILE : tile_cnt_rt;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint32_t wbase = warp * (32 * ITEMS) + lane;              // tile index of this thread's item 0
  const uint32_t lim = tile_cnt > wbase ? tile_cnt - wbase : 0u;  // item `it` exists iff it * 32 < lim

  uint32_t val[ITEMS], key[ITEMS];
  if (MODE == 2) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) val[it] = tbase + wbase + it * 32;
  } else {
    const uint2 *sp = src + off + tbase + wbase;
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const uint2 pr = (FULL || it * 32u < lim) ? sp[it * 32] : make_uint2(0u, 0u);
      key[it] = pr.x; val[it] = pr.y;
    }
  }
  {
    uint4 *z = reinterpret_cast<uint4 *>(&S.wcnt[0][0]);
    z[tid] = make_uint4(0u, 0u, 0u, 0u);
    z[tid + THREADS] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (MODE == 1) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = (FULL || it * 32u < lim) ? text_key4(T + off, val[it], n) : 0u;
  }
  if (MODE == 2) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) key[it] = (FULL || it * 32u < lim) ? text_key4(T + off, wrap_add(val[it], koff, n), n) : 0u;
  }
  __syncthreads();

  const uint32_t lt = lanemask_lt();
  uint32_t *wrow = &S.wcnt[warp][0];
  uint32_t rkp[ITEMS / 4];                                         // in-warp ranks (< 256), four per register
#pragma unroll
  for (int q = 0; q < ITEMS / 4; q++) rkp[q] = 0;
  if (RANKV == 1) {
    // All eight match operations are issued before the first result is used, and the per-warp digit
    // counters advance by shared-memory atomics of the group leaders whose results are independent of
    // each other: the ranking is no longer a chain of eight (match -> counter load -> counter store)
    // round trips (profiles/r02_ncu_pass_cs.txt: 33 % of the stall samples sat on that chain).
    uint32_t m[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const bool valid = FULL || it * 32u < lim;
      const uint32_t digit = valid ? ((key[it] >> shift) & 0xFFu) : 0x100u;
      m[it] = __match_any_sync(0xffffffffu, digit);
    }
    uint32_t old[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const bool valid = FULL || it * 32u < lim;
      old[it] = 0;
      if (valid && (m[it] & lt) == 0) old[it] = atomicAdd(&wrow[(key[it] >> shift) & 0xFFu], (uint32_t)__popc(m[it]));
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const uint32_t base = __shfl_sync(0xffffffffu, old[it], __ffs(m[it]) - 1);     // from the group leader
      rkp[it >> 2] |= (base + __popc(m[it] & lt)) << (8 * (it & 3));
    }
  } else {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const bool valid = FULL || it * 32u < lim;
      const uint32_t digit = valid ? ((key[it] >> shift) & 0xFFu) : 0x100u;   // missing items: a group of their own
      const uint32_t mask = __match_any_sync(0xffffffffu, digit);
      uint32_t base = 0;
      if (valid) base = wrow[digit];
      __syncwarp();
      if (valid && (mask & lt) == 0) wrow[digit] = base + __popc(mask);       // group leader
      __syncwarp();
      rkp[it >> 2] |= (base + __popc(mask & lt)) << (8 * (it & 3));
    }
  }
  __syncthreads();

  // per-digit totals: thread (d, half) sums eight warps, the halves meet in hsum
  const uint32_t d = tid & 255u, half = tid >> 8;
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const uint32_t c = S.wcnt[half * 8 + w][d]; S.wcnt[half * 8 + w][d] = run; run += c; }
    S.hsum[half][d] = run;
  }
  __syncthreads();
  uint32_t *mine = tstat + ((size_t)b * rtiles + tile) * 256 + d;
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  uint32_t inc = 0, total = 0, h0 = 0;
  if (half == 0) {
    h0 = S.hsum[0][d];
    total = h0 + S.hsum[1][d];
    st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total);
    inc = warp_incl_sum(total);
    if (lane == 31) S.ws[warp] = inc;
  }
  __syncthreads();
  uint32_t dst0 = 0;
  if (half == 0) {
    uint32_t wb = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) wb += (w < warp) ? S.ws[w] : 0u;
    dst0 = wb + inc - total;                                       // start of digit d inside the tile
    S.dstart[0][d] = dst0;                                         // warps 0..7
    S.dstart[1][d] = dst0 + h0;                                    // warps 8..15 come after the first half's items
  }
  __syncthreads();
  {
    const uint32_t *drow = &S.dstart[half][0];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      if (FULL || it * 32u < lim) {
        const uint32_t digit = (key[it] >> shift) & 0xFFu;
        const uint32_t slot = drow[digit] + wrow[digit] + ((rkp[it >> 2] >> (8 * (it & 3))) & 0xFFu);
        S.spair[slot] = make_uint2(key[it], val[it]);
      }
    }
  }
  if (half == 0) {
    uint32_t excl = 0;
    if (tile != 0) {
      // Look-back over the preceding tiles of this block, a window of TS_WINDOW status
      // words at a time: the loads of a window are independent and issued together, so
      // the walk costs one memory round trip per window instead of one per tile (the
      // predecessors are typically still in flight and only offer their aggregates).
      const uint32_t *row0 = mine - (size_t)tile * 256;            // status word of tile 0, this digit
      int t = (int)tile - 1;
      uint32_t spins = 0;
      bool done = false;
      while (!done) {
        uint32_t sw[TS_WINDOW];
#pragma unroll
        for (int k = 0; k < TS_WINDOW; k++)
          sw[k] = (t - k >= 0) ? ld_volatile_u32(row0 + (size_t)(t - k) * 256) : (TS_FLAG_PREFIX | ep);
        int used = 0;
#pragma unroll
        for (int k = 0; k < TS_WINDOW; k++) {
          if (!done && used == k) {
            const uint32_t w = sw[k];
            if ((w & TS_EPOCH_MASK) == ep && (w >> 30) != 0u) {
              excl += w & TS_VALUE_MASK;
              used = k + 1;
              if (w & TS_FLAG_PREFIX) done = true;
            }
          }
        }
        t -= used;
        if (!done && used < TS_WINDOW) {                             // tile t has not published yet
          if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }
          __nanosleep(40);
        }
      }
      st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total) & TS_VALUE_MASK));
    }
    S.delta[d] = gbase[(size_t)b * gstride + d] + excl - dst0 + off;   // slot offset folded in
  }
  __syncthreads();
  if (LAST) {
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
      const uint32_t i = tid + k * THREADS;
      if (FULL || i < tile_cnt) { const uint2 pr = S.spair[i]; sa_out[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr.y; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
      const uint32_t i = tid + k * THREADS;
      if (FULL || i < tile_cnt) { const uint2 pr = S.spair[i]; dst[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr; }
    }
  }
}

template <int MODE, int LAST, int MINB, int LIST = 0, int SH = -1, int RANKV = 0>
__global__ void __launch_bounds__(512, MINB)
k_text_pass2s(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint2 *__restrict__ src, uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
             uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
             uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff, uint32_t gstride,
             uint32_t xpose, uint32_t nb) {
  extern __shared__ __align__(16) unsigned char radix_smem_raw[];
  TextSmem2 &S = *reinterpret_cast<TextSmem2 *>(radix_smem_raw);
  constexpr uint32_t RTILE = 512 * 8;
  const uint32_t rtiles = g.S1 / RTILE;
  // xpose: grid = (blocks of a group, tiles, groups) -- CTAs are dispatched x-fastest, so consecutive
  // CTAs then work on different blocks and a tile's predecessor has usually published its inclusive
  // prefix.  The group is the whole batch except for the pass that gathers from the text (MODE 1):
  // there 32 blocks share the resident CTAs, so that their text (29 MB) stays in L2 -- spread over
  // all blocks of the batch every gathered sector came from DRAM (5.2 GB per launch instead of 0.8).
  const uint32_t b = xpose ? blockIdx.z * gridDim.x + blockIdx.x : blockIdx.y, tile = xpose ? blockIdx.y : blockIdx.x;
  if (b >= nb) return;
  const uint32_t n = meta[b].n;
  const uint32_t cnt = LIST ? meta[b].ul : n;
  const uint32_t tbase = tile * RTILE;
  if (tbase >= cnt) return;
  const uint32_t off = lbz_slot_off(g, b) + (LIST ? meta[b].lbase : 0u);
  const uint32_t tile_cnt = min(RTILE, cnt - tbase);
  if (MODE == 0 && xpose > 1u && threadIdx.x == 0) {
    // TMA bulk prefetch into L2 of the tile that the CTA dispatched `xpose` rows later will load
    // (block-fastest dispatch: that CTA starts about when the CTAs resident now retire)
    const uint32_t t2 = tbase + xpose * RTILE;
    if (t2 < cnt) {
      const uint32_t e0 = (off + t2) & ~1u;
      const uint32_t bytes = (min(RTILE, cnt - t2) * 8u) & ~15u;
      if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + e0), "r"(bytes) : "memory");
    }
  }
  if (tile_cnt == RTILE)
    text_pass2s_tile<MODE, LAST, LIST, SH, true, RANKV>(S, T, src, dst, sa_out, tstat, gbase, shift, epoch, err, koff, gstride, b, tile, n, off,
                                                tbase, tile_cnt, rtiles);
  else
    text_pass2s_tile<MODE, LAST, LIST, SH, false, RANKV>(S, T, src, dst, sa_out, tstat, gbase, shift, epoch, err, koff, gstride, b, tile, n, off,
                                                 tbase, tile_cnt, rtiles);
}

// ---------------------------------------------------------------------------
// Third implementation of the pass: a PERSISTENT kernel fed by TMA.
//
//   * k_worklist lists the non-empty tiles of the batch in tile-major order (tile 0 of every
//     block, then tile 1, ...).  The pass kernel runs 2 CTAs per SM; every CTA loops over work
//     items that it draws from an atomic ticket counter.  Consecutive tickets belong to
//     different blocks, so the look-back of a tile finds the inclusive prefix of its
//     predecessor (drawn a whole row of tickets earlier) after one or two steps instead of
//     walking over hundreds of in-flight aggregates; and a tile only ever waits for tiles with
//     smaller tickets, each of which is held by a CTA that is working on it or on an even
//     smaller one -- no reliance on the hardware's CTA dispatch order.
//   * one elected thread moves the next tile's (key, index) pairs HBM -> shared memory with a
//     1-D bulk copy (cp.async.bulk ... mbarrier::complete_tx::bytes) into the landing buffer
//     the CTA is not working on, so the load of item i+1 is in flight while item i is ranked,
//     staged and written out; it fetches ticket, work-list entry and block record for item
//     i+2 in the shadow of the other phases.
//   * warp-private counters and ranks as before, then one column-sum step, after which EVERY
//     warp scans the 256 digit totals for itself (8 digits per lane) and folds the result
//     into its own counter row: three barriers per tile instead of six; the landing buffer
//     doubles as the staging buffer of the digit-ordered tile.
// Status words, digit bases and the write-out are those of k_text_pass2.
#define WL_MAXBLOCKS 32768u
#define WL_INVALID 0xFFFFFFFFu
template <int LIST>
__global__ void __launch_bounds__(1024)
k_worklist(const LbzBlockMeta *__restrict__ meta, uint32_t nb, uint32_t rows, uint32_t *__restrict__ wl,
           uint32_t *__restrict__ wl_count) {
  __shared__ uint8_t nt[WL_MAXBLOCKS];
  __shared__ uint32_t rowstart[256];
  __shared__ uint32_t ws[40];
  const uint32_t tid = threadIdx.x;
  for (uint32_t b = tid; b < nb; b += 1024) {
    const uint32_t cnt = LIST ? meta[b].ul : meta[b].n;
    nt[b] = (uint8_t)min((cnt + 4095u) / 4096u, rows);
  }
  __syncthreads();
  uint32_t c = 0;
  if (tid < rows) for (uint32_t b = 0; b < nb; b++) c += (nt[b] > tid);
  uint32_t total;
  const uint32_t start = cta_excl_sum(c, ws, &total);
  if (tid < 256) rowstart[tid] = start;
  __syncthreads();
  if (tid < rows && c) {
    uint32_t o = rowstart[tid];
    for (uint32_t b = 0; b < nb; b++) if (nt[b] > tid) wl[o++] = (b << 8) | tid;
  }
  if (tid == 0) *wl_count = total;
}

struct PassDesc { uint32_t b, tile, cnt, off, n, valid, shift1, pad; };
struct PassSmem {
  uint2 buf[2][4096 + 2];
  uint32_t wcnt[16][256];
  uint32_t hsum[2][256];
  uint32_t dstart[256];
  uint32_t delta[256];
  PassDesc desc[2];
  unsigned long long full[2];
  // state of the elected thread (kept out of the register file: the other 511 threads would carry it too)
  PassDesc nxt;
  uint32_t qk, qend, total, pad;
};
#define P3_THREADS 512
#define P3_ELECTED 480u          // warp 15, lane 0: idle while the first eight warps look back
#define P3_GRAB 2u

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Returns false if the phase did not complete within ~2^24 polls (never expected; the caller
// raises the error flag and leaves instead of hanging the device).
__device__ __forceinline__ bool mbar_wait(void *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok, polls = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok && ++polls < (1u << 24));
  return ok != 0u;
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Elected thread: publish the descriptor of the next item in slot s and start its load.
template <int MODE>
__device__ __forceinline__ void pass3_issue(PassSmem &S, uint32_t s, const PassDesc &d0, const uint2 *__restrict__ src) {
  PassDesc d = d0;
  d.shift1 = 0u; d.pad = 0u;
  if (d.valid && MODE == 0) {
    // 16-byte aligned source: start one pair early when the tile starts at an odd element
    const uint32_t e0 = d.off + d.tile * 4096u;
    const uint32_t tile_cnt = min(4096u, d.cnt - d.tile * 4096u);
    d.shift1 = e0 & 1u;
    const uint32_t bytes = ((d.shift1 + tile_cnt + 1u) & ~1u) * 8u;
    S.desc[s] = d;
    mbar_arrive_expect_tx(&S.full[s], bytes);
    bulk_g2s(&S.buf[s][0], src + (e0 - d.shift1), bytes, &S.full[s]);
  } else {
    S.desc[s] = d;
    mbar_arrive(&S.full[s]);
  }
}
template <int LIST>
__device__ __forceinline__ PassDesc pass3_describe(const LbzGeom &g, const LbzBlockMeta *__restrict__ meta, uint32_t we) {
  PassDesc d;
  d.valid = 0u; d.b = 0u; d.tile = 0u; d.cnt = 0u; d.off = 0u; d.n = 0u; d.shift1 = 0u; d.pad = 0u;
  if (we != WL_INVALID) {
    d.b = we >> 8; d.tile = we & 255u; d.valid = 1u;
    d.n = meta[d.b].n;
    d.cnt = LIST ? meta[d.b].ul : d.n;
    d.off = lbz_slot_off(g, d.b) + (LIST ? meta[d.b].lbase : 0u);
  }
  return d;
}

template <int MODE, int LAST, int LIST>
__global__ void __launch_bounds__(P3_THREADS, 2)
k_text_pass3(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint2 *__restrict__ src, uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
             uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
             uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff, uint32_t gstride,
             uint32_t *__restrict__ ticket, const uint32_t *__restrict__ wl, const uint32_t *__restrict__ wl_count) {
  extern __shared__ __align__(128) unsigned char pass3_smem_raw[];
  PassSmem &S = *reinterpret_cast<PassSmem *>(pass3_smem_raw);
  constexpr int ITEMS = 8;
  constexpr uint32_t RTILE = 4096u;
  const uint32_t rtiles = g.S1 / RTILE;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const bool elected = tid == P3_ELECTED;

  if (tid == 0) {
    mbar_init(&S.full[0], 1u); mbar_init(&S.full[1], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // elected thread: ticket range [qk, qend), total number of work items, descriptor of the item after the next
  if (elected) {
    const uint32_t total = *wl_count;
    uint32_t qk = atomicAdd(ticket, P3_GRAB);
    const uint32_t qend = qk + P3_GRAB;
    uint32_t we = (qk < total) ? wl[qk] : WL_INVALID; qk++;
    pass3_issue<MODE>(S, 0u, pass3_describe<LIST>(g, meta, we), src);          // item 0 -> buffer 0
    we = (qk < total) ? wl[qk] : WL_INVALID; qk++;
    S.nxt = pass3_describe<LIST>(g, meta, we);                                   // item 1, issued after barrier A of item 0
    S.qk = qk; S.qend = qend; S.total = total;
  }

  const uint32_t lt = lanemask_lt();
  uint32_t *wrow = &S.wcnt[warp][0];
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  for (uint32_t it = 0;; it++) {
    const uint32_t s = it & 1u;
    {   // clear this warp's counter row while the tile lands
      uint4 *z = reinterpret_cast<uint4 *>(wrow);
      z[lane] = make_uint4(0u, 0u, 0u, 0u);
      z[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
    }
    // elected thread: the loads that describe item it+2 are issued early and consumed late (their
    // results stay in registers in between), so that this warp does not hold up the barriers
    uint32_t tk = 0, we = WL_INVALID, m_n = 0, m_cnt = 0, m_lbase = 0;
    const bool grab = elected && S.qk == S.qend;
    if (grab) tk = atomicAdd(ticket, P3_GRAB);               // consumed after barrier A
    if (!mbar_wait(&S.full[s], (it >> 1) & 1u)) { *err = 2u; break; }
    if (!S.desc[s].valid) break;
    const uint32_t b = S.desc[s].b, tile = S.desc[s].tile, n = S.desc[s].n, off = S.desc[s].off;
    const uint32_t tbase = tile * RTILE;
    const uint32_t tile_cnt = min(RTILE, S.desc[s].cnt - tbase);
    const uint32_t wbase = warp * (32 * ITEMS) + lane;
    const uint32_t lim = tile_cnt > wbase ? tile_cnt - wbase : 0u;
    uint2 *stage = &S.buf[s][0];

    uint32_t val[ITEMS], key[ITEMS];
    if (MODE == 0) {
      const uint2 *sp = stage + S.desc[s].shift1 + wbase;
#pragma unroll
      for (int q = 0; q < ITEMS; q++) {
        const uint2 pr = (q * 32u < lim) ? sp[q * 32] : make_uint2(0u, 0u);
        key[q] = pr.x; val[q] = pr.y;
      }
    } else if (MODE == 1) {
      const uint2 *sp = src + off + tbase + wbase;
#pragma unroll
      for (int q = 0; q < ITEMS; q++) val[q] = (q * 32u < lim) ? sp[q * 32].y : 0u;
#pragma unroll
      for (int q = 0; q < ITEMS; q++) key[q] = (q * 32u < lim) ? text_key4(T + off, val[q], n) : 0u;
    } else {
#pragma unroll
      for (int q = 0; q < ITEMS; q++) {
        val[q] = tbase + wbase + q * 32;
        key[q] = (q * 32u < lim) ? text_key4(T + off, wrap_add(val[q], koff, n), n) : 0u;
      }
    }
    __syncwarp();                                            // counter row cleared by all lanes

    uint32_t rkp[ITEMS / 4];
#pragma unroll
    for (int q = 0; q < ITEMS / 4; q++) rkp[q] = 0;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      const bool valid = q * 32u < lim;
      const uint32_t digit = valid ? ((key[q] >> shift) & 0xFFu) : 0x100u;
      const uint32_t mask = __match_any_sync(0xffffffffu, digit);
      uint32_t base = 0;
      if (valid) base = wrow[digit];
      __syncwarp();
      if (valid && (mask & lt) == 0) wrow[digit] = base + __popc(mask);
      __syncwarp();
      rkp[q >> 2] |= (base + __popc(mask & lt)) << (8 * (q & 3));
    }
    __syncthreads();                                         // A: all rows counted, all items in registers,
                                                             //    every thread is done with the other buffer
    if (elected) {
      pass3_issue<MODE>(S, s ^ 1u, S.nxt, src);              // item it+1 lands while this one is processed
      if (grab) { S.qk = tk; S.qend = tk + P3_GRAB; }
      const uint32_t k = S.qk;
      if (k < S.total) we = wl[k];                           // item it+2: work-list entry, consumed after barrier B
      S.qk = k + 1u;
    }
    const uint32_t d = tid & 255u, half = tid >> 8;
    {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) { const uint32_t c = S.wcnt[half * 8 + w][d]; S.wcnt[half * 8 + w][d] = run; run += c; }
      S.hsum[half][d] = run;
    }
    __syncthreads();                                         // B: column sums ready
    if (elected && we != WL_INVALID) {                       // block record of item it+2, consumed after the staging step
      const LbzBlockMeta *mb = meta + (we >> 8);
      m_n = mb->n;
      m_cnt = LIST ? mb->ul : m_n;
      m_lbase = LIST ? mb->lbase : 0u;
    }

    uint32_t *mine = tstat + ((size_t)b * rtiles + tile) * 256 + d;
    uint32_t total_d = 0;
    if (half == 0) {
      total_d = S.hsum[0][d] + S.hsum[1][d];
      st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total_d);
    }
    {   // every warp: exclusive scan of the 256 digit totals, 8 digits per lane, folded into its own row
      uint32_t run;
      {
        const uint4 a0 = *reinterpret_cast<const uint4 *>(&S.hsum[0][lane * 8]);
        const uint4 a1 = *reinterpret_cast<const uint4 *>(&S.hsum[0][lane * 8 + 4]);
        const uint4 b0 = *reinterpret_cast<const uint4 *>(&S.hsum[1][lane * 8]);
        const uint4 b1 = *reinterpret_cast<const uint4 *>(&S.hsum[1][lane * 8 + 4]);
        run = (a0.x + a0.y + a0.z + a0.w) + (a1.x + a1.y + a1.z + a1.w) + (b0.x + b0.y + b0.z + b0.w) + (b1.x + b1.y + b1.z + b1.w);
      }
      uint32_t e = warp_incl_sum(run) - run;                             // start of digit lane*8 inside the tile
      const bool hi = warp >= 8;                                         // second half: after the first half's items
      const bool pub = (lane >> 2) == warp;                              // warps 0..7: the 32 digits their threads look back for
#pragma unroll
      for (int hq = 0; hq < 2; hq++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(&S.hsum[0][lane * 8 + 4 * hq]);
        const uint4 c = *reinterpret_cast<const uint4 *>(&S.hsum[1][lane * 8 + 4 * hq]);
        const uint32_t e0 = e, e1 = e0 + a.x + c.x, e2 = e1 + a.y + c.y, e3 = e2 + a.z + c.z;
        e = e3 + a.w + c.w;
        if (pub) *reinterpret_cast<uint4 *>(&S.dstart[lane * 8 + 4 * hq]) = make_uint4(e0, e1, e2, e3);
        uint4 w = *reinterpret_cast<const uint4 *>(&wrow[lane * 8 + 4 * hq]);
        w.x += e0 + (hi ? a.x : 0u); w.y += e1 + (hi ? a.y : 0u); w.z += e2 + (hi ? a.z : 0u); w.w += e3 + (hi ? a.w : 0u);
        *reinterpret_cast<uint4 *>(&wrow[lane * 8 + 4 * hq]) = w;
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      if (q * 32u < lim) {
        const uint32_t digit = (key[q] >> shift) & 0xFFu;
        const uint32_t slot = wrow[digit] + ((rkp[q >> 2] >> (8 * (q & 3))) & 0xFFu);
        stage[slot] = make_uint2(key[q], val[q]);
      }
    }
    if (elected) {
      PassDesc nd;
      nd.valid = we != WL_INVALID; nd.b = we >> 8; nd.tile = we & 255u; nd.cnt = m_cnt; nd.n = m_n;
      nd.off = lbz_slot_off(g, we >> 8) + m_lbase; nd.shift1 = 0u; nd.pad = 0u;
      S.nxt = nd;
    }
    if (half == 0) {
      uint32_t excl = 0;
      if (tile != 0) {
        const uint32_t *row0 = mine - (size_t)tile * 256;
        int t = (int)tile - 1;
        uint32_t spins = 0;
        bool done = false;
        while (!done) {
          uint32_t sw[TS_WINDOW];
#pragma unroll
          for (int q = 0; q < TS_WINDOW; q++)
            sw[q] = (t - q >= 0) ? ld_volatile_u32(row0 + (size_t)(t - q) * 256) : (TS_FLAG_PREFIX | ep);
          int used = 0;
#pragma unroll
          for (int q = 0; q < TS_WINDOW; q++) {
            if (!done && used == q) {
              const uint32_t w = sw[q];
              if ((w & TS_EPOCH_MASK) == ep && (w >> 30) != 0u) {
                excl += w & TS_VALUE_MASK;
                used = q + 1;
                if (w & TS_FLAG_PREFIX) done = true;
              }
            }
          }
          t -= used;
          if (!done && used < TS_WINDOW) {
            if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }
            __nanosleep(20);
          }
        }
        st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total_d) & TS_VALUE_MASK));
      }
      S.delta[d] = gbase[(size_t)b * gstride + d] + excl - S.dstart[d];
    }
    __syncthreads();                                         // C: tile staged in digit order, offsets known
    if (LAST) {
      uint32_t *o = sa_out + off;
#pragma unroll
      for (int q = 0; q < ITEMS; q++) {
        const uint32_t i = tid + q * P3_THREADS;
        if (i < tile_cnt) { const uint2 pr = stage[i]; o[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr.y; }
      }
    } else {
      uint2 *o = dst + off;
#pragma unroll
      for (int q = 0; q < ITEMS; q++) {
        const uint32_t i = tid + q * P3_THREADS;
        if (i < tile_cnt) { const uint2 pr = stage[i]; o[S.delta[(pr.x >> shift) & 0xFFu] + i] = pr; }
      }
    }
    // this thread's generic-proxy accesses to the buffer are ordered before the bulk copy that
    // refills it (issued by the elected thread after the next barrier A)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
}

// ---------------------------------------------------------------------------
// Fourth implementation of the pass (default): one tile per CTA and three CTAs per SM like
// k_text_pass2 -- the persistent variant above loses more to its lower occupancy and to the
// coupling of consecutive tiles than the prefetch gains (measured: 1.31 ms vs 0.59 ms per pass) --
// with the tile algorithm of k_text_pass3:
//   * the tile's (key, index) pairs arrive by ONE 1-D bulk copy (cp.async.bulk, completion on an
//     mbarrier) issued by the first thread while all threads clear the counters; the landing
//     buffer doubles as the staging buffer of the digit-ordered tile;
//   * three barriers per tile instead of six: after the column sums every warp scans the 256
//     digit totals for itself and folds the result into its own counter row;
//   * full tiles (all but the last of a block) run a specialisation without bounds tests;
//   * CTAs are dispatched block-fastest (grid = blocks x tiles), so a tile's predecessor has
//     usually published its inclusive prefix and the look-back ends after one or two steps.
// Counter rows are skewed by two words per 32 digits (P4_IDX): the scan step reads eight consecutive
// digits per lane, which in a flat row is a stride of 32 B (4-way bank conflicts on 64-bit accesses;
// the first build of this kernel was bound by the shared-memory pipe, profiles/r02_ncu_text_pass4.txt);
// with the skew the 16 lanes of a half-warp cover all 32 banks, and single digits keep their low
// five bits as the bank, which is what spreads the letters of a text over the banks when ranking.
#define P4_ROW 272u
#define P4_IDX(d) ((d) + (((d) >> 5) << 1))
struct Pass4Smem {
  uint2 buf[4096 + 2];
  uint32_t wcnt[16][P4_ROW];
  uint32_t hsum[2][P4_ROW];
  uint32_t dstart[256];
  uint32_t delta[256];
  unsigned long long full;
};

template <int MODE, int LAST, bool FULL, bool RR>
__device__ __forceinline__ void pass4_body(Pass4Smem &S, const uint8_t *__restrict__ T, const uint2 *__restrict__ src,
                                           uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out, uint32_t *__restrict__ tstat,
                                           const uint32_t *__restrict__ gbase, uint32_t shift, uint32_t ep, uint32_t *__restrict__ err,
                                           uint32_t koff, uint32_t n, uint32_t tile, uint32_t tile_cnt, uint32_t shift1,
                                           uint32_t off, uint32_t stat_row, uint32_t gb_row) {
  constexpr int ITEMS = 8;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint32_t wbase = warp * (32 * ITEMS) + lane;
  const uint32_t lim = FULL ? 0xFFFFFFFFu : (tile_cnt > wbase ? tile_cnt - wbase : 0u);
  uint32_t *wrow = &S.wcnt[warp][0];
  uint2 *stage = &S.buf[0];
  // RR ("re-read"): only the digits stay in registers through the counting and scanning steps; the
  // pairs are read again from the landing buffer right before they are staged (one more barrier,
  // sixteen registers fewer while the scan is live -- the kernel is register-bound at 3 CTAs/SM)
  uint32_t val[ITEMS], key[ITEMS];
  uint32_t dgp[ITEMS / 4];
#pragma unroll
  for (int q = 0; q < ITEMS / 4; q++) dgp[q] = 0;
  if (MODE == 0) {
    const uint2 *sp = stage + shift1 + wbase;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      const uint2 pr = (FULL || q * 32u < lim) ? sp[q * 32] : make_uint2(0u, 0u);
      key[q] = pr.x; val[q] = pr.y;
    }
  } else if (MODE == 1) {
    const uint2 *sp = src + off + tile * 4096u + wbase;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) val[q] = (FULL || q * 32u < lim) ? sp[q * 32].y : 0u;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) key[q] = (FULL || q * 32u < lim) ? text_key4(T + off, val[q], n) : 0u;
  } else {
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      val[q] = tile * 4096u + wbase + q * 32;
      key[q] = (FULL || q * 32u < lim) ? text_key4(T + off, wrap_add(val[q], koff, n), n) : 0u;
    }
  }
  if (RR) {
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      if (MODE != 0 && (FULL || q * 32u < lim)) stage[wbase + q * 32] = make_uint2(key[q], val[q]);
      dgp[q >> 2] |= ((key[q] >> shift) & 0xFFu) << (8 * (q & 3));
    }
  }
  const uint32_t lt = lanemask_lt();
  uint32_t rkp[ITEMS / 4];
#pragma unroll
  for (int q = 0; q < ITEMS / 4; q++) rkp[q] = 0;
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    const bool valid = FULL || q * 32u < lim;
    const uint32_t dg = RR ? ((dgp[q >> 2] >> (8 * (q & 3))) & 0xFFu) : ((key[q] >> shift) & 0xFFu);
    const uint32_t digit = valid ? dg : 0x100u;
    const uint32_t mask = __match_any_sync(0xffffffffu, digit);
    const uint32_t pd = P4_IDX(dg);
    uint32_t base = 0;
    if (valid) base = wrow[pd];
    __syncwarp();
    if (valid && (mask & lt) == 0) wrow[pd] = base + __popc(mask);
    __syncwarp();
    rkp[q >> 2] |= (base + __popc(mask & lt)) << (8 * (q & 3));
  }
  __syncthreads();                                           // A: all rows counted, all items in registers

  const uint32_t d = tid & 255u, half = tid >> 8;
  const uint32_t pdd = P4_IDX(d);
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const uint32_t c = S.wcnt[half * 8 + w][pdd]; S.wcnt[half * 8 + w][pdd] = run; run += c; }
    S.hsum[half][pdd] = run;
  }
  __syncthreads();                                           // B: column sums ready

  uint32_t total_d = 0;
  uint32_t *mine = tstat + ((size_t)stat_row + tile) * 256 + d;
  if (half == 0) {
    total_d = S.hsum[0][pdd] + S.hsum[1][pdd];
    st_volatile_u32(mine, (tile == 0 ? TS_FLAG_PREFIX : TS_FLAG_AGG) | ep | total_d);
  }
  {   // every warp: exclusive scan of the 256 digit totals, 8 digits per lane, folded into its own row
    const uint32_t pl = P4_IDX(lane * 8u);                             // this lane's eight digits: one skewed, contiguous run
    uint32_t run = 0;
#pragma unroll
    for (int hq = 0; hq < 4; hq++) {
      const uint2 a = *reinterpret_cast<const uint2 *>(&S.hsum[0][pl + 2 * hq]);
      const uint2 c = *reinterpret_cast<const uint2 *>(&S.hsum[1][pl + 2 * hq]);
      run += a.x + a.y + c.x + c.y;
    }
    uint32_t e = warp_incl_sum(run) - run;                               // start of digit lane*8 inside the tile
    const bool hi = warp >= 8;                                           // second half: after the first half's items
    const bool pub = (lane >> 2) == warp;                                // warps 0..7: the 32 digits their threads look back for
#pragma unroll
    for (int hq = 0; hq < 4; hq++) {
      const uint2 a = *reinterpret_cast<const uint2 *>(&S.hsum[0][pl + 2 * hq]);
      const uint2 c = *reinterpret_cast<const uint2 *>(&S.hsum[1][pl + 2 * hq]);
      const uint32_t e0 = e, e1 = e0 + a.x + c.x;
      e = e1 + a.y + c.y;
      if (pub) *reinterpret_cast<uint2 *>(&S.dstart[lane * 8 + 2 * hq]) = make_uint2(e0, e1);
      uint2 w = *reinterpret_cast<const uint2 *>(&wrow[pl + 2 * hq]);
      w.x += e0 + (hi ? a.x : 0u); w.y += e1 + (hi ? a.y : 0u);
      *reinterpret_cast<uint2 *>(&wrow[pl + 2 * hq]) = w;
    }
  }
  __syncwarp();
  if (RR) {
    const uint2 *sp = stage + (MODE == 0 ? shift1 : 0u) + wbase;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      const uint2 pr = (FULL || q * 32u < lim) ? sp[q * 32] : make_uint2(0u, 0u);
      key[q] = pr.x; val[q] = pr.y;
    }
    __syncthreads();                                         // every pair is back in registers: the buffer may be overwritten
  }
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    if (FULL || q * 32u < lim) {
      const uint32_t digit = RR ? ((dgp[q >> 2] >> (8 * (q & 3))) & 0xFFu) : ((key[q] >> shift) & 0xFFu);
      const uint32_t slot = wrow[P4_IDX(digit)] + ((rkp[q >> 2] >> (8 * (q & 3))) & 0xFFu);
      stage[slot] = make_uint2(key[q], val[q]);
    }
  }
  if (half == 0) {
    uint32_t excl = 0;
    if (tile != 0) {
      const uint32_t *row0 = mine - (size_t)tile * 256;
      int t = (int)tile - 1;
      uint32_t spins = 0;
      bool done = false;
      while (!done) {
        uint32_t sw[TS_WINDOW];
#pragma unroll
        for (int q = 0; q < TS_WINDOW; q++)
          sw[q] = (t - q >= 0) ? ld_volatile_u32(row0 + (size_t)(t - q) * 256) : (TS_FLAG_PREFIX | ep);
        int used = 0;
#pragma unroll
        for (int q = 0; q < TS_WINDOW; q++) {
          if (!done && used == q) {
            const uint32_t w = sw[q];
            if ((w & TS_EPOCH_MASK) == ep && (w >> 30) != 0u) {
              excl += w & TS_VALUE_MASK;
              used = q + 1;
              if (w & TS_FLAG_PREFIX) done = true;
            }
          }
        }
        t -= used;
        if (!done && used < TS_WINDOW) {
          if (++spins > TS_SPIN_LIMIT) { *err = 1u; break; }
          __nanosleep(20);
        }
      }
      st_volatile_u32(mine, TS_FLAG_PREFIX | ep | ((excl + total_d) & TS_VALUE_MASK));
    }
    S.delta[d] = gbase[(size_t)gb_row + d] + excl - S.dstart[d];
  }
  __syncthreads();                                           // C: tile staged in digit order, offsets known
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    const uint32_t i = tid + q * 512u;
    if (FULL || i < tile_cnt) {
      const uint2 pr = stage[i];
      const uint32_t o = S.delta[(pr.x >> shift) & 0xFFu] + i;
      if (LAST) sa_out[off + o] = pr.y; else dst[off + o] = pr;
    }
  }
}

template <int MODE, int LAST, int LIST, bool RR>
__global__ void __launch_bounds__(512, 3)
k_text_pass4(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint2 *__restrict__ src, uint2 *__restrict__ dst, uint32_t *__restrict__ sa_out,
             uint32_t *__restrict__ tstat, const uint32_t *__restrict__ gbase,
             uint32_t shift, uint32_t epoch, uint32_t *__restrict__ err, uint32_t koff, uint32_t gstride) {
  extern __shared__ __align__(128) unsigned char pass4_smem_raw[];
  Pass4Smem &S = *reinterpret_cast<Pass4Smem *>(pass4_smem_raw);
  const uint32_t b = blockIdx.x, tile = blockIdx.y;           // block-fastest dispatch
  const uint32_t n = meta[b].n;
  const uint32_t cnt = LIST ? meta[b].ul : n;
  const uint32_t tbase = tile * 4096u;
  if (tbase >= cnt) return;
  const uint32_t off = lbz_slot_off(g, b) + (LIST ? meta[b].lbase : 0u);
  const uint32_t tile_cnt = min(4096u, cnt - tbase);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t shift1 = (MODE == 0) ? ((off + tbase) & 1u) : 0u;
  if (MODE == 0 && tid == 0) {
    // 16-byte aligned source: start one pair early when the tile starts at an odd element
    const uint32_t bytes = ((shift1 + tile_cnt + 1u) & ~1u) * 8u;
    mbar_init(&S.full, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(&S.full, bytes);
    bulk_g2s(&S.buf[0], src + (off + tbase - shift1), bytes, &S.full);
  }
  {
    uint4 *z = reinterpret_cast<uint4 *>(&S.wcnt[tid >> 5][0]);             // P4_ROW = 272 words = 68 uint4
    z[lane] = make_uint4(0u, 0u, 0u, 0u);
    z[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 4) z[lane + 64] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (MODE == 0) {
    __syncthreads();                                          // the mbarrier is initialised
    if (!mbar_wait(&S.full, 0u)) { *err = 2u; return; }
  } else {
    __syncwarp();
  }
  const uint32_t ep = (epoch << 20) & TS_EPOCH_MASK;
  const uint32_t stat_row = b * (g.S1 / 4096u), gb_row = b * gstride;
  if (tile_cnt == 4096u)
    pass4_body<MODE, LAST, true, RR>(S, T, src, dst, sa_out, tstat, gbase, shift, ep, err, koff, n, tile, tile_cnt, shift1, off, stat_row, gb_row);
  else
    pass4_body<MODE, LAST, false, RR>(S, T, src, dst, sa_out, tstat, gbase, shift, ep, err, koff, n, tile, tile_cnt, shift1, off, stat_row, gb_row);
}

// Pass variants (LBZ_TP_VER): 2 = one CTA per tile, dispatched block-fastest, TMA bulk prefetch of the tile
// two rows ahead into L2 (k_text_pass2, default: the fastest measured, profiles/r02_pass_variants.md),
// 4 = one CTA per tile, TMA-fed, three barriers (k_text_pass4: bound by the shared-memory pipe),
// 3 = persistent TMA kernel (k_text_pass3: bound by occupancy), 2 = one CTA per tile
// (k_text_pass2; LBZ_TP_XPOSE=1 dispatches it block-fastest), 1 = first implementation.
static int tp_version() {
  static int v = -1;
  if (v < 0) { const char *ev = getenv("LBZ_TP_VER"); v = ev ? atoi(ev) : 2; if (v < 1 || v > 4) v = 2; }
  return v;
}
// 0 = tile-fastest dispatch, 1 = block-fastest, n >= 2 = block-fastest + TMA bulk prefetch into L2 of the
// tile n rows ahead
static uint32_t tp_xpose() {
  static int v = -1;
  if (v < 0) { const char *ev = getenv("LBZ_TP_XPOSE"); v = ev ? atoi(ev) : 2; if (v < 0 || v > 8) v = 2; }
  return (uint32_t)v;
}
// 1 (default): kernels with the digit offset as a compile-time constant; 0: one kernel per pass kind, offset
// read from the launch parameters
static bool tp_const_shift() {
  static int v = -1;
  if (v < 0) { const char *ev = getenv("LBZ_TP_CONST_SHIFT"); v = ev ? (atoi(ev) != 0) : 1; }
  return v != 0;
}
// 0 (default): k_text_pass2; 1: k_text_pass2s, the variant specialised for full tiles and constant digit offsets
static bool tp_spec() {
  static int v = -1;
  if (v < 0) { const char *ev = getenv("LBZ_TP_SPEC"); v = ev ? (atoi(ev) != 0) : 0; }
  return v != 0;
}
// ranking inside a warp: 0 = counter load / store per item, 1 = matches first + shared-memory atomics
static int tp_rankv() {
  static int v = -1;
  if (v < 0) { const char *ev = getenv("LBZ_TP_RANK"); v = ev ? atoi(ev) : 0; if (v < 0 || v > 1) v = 0; }
  return v;
}
static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int MODE, int LAST, int LIST>
static int launch_pass3(uint32_t tiles, uint32_t nb, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta, const uint8_t *T,
                        const uint2 *src, uint2 *dst, uint32_t *sa_out, uint32_t *tstat, const uint32_t *gbase, uint32_t gstride,
                        uint32_t shift, uint32_t epoch, uint32_t *err, uint32_t koff, const BwtBuffers &B) {
  static bool attr_set = false;
  if (!attr_set) {
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass3<MODE, LAST, LIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)));
    attr_set = true;
  }
  const uint32_t total = tiles * nb;                       // upper bound of the work list
  if (total == 0) return 0;
  const uint32_t grid = min(total, 2u * (uint32_t)sm_count());
  k_text_pass3<MODE, LAST, LIST><<<grid, P3_THREADS, sizeof(PassSmem), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase, shift, epoch,
                                                                            err, koff, gstride, B.tickets + (epoch & 1023u),
                                                                            B.wl + (LIST ? B.wl_list_off : 0u), B.wl_count + LIST);
  return 0;
}

template <int MODE, int LAST, int LIST>
static int launch_pass4(uint32_t tiles, uint32_t nb, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta, const uint8_t *T,
                        const uint2 *src, uint2 *dst, uint32_t *sa_out, uint32_t *tstat, const uint32_t *gbase, uint32_t gstride,
                        uint32_t shift, uint32_t epoch, uint32_t *err, uint32_t koff) {
  static bool attr_set = false;
  if (!attr_set) {
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass4<MODE, LAST, LIST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Pass4Smem)));
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass4<MODE, LAST, LIST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Pass4Smem)));
    attr_set = true;
  }
  if (tiles == 0 || nb == 0) return 0;
  static int rr = -1;
  if (rr < 0) { const char *ev = getenv("LBZ_TP_RR"); rr = ev ? (atoi(ev) != 0) : 1; }
  if (rr)
    k_text_pass4<MODE, LAST, LIST, true><<<dim3(nb, tiles), 512, sizeof(Pass4Smem), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase, shift,
                                                                                       epoch, err, koff, gstride);
  else
    k_text_pass4<MODE, LAST, LIST, false><<<dim3(nb, tiles), 512, sizeof(Pass4Smem), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase, shift,
                                                                                        epoch, err, koff, gstride);
  return 0;
}

template <int MODE, int LAST, int MINB>
static int launch_text_pass_b(uint32_t nb, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta, const uint8_t *T,
                            const uint2 *src, uint2 *dst, uint32_t *sa_out, uint32_t *tstat, const uint32_t *gbase,
                            uint32_t shift, uint32_t epoch, uint32_t *err, uint32_t koff) {
  if (tp_version() == 1) {
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass<MODE, LAST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TextSmem)));
    k_text_pass<MODE, LAST, MINB><<<dim3(g.S1 / 4096u, nb), 512, sizeof(TextSmem), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase,
                                                                                          shift, epoch, err, koff);
    return 0;
  }
  const uint32_t xp = tp_xpose();
  static int ggrp = -1;
  if (ggrp < 0) { const char *ev = getenv("LBZ_TP_GATHER_GROUP"); ggrp = ev ? atoi(ev) : 32; if (ggrp < 1) ggrp = 1; }
  const uint32_t gsz = (MODE == 1) ? min(nb, (uint32_t)ggrp) : nb;
  const dim3 grid = xp ? dim3(gsz, g.S1 / 4096u, (nb + gsz - 1) / gsz) : dim3(g.S1 / 4096u, nb);
#define LBZ_TP2_LAUNCH_R(SH, RV)                                                                                            \
  do {                                                                                                                      \
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass2s<MODE, LAST, MINB, 0, SH, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)sizeof(TextSmem2)));                                                           \
    k_text_pass2s<MODE, LAST, MINB, 0, SH, RV><<<grid, 512, sizeof(TextSmem2), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase, \
                                                                                    shift, epoch, err, koff, 256u, xp, nb);  \
  } while (0)
#define LBZ_TP2_LAUNCH(SH)                                                                                                  \
  do {                                                                                                                      \
    if (SH >= 0 && tp_rankv() == 1) LBZ_TP2_LAUNCH_R(SH, (SH >= 0 ? 1 : 0));                                                \
    else LBZ_TP2_LAUNCH_R(SH, 0);                                                                                           \
  } while (0)
  // the digit offsets of the default sort depth (8 bytes) as compile-time constants: the passes that
  // build keys from the text (MODE 1, 2) work on the lowest digit, the last pass on the highest
  if (!tp_spec()) {
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass2<MODE, LAST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TextSmem2)));
    if (MINB == 4)
      LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass2<MODE, LAST, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    k_text_pass2<MODE, LAST, MINB><<<grid, 512, sizeof(TextSmem2), st>>>(g, meta, T, src, dst, sa_out, tstat, gbase,
                                                                          shift, epoch, err, koff, 256u, xp, nb);
    return 0;
  }
  if (MINB == 3 && tp_const_shift()) {
    if constexpr (MODE != 0 && !LAST) {
      if (shift == 0u) { LBZ_TP2_LAUNCH(0); return 0; }
    } else if constexpr (MODE == 0 && LAST) {
      if (shift == 24u) { LBZ_TP2_LAUNCH(24); return 0; }
    } else if constexpr (MODE == 0 && !LAST) {
      if (shift == 8u) { LBZ_TP2_LAUNCH(8); return 0; }
      if (shift == 16u) { LBZ_TP2_LAUNCH(16); return 0; }
      if (shift == 24u) { LBZ_TP2_LAUNCH(24); return 0; }
    }
  }
  LBZ_TP2_LAUNCH(-1);
#undef LBZ_TP2_LAUNCH
#undef LBZ_TP2_LAUNCH_R
  return 0;
}
template <int MODE, int LAST>
static int launch_text_pass(uint32_t nb, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta, const uint8_t *T,
                            const uint2 *src, uint2 *dst, uint32_t *sa_out, uint32_t *tstat, const uint32_t *gbase,
                            uint32_t shift, uint32_t epoch, uint32_t *err, uint32_t koff, const BwtBuffers &B) {
  if (tp_version() == 4)
    return launch_pass4<MODE, LAST, 0>(g.S1 / 4096u, nb, st, g, meta, T, src, dst, sa_out, tstat, gbase, 256u, shift, epoch, err, koff);
  if (tp_version() == 3)
    return launch_pass3<MODE, LAST, 0>(g.S1 / 4096u, nb, st, g, meta, T, src, dst, sa_out, tstat, gbase, 256u, shift, epoch, err, koff, B);
  static int minb = 0;
  if (!minb) { const char *ev = getenv("LBZ_TP_MINB"); minb = (ev && atoi(ev) == 2) ? 2 : (ev && atoi(ev) == 4) ? 4 : 3; }
  // 4 CTAs per SM: 32 registers per thread (spills) against 64 instead of 48 resident warps; needs the largest shared-memory carve-out
  if (minb == 4) return launch_text_pass_b<MODE, LAST, 4>(nb, st, g, meta, T, src, dst, sa_out, tstat, gbase, shift, epoch, err, koff);
  if (minb == 3) return launch_text_pass_b<MODE, LAST, 3>(nb, st, g, meta, T, src, dst, sa_out, tstat, gbase, shift, epoch, err, koff);
  return launch_text_pass_b<MODE, LAST, 2>(nb, st, g, meta, T, src, dst, sa_out, tstat, gbase, shift, epoch, err, koff);
}

// One pass over the 32-bit-key pairs of every block's list L.
static int launch_list_pass(uint32_t max_count, uint32_t nb, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta,
                            const uint2 *src, uint2 *dst, uint32_t *tstat, const uint32_t *gbase, uint32_t gstride,
                            uint32_t shift, uint32_t epoch, uint32_t *err, const BwtBuffers &B) {
  const uint32_t tiles = (max_count + 4095u) / 4096u;
  if (tp_version() == 4)
    return launch_pass4<0, 0, 1>(tiles, nb, st, g, meta, nullptr, src, dst, nullptr, tstat, gbase, gstride, shift, epoch, err, 0u);
  if (tp_version() == 3)
    return launch_pass3<0, 0, 1>(tiles, nb, st, g, meta, nullptr, src, dst, nullptr, tstat, gbase, gstride, shift, epoch, err, 0u, B);
  const uint32_t xp = tp_xpose();
  const dim3 grid = xp ? dim3(nb, tiles) : dim3(tiles, nb);
#define LBZ_TP2_LIST_R(SH, RV)                                                                                            \
  do {                                                                                                                    \
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass2s<0, 0, 3, 1, SH, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        (int)sizeof(TextSmem2)));                                                         \
    k_text_pass2s<0, 0, 3, 1, SH, RV><<<grid, 512, sizeof(TextSmem2), st>>>(g, meta, nullptr, src, dst, nullptr, tstat, gbase, \
                                                                           shift, epoch, err, 0u, gstride, xp, nb);       \
  } while (0)
#define LBZ_TP2_LIST(SH)                                                                                                  \
  do {                                                                                                                    \
    if (SH >= 0 && tp_rankv() == 1) LBZ_TP2_LIST_R(SH, (SH >= 0 ? 1 : 0));                                                \
    else LBZ_TP2_LIST_R(SH, 0);                                                                                           \
  } while (0)
  if (!tp_spec()) {
    LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_text_pass2<0, 0, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TextSmem2)));
    k_text_pass2<0, 0, 3, 1><<<grid, 512, sizeof(TextSmem2), st>>>(g, meta, nullptr, src, dst, nullptr, tstat, gbase, shift, epoch, err, 0u,
                                                                    gstride, xp, nb);
    return 0;
  }
  if (tp_const_shift() && shift == 0u) LBZ_TP2_LIST(0);
  else if (tp_const_shift() && shift == 8u) LBZ_TP2_LIST(8);
  else if (tp_const_shift() && shift == 16u) LBZ_TP2_LIST(16);
  else if (tp_const_shift() && shift == 24u) LBZ_TP2_LIST(24);
  else LBZ_TP2_LIST(-1);
#undef LBZ_TP2_LIST
#undef LBZ_TP2_LIST_R
  return 0;
}

// Digit bases of the text passes: every pass of the initial sort sees the same
// multiset of digits (each rotation index occurs once, so the digits at any
// depth are the bytes of the block), hence one byte histogram per block serves
// all BWT_K passes.  One CTA per block.
__global__ void __launch_bounds__(1024)
k_text_bases(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             uint32_t *__restrict__ gbase) {
  const uint32_t b = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  const uint8_t *Tb = T + lbz_slot_off(g, b);
  __shared__ uint32_t sh[256];
  __shared__ uint32_t ws[40];
  if (threadIdx.x < 256) sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t nvec = n / 16;
  for (uint32_t i = threadIdx.x; i < nvec; i += 1024) {
    const uint4 v = reinterpret_cast<const uint4 *>(Tb)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      atomicAdd(&sh[w[q] & 0xFFu], 1u); atomicAdd(&sh[(w[q] >> 8) & 0xFFu], 1u);
      atomicAdd(&sh[(w[q] >> 16) & 0xFFu], 1u); atomicAdd(&sh[w[q] >> 24], 1u);
    }
  }
  for (uint32_t i = nvec * 16 + threadIdx.x; i < n; i += 1024) atomicAdd(&sh[Tb[i]], 1u);
  __syncthreads();
  const uint32_t v = threadIdx.x < 256 ? sh[threadIdx.x] : 0u;
  uint32_t tot;
  const uint32_t ex = cta_excl_sum(v, ws, &tot);
  if (threadIdx.x < 256) gbase[b * 256 + threadIdx.x] = ex;
}

// Digit bases of the five key passes of a round from the histograms that
// k_round_keys accumulated.
__global__ void __launch_bounds__(256)
k_key_bases(const LbzBlockMeta *__restrict__ meta, const uint32_t *__restrict__ khist, uint32_t *__restrict__ gbase1) {
  const uint32_t b = blockIdx.x;
  if (meta[b].ul == 0) return;
  __shared__ uint32_t ws[40];
  for (uint32_t p = 0; p < 5; p++) {
    const uint32_t v = khist[(b * 5 + p) * 256 + threadIdx.x];
    uint32_t tot;
    const uint32_t ex = cta_excl_sum(v, ws, &tot);
    gbase1[(b * 5 + p) * 256 + threadIdx.x] = ex;
  }
}

// ---------------------------------------------------------------------------
// Group heads after the initial sort: head[p] = first BWT_K bytes of rotation
// sa[p] differ from those of sa[p-1].  Keys are only compared for equality, so
// the 8 bytes are fetched as three aligned words and funnel-shifted.
__device__ __forceinline__ uint64_t text_key(const uint8_t *__restrict__ Tb, uint32_t v, uint32_t n, uint32_t K) {
  const uint64_t kmask = K >= 8u ? ~0ull : ((1ull << (8u * K)) - 1ull);    // byte d of the rotation sits at bits 8d
  if (v + 8u <= n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(Tb + (v & ~3u));
    const uint32_t sh = 8u * (v & 3u);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];      // slot capacity >= n + 64: in bounds
    const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
    return (((uint64_t)x1 << 32) | x0) & kmask;
  }
  uint64_t k = 0;
  uint32_t j = v;
#pragma unroll
  for (uint32_t d = 0; d < 8u; d++) { if (d < K) k |= (uint64_t)Tb[j] << (8 * d); if (++j >= n) j = 0; }
  return k;
}

#define SMALL_GROUP 32u      // groups up to this size are refined by the local sort

struct TileAgg { uint32_t small; uint32_t large; int last; uint32_t lheads; };   // lheads: heads of large tied groups
#define ORD_LIMIT 4096u       // list-L group ordinals below this fit the 12 spare bits of a gs entry / a 32-bit round key

__device__ __forceinline__ void list_sel(const LbzBlockMeta &m, uint32_t sel, uint32_t &base, uint32_t &cnt) {
  if (sel == 0) { base = 0; cnt = m.us; } else { base = m.lbase; cnt = m.ul; }
}

// Distance from bit x back to the nearest set bit at or before x (0..31), or 32 if
// none within 32 bits; distance forward to the nearest set bit after x (1..32), or
// 33 if none.  `bits` is a shared-memory bitmask, x >= 32 and x + 32 in range.
__device__ __forceinline__ uint32_t win_back(const uint32_t *bits, uint32_t x) {
  const uint32_t wi = x >> 5, bi = x & 31u;
  const uint32_t t = (bi == 31u) ? bits[wi] : __funnelshift_r(bits[wi - 1], bits[wi], bi + 1u);
  return t ? (uint32_t)__clz(t) : 32u;
}
__device__ __forceinline__ uint32_t win_fwd(const uint32_t *bits, uint32_t x) {
  const uint32_t wi = x >> 5, bi = x & 31u;
  const uint32_t u = (bi == 31u) ? bits[wi + 1] : __funnelshift_r(bits[wi], bits[wi + 1], bi + 1u);
  return u ? (uint32_t)__ffs(u) : 33u;
}

// Tile-parallel pass A: head flags of a tile (bit 0) plus the size class of every
// tied rotation's group (bit 1: the group has <= SMALL_GROUP members), numbers of
// small/large tied rotations, last head position.  Per-tile aggregates replace a
// serial scan: pass B sums the (<= 220) aggregates of its block itself.
#define HALO 32u
__global__ void __launch_bounds__(256)
k_heads_agg(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
            const uint32_t *__restrict__ sa, uint8_t *__restrict__ head, TileAgg *__restrict__ agg, uint32_t K) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *Tb = T + off;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  // flags of positions tbase-HALO .. tbase+LBZ_TILE+HALO ; index = p - tbase + HALO
  __shared__ __align__(16) uint8_t sflag[HALO + LBZ_TILE + HALO + 16];
  __shared__ uint32_t sbits[(HALO + LBZ_TILE + HALO) / 32 + 2];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
#pragma unroll 2
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {     // uniform trip count: shuffles stay converged
    const uint32_t p = tbase + it * 256 + tid;
    const bool valid = p < n;
    uint64_t k = 0;
    if (valid) k = text_key(Tb, sa[off + p], n, K);
    uint64_t kprev = __shfl_up_sync(0xffffffffu, k, 1);
    bool hd = true;                                      // end of block acts as a head
    if (valid) {
      if (lane == 0) kprev = p ? text_key(Tb, sa[off + p - 1], n, K) : ~k;
      hd = (p == 0) || (k != kprev);
    }
    sflag[HALO + p - tbase] = hd;
    const uint32_t bw = __ballot_sync(0xffffffffu, hd);
    if (lane == 0) sbits[1 + ((p - tbase) >> 5)] = bw;   // HALO == 32: the tile starts at word 1
  }
  if (tid < 2 * HALO + 1) {                              // halo flags on both sides (+ the tile's end flag)
    const int64_t p = (tid < HALO) ? (int64_t)tbase - HALO + tid : (int64_t)tbase + LBZ_TILE + (tid - HALO);
    const uint32_t x = (tid < HALO) ? tid : HALO + LBZ_TILE + (tid - HALO);
    uint8_t f = 1;
    if (p > 0 && p < (int64_t)n)
      f = text_key(Tb, sa[off + (uint32_t)p], n, K) != text_key(Tb, sa[off + (uint32_t)p - 1], n, K);
    sflag[x] = f;
  }
  __syncthreads();
  if (tid < 2) {                                         // pack the two halo words from the flag bytes
    const uint32_t base = tid ? HALO + LBZ_TILE : 0u;
    uint32_t wv = 0;
    for (uint32_t q = 0; q < 32; q++) wv |= (uint32_t)(sflag[base + q] & 1u) << q;
    sbits[tid ? (HALO + LBZ_TILE) / 32 : 0] = wv;
  }
  __syncthreads();
  const bool tracking = K < n;
  const uint32_t q0 = tid * 16;
  uint32_t small = 0, large = 0;
  int last = -1;
  uint32_t outw[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t p = tbase + q0 + j;
    const uint32_t x = HALO + q0 + j;
    uint32_t fb = sflag[x] & 1u;
    if (p < n) {
      const bool f = fb, f1 = sflag[x + 1] & 1u;
      if (f) last = (int)p;
      if (tracking && !(f && f1)) {
        const uint32_t back = win_back(sbits, x), fwd = win_fwd(sbits, x);   // group = [x-back, x+fwd)
        if (back + fwd <= SMALL_GROUP) { small++; fb |= 2u; } else large++;
      }
    } else {
      fb = 1;
    }
    outw[j >> 2] |= fb << (8 * (j & 3));
  }
  *reinterpret_cast<uint4 *>(&head[off + tbase + q0]) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
  uint32_t tots, totl;
  int tmax;
  (void)cta_excl_sum(small, ws, &tots);
  (void)cta_excl_sum(large, ws, &totl);
  (void)cta_excl_max(last, -1, wsi, &tmax);
  if (tid == 0) { TileAgg a; a.small = tots; a.large = totl; a.last = tmax; a.lheads = 0; agg[(size_t)b * g.tiles1 + tile] = a; }
}

// Tile-parallel pass B: ranks (rank[i] = first position of i's group) and
// compaction of the members of tied groups into the two round lists.
__global__ void __launch_bounds__(256, 5)
k_ranks_compact(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B, const TileAgg *__restrict__ agg) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t tid = threadIdx.x;
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  // carries from the preceding tiles of this block; list L starts where S ends
  const uint32_t ntiles = (n + LBZ_TILE - 1) / LBZ_TILE;
  uint32_t cs = 0, cl = 0, alls = 0, alll = 0;
  int l = -1;
  for (uint32_t t = tid; t < ntiles; t += 256) {
    const TileAgg a = agg[(size_t)b * g.tiles1 + t];
    alls += a.small; alll += a.large;
    if (t < tile) { cs += a.small; cl += a.large; l = max(l, a.last); }
  }
  uint32_t carry_s, carry_l, total_s, total_l;
  int carry_last;
  (void)cta_excl_sum(cs, ws, &carry_s);
  (void)cta_excl_sum(cl, ws, &carry_l);
  (void)cta_excl_sum(alls, ws, &total_s);
  (void)cta_excl_sum(alll, ws, &total_l);
  (void)cta_excl_max(l, -1, wsi, &carry_last);
  const uint32_t lbase = total_s;

  const uint32_t p0 = tbase + tid * 16;
  uint8_t f[17];
  {
    const uint4 fv = *reinterpret_cast<const uint4 *>(&B.head[off + p0]);
    const uint32_t wv[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
    for (int j = 0; j < 16; j++) f[j] = (p0 + j < n) ? (uint8_t)((wv[j >> 2] >> (8 * (j & 3))) & 0xFFu) : (uint8_t)1;
    f[16] = (p0 + 16 < n) ? B.head[off + p0 + 16] : (uint8_t)1;
  }
  int last = -1;
#pragma unroll
  for (int j = 0; j < 16; j++) if (p0 + j < n && (f[j] & 1u)) last = (int)(p0 + j);
  int tmax;
  int st = cta_excl_max(last, -1, wsi, &tmax);
  st = max(st, carry_last);
  const bool tracking = B.K < n;
  uint32_t smask = 0, lmask = 0;
  uint32_t starts[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (p0 + j < n) {
      if (f[j] & 1u) st = (int)(p0 + j);
      starts[j] = (uint32_t)st;
      if (tracking && !((f[j] & 1u) && (f[j + 1] & 1u))) {
        if (f[j] & 2u) smask |= 1u << j; else lmask |= 1u << j;
      }
    }
  }
  uint32_t tots, totl;
  uint32_t os = carry_s + cta_excl_sum(__popc(smask), ws, &tots);
  uint32_t ol = lbase + carry_l + cta_excl_sum(__popc(lmask), ws, &totl);
  if (p0 < n) {
    uint32_t v[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 x = *reinterpret_cast<const uint4 *>(&B.sa[off + p0 + 4 * q]);   // slot capacity covers the overread
      v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
    const uint64_t pol = l2_policy_evict_last();
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (p0 + j < n) {
        if (B.hints) st_u32_hint(&B.rank[off + v[j]], starts[j], pol); else B.rank[off + v[j]] = starts[j];
        if ((smask | lmask) & (1u << j)) {
          const uint32_t o = (smask & (1u << j)) ? os++ : ol++;
          if (B.hints) {
            __stcs(&B.pos[off + o], p0 + j); __stcs(&B.val[off + o], v[j]); __stcs(&B.gs[off + o], starts[j]);
          } else {
            B.pos[off + o] = p0 + j; B.val[off + o] = v[j]; B.gs[off + o] = starts[j];
          }
        }
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= n) {               // last tile of the block
    meta[b].us_next = total_s;                            // committed by k_round_commit
    meta[b].ul_next = total_l;
    meta[b].lbase = lbase;
    meta[b].lbase_next = lbase;
    meta[b].depth = B.K;
    atomicMax(&B.counters[0], max(total_s, total_l));
    atomicAdd(&B.counters[1], total_s + total_l);
  }
}

// ---------------------------------------------------------------------------
// Round key: (group start << 20) | rank of rotation (i + h) for one of the two
// lists; for list L it also accumulates the digit histograms of the five radix
// passes that follow.
__global__ void __launch_bounds__(256)
k_round_keys(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint32_t *__restrict__ val,
             const uint32_t *__restrict__ gs, const uint32_t *__restrict__ rank,
             uint64_t *__restrict__ key, uint32_t *__restrict__ khist, uint32_t h, uint32_t sel, int hints, int pair) {
  const uint64_t pol = l2_policy_evict_last();
  const uint32_t b = blockIdx.y;
  uint32_t lb, U;
  list_sel(meta[b], sel, lb, U);
  const uint32_t tbase = blockIdx.x * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t lo = off + lb;
  __shared__ uint32_t sh[5][256];
  if (sel) {
    for (uint32_t i = threadIdx.x; i < 5 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
  }
  uint32_t v[LBZ_TILE / 256], gg[LBZ_TILE / 256];
#pragma unroll
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t j = tbase + it * 256 + threadIdx.x;
    v[it] = (j < U) ? val[lo + j] : 0xFFFFFFFFu;
    gg[it] = (j < U) ? gs[lo + j] : 0u;
  }
#pragma unroll
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    if (v[it] != 0xFFFFFFFFu) v[it] = hints ? ld_u32_hint(&rank[off + wrap_add(v[it], h, n)], pol) : rank[off + wrap_add(v[it], h, n)];
  }
#pragma unroll
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t j = tbase + it * 256 + threadIdx.x;
    if (j < U) {
      if (sel && pair) {
        // list L, 32-bit form: (group ordinal, rank of rotation i+h) next to the index
        const uint32_t k32 = (gg[it] & 0xFFF00000u) | v[it];
        reinterpret_cast<uint2 *>(key)[lo + j] = make_uint2(k32, val[lo + j]);
#pragma unroll
        for (int p = 0; p < 4; p++) atomicAdd(&sh[p][(k32 >> (8 * p)) & 0xFFu], 1u);
      } else {
        const uint64_t k = ((uint64_t)(gg[it] & 0xFFFFFu) << 20) | v[it];
        key[lo + j] = k;
        if (sel) {
#pragma unroll
          for (int p = 0; p < 5; p++) atomicAdd(&sh[p][(uint32_t)(k >> (8 * p)) & 0xFFu], 1u);
        }
      }
    }
  }
  if (sel) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 5 * 256; i += 256) {
      const uint32_t c = (&sh[0][0])[i];
      if (c) atomicAdd(&khist[(size_t)b * 5 * 256 + i], c);
    }
  }
}

// Local refinement of list S: every group has at most SMALL_GROUP members, so a
// member's place inside its group is found by counting the members with a smaller
// key (ties by list index: stable).  One tile of 4096 list entries per CTA with a
// halo of SMALL_GROUP entries on both sides staged in shared memory.
__global__ void __launch_bounds__(256)
k_small_sort(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint64_t *__restrict__ kin,
             const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout, uint32_t *__restrict__ vout) {
  const uint32_t b = blockIdx.y;
  const uint32_t U = meta[b].us;
  const uint32_t tbase = blockIdx.x * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t off = lbz_slot_off(g, b);
  constexpr uint32_t W = SMALL_GROUP + LBZ_TILE + SMALL_GROUP;        // staged window, SMALL_GROUP == 32
  __shared__ uint32_t srk[W];                                          // rank part of the key (20 bits)
  __shared__ uint32_t sbits[W / 32 + 2];                               // group-head bits
  const int64_t lo = (int64_t)tbase - SMALL_GROUP;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t x0 = 0; x0 < W; x0 += 256) {                           // W is a multiple of 32: whole warps
    const uint32_t x = x0 + threadIdx.x;
    bool hd = true;
    if (x < W) {
      const int64_t j = lo + x;
      if (j >= 0 && j < (int64_t)U) {
        const uint64_t k = kin[off + j];
        srk[x] = (uint32_t)k & 0xFFFFFu;
        hd = (j == 0) || ((kin[off + j - 1] >> 20) != (k >> 20));
      } else {
        srk[x] = 0;
      }
    }
    const uint32_t bw = __ballot_sync(0xffffffffu, hd);
    if (lane == 0 && x < W) sbits[x >> 5] = bw;
  }
  if (threadIdx.x == 0) sbits[W / 32] = 0xFFFFFFFFu;
  __syncthreads();
  for (uint32_t q = threadIdx.x; q < LBZ_TILE; q += 256) {
    const uint32_t j = tbase + q;
    if (j >= U) break;
    const uint32_t x = SMALL_GROUP + q;
    const uint32_t hs = x - win_back(sbits, x), he = x + win_fwd(sbits, x);   // group = [hs, he), <= 32 members
    const uint32_t k = srk[x];
    uint32_t r = 0;
    for (uint32_t y = hs; y < he; y++) { const uint32_t ky = srk[y]; r += (ky < k) || (ky == k && y < x); }
    const uint32_t dst = (uint32_t)(lo + hs) + r;                       // list index of the sorted place
    kout[off + dst] = kin[off + j];
    vout[off + dst] = vin[off + j];
  }
}

// After the round sort (tile-parallel, two passes like the initial ranks):
// pass A counts, per tile of the sorted list, the elements that stay tied and
// finds the last group head; pass B refines groups, writes order and ranks
// back and compacts the survivors for the next round.
__device__ __forceinline__ void load_round_keys(const uint64_t *__restrict__ skey, uint32_t off, uint32_t j0,
                                                uint32_t U, uint64_t k[18]) {
  // k[q] = key of list index j0 - 1 + q ; out of range = ~0 (never a real key: 40 bits)
#pragma unroll
  for (int q = 0; q < 18; q++) {
    const int64_t j = (int64_t)j0 + q - 1;
    k[q] = (j >= 0 && j < (int64_t)U) ? skey[off + j] : ~0ull;
  }
}

__global__ void __launch_bounds__(256)
k_round_agg(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint64_t *__restrict__ skey,
            TileAgg *__restrict__ agg, uint32_t h, uint32_t sel) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  uint32_t lb, U;
  list_sel(meta[b], sel, lb, U);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b) + lb;
  const uint32_t tid = threadIdx.x;
  const bool more = (2u * h < n);
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  const uint32_t j0 = tbase + tid * 16;
  uint64_t k[18];
  load_round_keys(skey, off, j0, U, k);
  uint32_t uns = 0;
  int last = -1;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const uint32_t j = j0 + q;
    if (j < U) {
      const bool hd = (j == 0) || (k[q + 1] != k[q]);
      const bool hn = (j + 1 >= U) || (k[q + 2] != k[q + 1]);
      if (hd) last = (int)j;
      uns += (more && !(hd && hn));
    }
  }
  uint32_t tot;
  int tmax;
  (void)cta_excl_sum(uns, ws, &tot);
  (void)cta_excl_max(last, -1, wsi, &tmax);
  if (tid == 0) {
    TileAgg a; a.small = tot; a.large = 0; a.last = tmax; a.lheads = 0;
    agg[((size_t)sel * gridDim.y + b) * g.tiles1 + tile] = a;
  }
}

__global__ void __launch_bounds__(256, 3)
k_round_apply(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B,
              const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sval,
              const uint32_t *__restrict__ pos, uint32_t *__restrict__ nval,
              uint32_t *__restrict__ npos, uint32_t *__restrict__ ngs,
              const TileAgg *__restrict__ agg, uint32_t h, uint32_t sel) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  uint32_t lb, U;
  list_sel(meta[b], sel, lb, U);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t sa_off = lbz_slot_off(g, b);       // order / rank arrays
  const uint32_t off = sa_off + lb;                  // list arrays
  const uint32_t tid = threadIdx.x;
  const bool more = (2u * h < n);
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  uint32_t c = 0;
  int l = -1;
  for (uint32_t t = tid; t < tile; t += 256) {
    const TileAgg a = agg[((size_t)sel * gridDim.y + b) * g.tiles1 + t];
    c += a.small; l = max(l, a.last);
  }
  uint32_t carry_cnt;
  int carry_last;
  (void)cta_excl_sum(c, ws, &carry_cnt);
  (void)cta_excl_max(l, -1, wsi, &carry_last);

  const uint32_t j0 = tbase + tid * 16;
  uint64_t k[18];
  load_round_keys(skey, off, j0, U, k);
  int last = -1;
  uint32_t hdmask = 0;
#pragma unroll
  for (int q = 0; q < 17; q++) {
    const uint32_t j = j0 + q;
    const bool hd = (j >= U) || (j == 0) || (k[q + 1] != k[q]);
    if (hd) hdmask |= 1u << q;
    if (q < 16 && j < U && hd) last = (int)j;
  }
  int tmax;
  int st = cta_excl_max(last, -1, wsi, &tmax);
  st = max(st, carry_last);
  uint32_t unsmask = 0;
  uint32_t mygs[16], myp[16], myv[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const uint32_t j = j0 + q;
    if (j < U) {
      if (hdmask & (1u << q)) st = (int)j;
      myp[q] = pos[off + j];
      myv[q] = sval[off + j];
      mygs[q] = pos[off + (uint32_t)st];                  // SA position of the group's head element
      const bool single = (hdmask & (1u << q)) && (hdmask & (2u << q));
      if (more && !single) unsmask |= 1u << q;
    }
  }
  uint32_t tot;
  uint32_t o = carry_cnt + cta_excl_sum(__popc(unsmask), ws, &tot);
  const uint64_t pol = l2_policy_evict_last();
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const uint32_t j = j0 + q;
    if (j < U) {
      // the order entry is only final once the rotation leaves the tied lists; the rank
      // only changes for rotations that are not in the first subgroup of their old group
      if (!(unsmask & (1u << q))) B.sa[sa_off + myp[q]] = myv[q];
      if (mygs[q] != (uint32_t)(k[q + 1] >> 20)) {
        if (B.hints) st_u32_hint(&B.rank[sa_off + myv[q]], mygs[q], pol); else B.rank[sa_off + myv[q]] = mygs[q];
      }
      if (unsmask & (1u << q)) {
        if (B.hints) { __stcs(&npos[off + o], myp[q]); __stcs(&nval[off + o], myv[q]); __stcs(&ngs[off + o], mygs[q]); }
        else { npos[off + o] = myp[q]; nval[off + o] = myv[q]; ngs[off + o] = mygs[q]; }
        o++;
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= U) {                // last tile of this list
    const uint32_t U2 = carry_cnt + tot;
    if (sel) meta[b].ul_next = U2; else meta[b].us_next = U2;   // committed by k_round_commit
    meta[b].depth = 2u * h;
    atomicMax(&B.counters[0], U2);
    atomicAdd(&B.counters[1], U2);
    if (!sel) atomicAdd(&B.counters[4], U2);              // diagnostics only (LBZ_ROUND_STATS)
  }
}


// ===========================================================================
// Bitmask formulation of the rank/refine passes.  Group structure is kept as one
// bit per position (head bit) so that group starts, "still tied" flags and
// compaction offsets come from popcounts and find-highest-bit on shared-memory
// words, and every global access is striped: consecutive lanes touch consecutive
// elements (the earlier kernels gave each thread 16 consecutive elements, which
// made every load and store instruction touch 32 different cache lines).
__device__ __forceinline__ uint32_t lanemask_le() { return lanemask_lt() | (1u << (threadIdx.x & 31u)); }
__device__ __forceinline__ uint32_t valid_bits(uint32_t base, uint32_t n) {     // bits of word [base, base+32) below n
  if (base >= n) return 0u;
  const uint32_t c = n - base;
  return c >= 32u ? 0xFFFFFFFFu : ((1u << c) - 1u);
}

// Tile-parallel pass A (replaces k_heads_agg): head bits + size-class bits of a tile
// as bitmaps, numbers of small/large tied rotations, last head position.
__global__ void __launch_bounds__(256)
k_heads_bits(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
             const uint32_t *__restrict__ sa, uint32_t *__restrict__ hbits, uint32_t *__restrict__ cbits,
             TileAgg *__restrict__ agg, uint32_t K) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *Tb = T + off;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  // sbits[0] = positions tbase-32..tbase-1, sbits[1..128] = the tile, sbits[129] = the 32 positions after it
  __shared__ uint32_t sbits[132];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
#pragma unroll 2
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t p = tbase + it * 256 + tid;
    const bool valid = p < n;
    uint64_t k = 0;
    if (valid) k = text_key(Tb, sa[off + p], n, K);
    uint64_t kprev = __shfl_up_sync(0xffffffffu, k, 1);
    bool hd = true;                                      // past the end of the block: head
    if (valid) {
      if (lane == 0) kprev = p ? text_key(Tb, sa[off + p - 1], n, K) : ~k;
      hd = (p == 0) || (k != kprev);
    }
    const uint32_t bw = __ballot_sync(0xffffffffu, hd);
    if (lane == 0) sbits[1 + it * 8 + warp] = bw;
  }
  if (tid < 64) {                                        // two whole warps: the halo words
    const int64_t p = (warp == 0) ? (int64_t)tbase - 32 + lane : (int64_t)tbase + LBZ_TILE + lane;
    bool f = true;
    if (p > 0 && p < (int64_t)n)
      f = text_key(Tb, sa[off + (uint32_t)p], n, K) != text_key(Tb, sa[off + (uint32_t)p - 1], n, K);
    const uint32_t bw = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sbits[warp == 0 ? 0 : 129] = bw;
  }
  if (tid == 64) { sbits[130] = 0xFFFFFFFFu; sbits[131] = 0xFFFFFFFFu; }
  __syncthreads();
  const bool tracking = K < n;
  const uint32_t wordbase = (off + tbase) >> 5;
  uint32_t small = 0, large = 0, lheads = 0;
  int last = -1;
#pragma unroll 2
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t w = it * 8 + warp;
    const uint32_t p = tbase + w * 32 + lane;
    const uint32_t x = HALO + w * 32 + lane;
    const uint32_t H = sbits[1 + w];
    const uint32_t Hn = (H >> 1) | (sbits[2 + w] << 31);
    const bool tied = (p < n) && tracking && !(((H >> lane) & 1u) && ((Hn >> lane) & 1u));
    bool sm = false;
    if (tied) sm = (win_back(sbits, x) + win_fwd(sbits, x)) <= SMALL_GROUP;   // group = [x-back, x+fwd)
    const uint32_t bt = __ballot_sync(0xffffffffu, tied);
    const uint32_t bs = __ballot_sync(0xffffffffu, sm);
    if (lane == 0) {
      hbits[wordbase + w] = H;
      cbits[wordbase + w] = bs;
      small += __popc(bs);
      large += __popc(bt & ~bs);
      lheads += __popc(bt & ~bs & H);
      const uint32_t hv = H & valid_bits(tbase + w * 32, n);
      if (hv) last = max(last, (int)(tbase + w * 32 + 31u - (uint32_t)__clz(hv)));
    }
  }
  uint32_t tots, totl, toth;
  int tmax;
  (void)cta_excl_sum(small, ws, &tots);
  (void)cta_excl_sum(large, ws, &totl);
  (void)cta_excl_sum(lheads, ws, &toth);
  (void)cta_excl_max(last, -1, wsi, &tmax);
  if (tid == 0) { TileAgg a; a.small = tots; a.large = totl; a.last = tmax; a.lheads = toth; agg[(size_t)b * g.tiles1 + tile] = a; }
}

// Tile-parallel pass B (replaces k_ranks_compact): ranks and compaction of the tied
// rotations into the two round lists, from the bitmaps.
__global__ void __launch_bounds__(256, 4)
k_ranks_compact2(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B, const TileAgg *__restrict__ agg) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  __shared__ uint32_t hb[130], sb[128], lbm[128], spre[128], lpre[128], hpre[128];
  __shared__ int lastpre[128];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  const uint32_t ntiles = (n + LBZ_TILE - 1) / LBZ_TILE;
  uint32_t cs = 0, cl = 0, ch = 0, alls = 0, alll = 0, allh = 0;
  int l = -1;
  for (uint32_t t = tid; t < ntiles; t += 256) {
    const TileAgg a = agg[(size_t)b * g.tiles1 + t];
    alls += a.small; alll += a.large; allh += a.lheads;
    if (t < tile) { cs += a.small; cl += a.large; ch += a.lheads; l = max(l, a.last); }
  }
  uint32_t carry_s, carry_l, carry_h, total_s, total_l, total_h;
  int carry_last;
  (void)cta_excl_sum(cs, ws, &carry_s);
  (void)cta_excl_sum(cl, ws, &carry_l);
  (void)cta_excl_sum(ch, ws, &carry_h);
  (void)cta_excl_sum(alls, ws, &total_s);
  (void)cta_excl_sum(alll, ws, &total_l);
  (void)cta_excl_sum(allh, ws, &total_h);
  (void)cta_excl_max(l, -1, wsi, &carry_last);
  const uint32_t lbase = total_s;
  const uint32_t wordbase = (off + tbase) >> 5;
  if (tid < 128) hb[tid] = B.hbits[wordbase + tid];
  if (tid == 128) { hb[128] = (tbase + LBZ_TILE >= n) ? 1u : B.hbits[wordbase + 128]; hb[129] = 0u; }
  __syncthreads();
  const bool tracking = B.K < n;
  uint32_t sw = 0, lw = 0;
  int lastw = -1;
  if (tid < 128) {
    const uint32_t H = hb[tid], Hn = (H >> 1) | (hb[tid + 1] << 31);
    const uint32_t V = valid_bits(tbase + tid * 32, n);
    const uint32_t tied = tracking ? (~(H & Hn) & V) : 0u;
    const uint32_t cw = B.cbits[wordbase + tid];
    sw = tied & cw; lw = tied & ~cw;
    sb[tid] = sw; lbm[tid] = lw;
    if (H & V) lastw = (int)(tbase + tid * 32 + 31u - (uint32_t)__clz(H & V));
  }
  uint32_t tots, totl;
  int tmax;
  const uint32_t exs = cta_excl_sum(__popc(sw), ws, &tots);
  const uint32_t exl = cta_excl_sum(__popc(lw), ws, &totl);
  uint32_t toth;
  const uint32_t exh = cta_excl_sum(tid < 128 ? __popc(lw & hb[tid]) : 0u, ws, &toth);
  const int exm = cta_excl_max(lastw, -1, wsi, &tmax);
  if (tid < 128) { spre[tid] = exs; lpre[tid] = exl; hpre[tid] = exh; lastpre[tid] = max(exm, carry_last); }
  __syncthreads();
  const uint32_t lt = lanemask_lt(), le = lanemask_le();
  const uint64_t pol = l2_policy_evict_last();
#pragma unroll 4
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t w = it * 8 + warp;
    const uint32_t p = tbase + w * 32 + lane;
    if (p < n) {
      const uint32_t v = B.sa[off + p];
      const uint32_t m = hb[w] & le;
      const uint32_t st = m ? (tbase + w * 32 + 31u - (uint32_t)__clz(m)) : (uint32_t)lastpre[w];
      if (B.hints) st_u32_hint(&B.rank[off + v], st, pol); else B.rank[off + v] = st;
      const uint32_t sm = sb[w], lm = lbm[w];
      if ((sm >> lane) & 1u) {
        const uint32_t o = carry_s + spre[w] + __popc(sm & lt);
        B.pos[off + o] = p; B.val[off + o] = v; B.gs[off + o] = st;
      } else if ((lm >> lane) & 1u) {
        const uint32_t o = lbase + carry_l + lpre[w] + __popc(lm & lt);
        // ordinal of the group inside list L (heads of large groups at or before p, minus one)
        const uint32_t ord = min(carry_h + hpre[w] + __popc(lm & hb[w] & le) - 1u, ORD_LIMIT - 1u);
        B.pos[off + o] = p; B.val[off + o] = v; B.gs[off + o] = st | (ord << 20);
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= n) {               // last tile of the block
    meta[b].us_next = total_s;                            // committed by k_round_commit
    meta[b].ul_next = total_l;
    meta[b].lbase = lbase;
    meta[b].lbase_next = lbase;
    meta[b].depth = B.K;
    atomicMax(&B.counters[0], max(total_s, total_l));
    atomicAdd(&B.counters[1], total_s + total_l);
    atomicAdd(&B.counters[4], total_s);                   // diagnostics only (LBZ_ROUND_STATS)
    atomicMax(&B.counters[5], total_h);                   // groups in list L: decides the 32-bit key path
  }
}

// Head bits of one tile of a sorted round list with a 32-position halo on both sides,
// in the layout win_back()/win_fwd() expect: sb[0] = list indices tbase-32..tbase-1,
// sb[1..128] = the tile, sb[129] = the 32 indices after it, sb[130..131] = all ones.
// Bit = key differs from its predecessor's, or the index starts / lies beyond the list.
// PAIR: the list holds (32-bit key, index) pairs in the same 8-byte slots; only the key half is compared.
template <int PAIR>
__device__ __forceinline__ uint64_t round_key_at(const uint64_t *__restrict__ skey, uint32_t idx) {
  if (PAIR) return (uint64_t)reinterpret_cast<const uint2 *>(skey)[idx].x;
  return skey[idx];
}
template <int PAIR>
__device__ __forceinline__ void round_head_bits(const uint64_t *__restrict__ skey, uint32_t off, uint32_t tbase,
                                                uint32_t U, uint32_t *sb) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#pragma unroll 4
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t j = tbase + it * 256 + tid;
    const uint64_t k = (j < U) ? round_key_at<PAIR>(skey, off + j) : ~0ull;      // real keys have <= 40 bits
    uint64_t kprev = __shfl_up_sync(0xffffffffu, k, 1);
    if (lane == 0) kprev = (j > 0 && j < U) ? round_key_at<PAIR>(skey, off + j - 1) : ~0ull;
    const bool hd = (j >= U) || (j == 0) || (k != kprev);
    const uint32_t bw = __ballot_sync(0xffffffffu, hd);
    if (lane == 0) sb[1 + it * 8 + warp] = bw;
  }
  if (tid < 64) {                                        // two whole warps: the halo words
    const int64_t j = (warp == 0) ? (int64_t)tbase - 32 + lane : (int64_t)tbase + LBZ_TILE + lane;
    bool f = true;
    if (j > 0 && j < (int64_t)U)
      f = round_key_at<PAIR>(skey, off + (uint32_t)j) != round_key_at<PAIR>(skey, off + (uint32_t)j - 1);
    const uint32_t bw = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sb[warp == 0 ? 0 : 129] = bw;
  }
  if (tid == 64) { sb[130] = 0xFFFFFFFFu; sb[131] = 0xFFFFFFFFu; }
  __syncthreads();
}

// "Still tied" bits of the tile, split by the size of the (refined) group: mv = members of
// groups of <= SMALL_GROUP (they belong in list S from now on), st = members of larger
// groups.  CLS = 0 (list S: groups only shrink) skips the size test.  One word per
// (it, warp), written by lane 0; the caller synchronises.
template <int CLS>
__device__ __forceinline__ void round_class_bits(const uint32_t *sb, uint32_t tbase, uint32_t U, bool more,
                                                 uint32_t *mvb, uint32_t *stb) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#pragma unroll 2
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t w = it * 8 + warp;
    const uint32_t H = sb[1 + w], Hn = (H >> 1) | (sb[2 + w] << 31);
    const uint32_t tied = more ? (~(H & Hn) & valid_bits(tbase + w * 32, U)) : 0u;
    uint32_t mv = tied, st = 0u;
    if (CLS) {
      const uint32_t x = 32u + w * 32 + lane;
      bool sm = false;
      if ((tied >> lane) & 1u) sm = (win_back(sb, x) + win_fwd(sb, x)) <= SMALL_GROUP;
      mv = __ballot_sync(0xffffffffu, sm);
      st = tied & ~mv;
    }
    if (lane == 0) { mvb[w] = mv; stb[w] = st; }
  }
}

// Pass A of a round (per list): per tile, the numbers of members that stay tied in
// small / large groups, the heads of the large ones, the last group head.
template <int PAIR, int CLS>
__global__ void __launch_bounds__(256)
k_round_agg2(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint64_t *__restrict__ skey,
             TileAgg *__restrict__ agg, uint32_t h, uint32_t sel) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  uint32_t lb, U;
  list_sel(meta[b], sel, lb, U);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b) + lb;
  const uint32_t tid = threadIdx.x;
  const bool more = (2u * h < n);
  __shared__ uint32_t sb[132], mvb[128], stb[128];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  round_head_bits<PAIR>(skey, off, tbase, U, sb);
  round_class_bits<CLS>(sb, tbase, U, more, mvb, stb);
  __syncthreads();
  uint32_t nmv = 0, nst = 0, nhd = 0;
  int last = -1;
  if (tid < 128) {
    const uint32_t H = sb[1 + tid];
    const uint32_t V = valid_bits(tbase + tid * 32, U);
    nmv = __popc(mvb[tid]); nst = __popc(stb[tid]); nhd = __popc(stb[tid] & H);
    if (H & V) last = (int)(tbase + tid * 32 + 31u - (uint32_t)__clz(H & V));
  }
  uint32_t tmv, tst, thd;
  int tmax;
  (void)cta_excl_sum(nmv, ws, &tmv);
  (void)cta_excl_sum(nst, ws, &tst);
  (void)cta_excl_sum(nhd, ws, &thd);
  (void)cta_excl_max(last, -1, wsi, &tmax);
  if (tid == 0) {
    TileAgg a; a.small = tmv; a.large = tst; a.last = tmax; a.lheads = thd;
    agg[((size_t)sel * gridDim.y + b) * g.tiles1 + tile] = a;
  }
}

// Pass B of a round (per list): refine the groups, write final order entries and new
// ranks, and compact the members that stay tied into next round's lists:
//   S' = [survivors of list S | members of list-L groups that shrank to <= SMALL_GROUP]
//   L' = the rest of list L, at list offset |S'|.
template <int PAIR, int CLS>
__global__ void __launch_bounds__(256, 4)
k_round_apply2(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B,
               const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sval,
               const uint32_t *__restrict__ pos, const uint32_t *__restrict__ ogs, uint32_t *__restrict__ nval,
               uint32_t *__restrict__ npos, uint32_t *__restrict__ ngs,
               const TileAgg *__restrict__ agg, uint32_t h, uint32_t sel) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  uint32_t lb, U;
  list_sel(meta[b], sel, lb, U);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t sa_off = lbz_slot_off(g, b);       // order / rank arrays, and the new lists
  const uint32_t off = sa_off + lb;                  // this (old) list
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const bool more = (2u * h < n);
  __shared__ uint32_t sb[132], mvb[128], stb[128], mpre[128], spre[128], hpre[128];
  __shared__ int lastpre[128];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  // totals of both lists (list sizes of the next round) and this tile's carries
  const uint32_t us = meta[b].us, ul = meta[b].ul;
  const uint32_t tiles_s = (us + LBZ_TILE - 1) / LBZ_TILE, tiles_l = (ul + LBZ_TILE - 1) / LBZ_TILE;
  const TileAgg *agg_s = agg + (size_t)b * g.tiles1;
  const TileAgg *agg_l = agg + ((size_t)gridDim.y + b) * g.tiles1;
  uint32_t s_tot = 0, l_mv = 0, l_st = 0, l_hd = 0, c_mv = 0, c_st = 0, c_hd = 0;
  int l = -1;
  for (uint32_t t = tid; t < tiles_s; t += 256) {
    const TileAgg a = agg_s[t];
    s_tot += a.small;
    if (sel == 0 && t < tile) { c_mv += a.small; l = max(l, a.last); }
  }
  for (uint32_t t = tid; t < tiles_l; t += 256) {
    const TileAgg a = agg_l[t];
    l_mv += a.small; l_st += a.large; l_hd += a.lheads;
    if (sel == 1 && t < tile) { c_mv += a.small; c_st += a.large; c_hd += a.lheads; l = max(l, a.last); }
  }
  uint32_t S_total, L_movers, L_stay, L_heads, carry_mv, carry_st, carry_hd;
  int carry_last;
  (void)cta_excl_sum(s_tot, ws, &S_total);
  (void)cta_excl_sum(l_mv, ws, &L_movers);
  (void)cta_excl_sum(l_st, ws, &L_stay);
  (void)cta_excl_sum(l_hd, ws, &L_heads);
  (void)cta_excl_sum(c_mv, ws, &carry_mv);
  (void)cta_excl_sum(c_st, ws, &carry_st);
  (void)cta_excl_sum(c_hd, ws, &carry_hd);
  (void)cta_excl_max(l, -1, wsi, &carry_last);
  const uint32_t us2 = S_total + L_movers;           // |S'| = list offset of L'
  const uint32_t mv_base = (sel ? S_total : 0u) + carry_mv;
  const uint32_t st_base = us2 + carry_st;

  round_head_bits<PAIR>(skey, off, tbase, U, sb);
  round_class_bits<CLS>(sb, tbase, U, more, mvb, stb);
  __syncthreads();
  uint32_t nmv = 0, nst = 0, nhd = 0;
  int lastw = -1;
  if (tid < 128) {
    const uint32_t H = sb[1 + tid];
    const uint32_t V = valid_bits(tbase + tid * 32, U);
    nmv = __popc(mvb[tid]); nst = __popc(stb[tid]); nhd = __popc(stb[tid] & H);
    if (H & V) lastw = (int)(tbase + tid * 32 + 31u - (uint32_t)__clz(H & V));
  }
  uint32_t tmv, tst, thd;
  int tmax;
  const uint32_t exmv = cta_excl_sum(nmv, ws, &tmv);
  const uint32_t exst = cta_excl_sum(nst, ws, &tst);
  const uint32_t exhd = cta_excl_sum(nhd, ws, &thd);
  const int exm = cta_excl_max(lastw, -1, wsi, &tmax);
  if (tid < 128) { mpre[tid] = exmv; spre[tid] = exst; hpre[tid] = exhd; lastpre[tid] = max(exm, carry_last); }
  __syncthreads();
  const uint32_t lt = lanemask_lt(), le = lanemask_le();
  const uint64_t pol = l2_policy_evict_last();
#pragma unroll 4
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {
    const uint32_t w = it * 8 + warp;
    const uint32_t j = tbase + w * 32 + lane;
    if (j < U) {
      const uint32_t myp = pos[off + j];
      const uint32_t myv = PAIR ? reinterpret_cast<const uint2 *>(skey)[off + j].y : sval[off + j];
      // sorting permutes a group only inside its own list range, so the unsorted gs
      // entry at this list index still names this element's old group
      const uint32_t oldgs = ogs[off + j] & 0xFFFFFu;
      const uint32_t H = sb[1 + w];
      const uint32_t m = H & le;
      const uint32_t st = m ? (tbase + w * 32 + 31u - (uint32_t)__clz(m)) : (uint32_t)lastpre[w];
      const uint32_t mygs = pos[off + st];                 // SA position of the group's head element
      const uint32_t mm = mvb[w], sm = stb[w];
      const bool mv = (mm >> lane) & 1u, stay = (sm >> lane) & 1u;
      // the order entry is only final once the rotation leaves the tied lists; the rank
      // only changes for rotations that are not in the first subgroup of their old group
      if (!mv && !stay) B.sa[sa_off + myp] = myv;
      if (mygs != oldgs) {
        if (B.hints) st_u32_hint(&B.rank[sa_off + myv], mygs, pol); else B.rank[sa_off + myv] = mygs;
      }
      if (mv) {
        const uint32_t o = sa_off + mv_base + mpre[w] + __popc(mm & lt);
        npos[o] = myp; nval[o] = myv; ngs[o] = mygs;
      } else if (stay) {
        const uint32_t o = sa_off + st_base + spre[w] + __popc(sm & lt);
        // ordinal of the surviving large group inside L' (its heads at or before j, minus one)
        const uint32_t ord = min(carry_hd + hpre[w] + __popc(sm & H & le) - 1u, ORD_LIMIT - 1u);
        npos[o] = myp; nval[o] = myv; ngs[o] = mygs | (ord << 20);
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= U) {                // last tile of this list
    // every last tile writes the same next-round geometry (committed by k_round_commit)
    meta[b].us_next = us2;
    meta[b].ul_next = L_stay;
    meta[b].lbase_next = us2;
    meta[b].depth = 2u * h;
    atomicMax(&B.counters[0], max(us2, L_stay));
    atomicAdd(&B.counters[1], sel ? (L_movers + L_stay) : S_total);
    atomicAdd(&B.counters[4], sel ? L_movers : S_total);  // diagnostics only (LBZ_ROUND_STATS)
    if (sel) atomicMax(&B.counters[5], L_heads);          // groups in L': decides the 32-bit key path
  }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bwt_final(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B) {
  const uint32_t b = blockIdx.y;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = blockIdx.x * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  for (uint32_t p = tbase + threadIdx.x; p < min(tbase + LBZ_TILE, n); p += 256) {
    const uint32_t v = B.sa[off + p];
    B.bwt[off + p] = B.T[off + (v ? v - 1 : n - 1)];
  }
}

// Primary index = first position of rotation 0's group; the group is larger than
// one only for exactly periodic blocks.  Its members are contiguous in the order,
// so the size is found by walking forward from the group start (one CTA per block).
__global__ void __launch_bounds__(256)
k_primary_index(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B) {
  const uint32_t b = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t r0 = B.rank[off];
  uint32_t total = 0;
  for (uint32_t base = r0; base < n; base += 256) {
    const uint32_t p = base + threadIdx.x;
    const bool same = (p < n) && (B.rank[off + B.sa[off + p]] == r0);
    const uint32_t c = (uint32_t)__syncthreads_count(same);
    total += c;
    if (c < 256) break;
  }
  if (threadIdx.x == 0) { meta[b].bwt_idx = r0; meta[b].tie_count = total; }
}

__global__ void k_bwt_prep(LbzBlockMeta *__restrict__ meta, uint32_t nblocks, uint32_t *__restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblocks) {
    meta[i].tie_count = 0; meta[i].unsorted = 0;
    meta[i].us = 0; meta[i].ul = 0; meta[i].lbase = 0; meta[i].us_next = 0; meta[i].ul_next = 0; meta[i].lbase_next = 0;
  }
  if (i == 0) { counters[0] = 0; counters[1] = 0; counters[3] = 0; counters[4] = 0; counters[5] = 0; }
}
// Between rounds: the tied-set size written by the last tile of a block becomes
// the segment size of the next round (kept apart so that no kernel reads and
// writes meta.unsorted at the same time).
__global__ void k_round_commit(LbzBlockMeta *__restrict__ meta, uint32_t nblocks, uint32_t *__restrict__ counters,
                               uint32_t *__restrict__ khist) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblocks) {
    meta[i].us = meta[i].us_next; meta[i].ul = meta[i].ul_next; meta[i].lbase = meta[i].lbase_next;
    meta[i].unsorted = meta[i].us + meta[i].ul;
  }
  if (i < nblocks * 5 * 256) khist[i] = 0;
  if (i == 0) { counters[0] = 0; counters[1] = 0; counters[4] = 0; counters[5] = 0; }
}

template <typename K, int COUNT_U, int GATHER, int THREADS, int ITEMS>
static int launch_radix_t(dim3 grid, cudaStream_t st, const LbzGeom &g, const LbzBlockMeta *meta, const uint8_t *T,
                          const uint32_t *sv, uint32_t *dv, const K *sk, K *dk, uint32_t *tstat, const uint32_t *gbase,
                          uint32_t gstride, uint32_t shift, uint32_t epoch, uint32_t *err) {
  const size_t smem = sizeof(RadixSmem<K, THREADS, ITEMS>);
  LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_radix_pass<K, COUNT_U, GATHER, THREADS, ITEMS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_radix_pass<K, COUNT_U, GATHER, THREADS, ITEMS><<<grid, THREADS, smem, st>>>(g, meta, T, sv, dv, sk, dk, tstat, gbase,
                                                                                 gstride, shift, epoch, err);
  return 0;
}
// cfg 0: 256 threads x 8 items (tile 2048); cfg 1: 512 threads x 8 items (tile 4096)
template <typename K, int COUNT_U>
static int launch_radix(int cfg, bool gather, uint32_t max_count, uint32_t nb, cudaStream_t st, const LbzGeom &g,
                        const LbzBlockMeta *meta, const uint8_t *T, const uint32_t *sv, uint32_t *dv, const K *sk, K *dk,
                        uint32_t *tstat, const uint32_t *gbase, uint32_t gstride, uint32_t shift, uint32_t epoch,
                        uint32_t *err) {
  const uint32_t rtile = cfg ? 4096u : 2048u;
  const dim3 grid((max_count + rtile - 1) / rtile, nb);
  if (cfg) {
    if (gather) return launch_radix_t<K, COUNT_U, 1, 512, 8>(grid, st, g, meta, T, sv, dv, sk, dk, tstat, gbase, gstride, shift, epoch, err);
    return launch_radix_t<K, COUNT_U, 0, 512, 8>(grid, st, g, meta, T, sv, dv, sk, dk, tstat, gbase, gstride, shift, epoch, err);
  }
  if (gather) return launch_radix_t<K, COUNT_U, 1, 256, 8>(grid, st, g, meta, T, sv, dv, sk, dk, tstat, gbase, gstride, shift, epoch, err);
  return launch_radix_t<K, COUNT_U, 0, 256, 8>(grid, st, g, meta, T, sv, dv, sk, dk, tstat, gbase, gstride, shift, epoch, err);
}

// Status-word epoch: 1..1023, the status array is cleared when it wraps.
static uint32_t next_epoch(const BwtBuffers &B, uint32_t nb, const LbzGeom &g, cudaStream_t st) {
  uint32_t e = *B.epoch + 1;
  if (e >= 1024) {
    cudaMemsetAsync(B.tstat, 0, (size_t)nb * (g.S1 / 2048u) * 256 * sizeof(uint32_t), st);
    cudaMemsetAsync(B.tickets, 0, 1024 * sizeof(uint32_t), st);
    e = 1;
  }
  *B.epoch = e;
  return e;
}

// ---------------------------------------------------------------------------
// Host driver.  `h_counters` is pinned host memory for the per-round readback.
extern "C" int lbz_run_bwt(const LbzGeom *gp, LbzBlockMeta *d_meta, BwtBuffers B,
                           uint32_t *h_counters, uint32_t *rounds_out, uint64_t *launches,
                           const LbzTimers *tm, cudaStream_t st) {
  uint64_t nl = 0;
  const LbzGeom g = *gp;
  const uint32_t nb = 2 * g.nchunks;
  if (nb == 0) return 0;
  const dim3 grid_full(g.tiles1, nb);

  k_bwt_prep<<<(nb + 255) / 256, 256, 0, st>>>(d_meta, nb, B.counters);
  uint2 *pa = reinterpret_cast<uint2 *>(B.key), *pb = reinterpret_cast<uint2 *>(B.key2);
  const uint32_t K = B.K;
  nl += 2 + K + 2;
  static int cfg = -1;
  if (cfg < 0) { const char *ev = getenv("LBZ_RADIX_CFG"); cfg = ev ? (atoi(ev) != 0) : 1; }

  k_text_bases<<<nb, 1024, 0, st>>>(g, d_meta, B.T, B.gbase);
  if (tp_version() == 3) { k_worklist<0><<<1, 1024, 0, st>>>(d_meta, nb, g.tiles1, B.wl, B.wl_count); nl++; }
  // LSD over text bytes 7..0 of every rotation: bytes 4..7 travel as a 32-bit key
  // next to the index for the first four passes, bytes 0..3 are fetched once by
  // the fifth pass; the first pass builds its pairs from the text, the last one
  // writes the order only.
  // K <= 4 would need a single-phase variant; the engine only offers 5..8.
  const uint32_t koff = K - 4u;
  for (uint32_t p = 0; p < K; p++) {
    if (tm && tm->enabled && p < LBZ_NK0) cudaEventRecord(tm->k0[2 * p], st);
    const uint32_t ep = next_epoch(B, nb, g, st);
    const uint32_t shift = p < 4u ? 8u * p : 8u * (p + 4u - K);
    const bool last = (p == K - 1u);
    int rc;
    uint32_t *const er = B.counters + 3;
    if (p == 0) rc = launch_text_pass<2, 0>(nb, st, g, d_meta, B.T, pa, pb, B.sa, B.tstat, B.gbase, shift, ep, er, koff, B);
    else if (p == 4 && last) rc = launch_text_pass<1, 1>(nb, st, g, d_meta, B.T, pa, pb, B.sa, B.tstat, B.gbase, shift, ep, er, koff, B);
    else if (p == 4) rc = launch_text_pass<1, 0>(nb, st, g, d_meta, B.T, pa, pb, B.sa, B.tstat, B.gbase, shift, ep, er, koff, B);
    else if (last) rc = launch_text_pass<0, 1>(nb, st, g, d_meta, B.T, pa, pb, B.sa, B.tstat, B.gbase, shift, ep, er, koff, B);
    else rc = launch_text_pass<0, 0>(nb, st, g, d_meta, B.T, pa, pb, B.sa, B.tstat, B.gbase, shift, ep, er, koff, B);
    if (rc) return -1;
    if (tm && tm->enabled && p < LBZ_NK0) cudaEventRecord(tm->k0[2 * p + 1], st);
    { uint2 *t = pa; pa = pb; pb = t; }
  }
  TileAgg *agg = reinterpret_cast<TileAgg *>(B.agg);
  static int rv1 = -1;
  if (rv1 < 0) { const char *ev = getenv("LBZ_ROUND_V1"); rv1 = (ev && atoi(ev) != 0) ? 1 : 0; }
  if (rv1) {
    k_heads_agg<<<grid_full, 256, 0, st>>>(g, d_meta, B.T, B.sa, B.head, agg, K);
    k_ranks_compact<<<grid_full, 256, 0, st>>>(g, d_meta, B, agg);
  } else {
    k_heads_bits<<<grid_full, 256, 0, st>>>(g, d_meta, B.T, B.sa, B.hbits, B.cbits, agg, K);
    k_ranks_compact2<<<grid_full, 256, 0, st>>>(g, d_meta, B, agg);
  }
  LBZ_CUDA_CHECK(cudaGetLastError());
  if (tm && tm->enabled) cudaEventRecord(tm->stage[2], st);
  if (B.on_sorted) B.on_sorted(B.on_sorted_arg);

  uint32_t rounds = 0;
  uint32_t h = K;
  // Per round: keys are built into kA from the lists (vA, gA); list S is sorted by
  // k_small_sort into (kB, vB); list L either by four passes over 32-bit-key pairs
  // (kA -> kB -> kA -> kB -> kA, while every block has fewer than ORD_LIMIT groups in L)
  // or by five passes over 40-bit keys ((kA,vA) -> ... -> (kB,vB)); the survivors are
  // compacted into (vA, pB, gB), which become the next round's lists.
  uint64_t *kA = B.key, *kB = B.key2;
  uint32_t *vA = B.val, *vB = B.val2;
  uint32_t *pA = B.pos, *pB = B.pos2;
  uint32_t *gA = B.gs, *gB = B.gs2;
  static int l64 = -1;
  if (l64 < 0) { const char *ev = getenv("LBZ_L64"); l64 = (ev && atoi(ev) != 0) ? 1 : 0; }
  for (;;) {
    LBZ_CUDA_CHECK(cudaMemcpyAsync(h_counters, B.counters, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    LBZ_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_counters[3]) {
      fprintf(stderr, "lbzip2_b200: sort pass failed (%s)\n", h_counters[3] == 2u ? "bulk copy never landed" : "chained scan timed out");
      return -1;
    }
    const uint32_t maxU = h_counters[0];
    const bool pair = !rv1 && !l64 && h_counters[5] <= ORD_LIMIT;
    static int round_stats = -1;
    if (round_stats < 0) round_stats = getenv("LBZ_ROUND_STATS") != nullptr;
    if (round_stats)
      fprintf(stderr, "lbzip2_b200: sort depth %u: %u rotations still tied, %u of them in groups <= %u (largest list %u, "
              "at most %u groups in a list L: %s keys)\n", h, h_counters[1], h_counters[4], SMALL_GROUP, maxU, h_counters[5],
              pair ? "32-bit" : "40-bit");
    if (maxU == 0) break;
    rounds++;
    nl += 5 + (pair ? 4 : 5) + 4;
    const dim3 grid_u((maxU + LBZ_TILE - 1) / LBZ_TILE, nb);
    const uint32_t *gb1 = B.gbase + (size_t)nb * 256;
    k_round_commit<<<(nb * 5 * 256 + 255) / 256, 256, 0, st>>>(d_meta, nb, B.counters, B.khist);
    k_round_keys<<<grid_u, 256, 0, st>>>(g, d_meta, vA, gA, B.rank, kA, B.khist, h, 0u, B.hints, 0);
    k_round_keys<<<grid_u, 256, 0, st>>>(g, d_meta, vA, gA, B.rank, kA, B.khist, h, 1u, B.hints, pair ? 1 : 0);
    k_key_bases<<<nb, 256, 0, st>>>(d_meta, B.khist, B.gbase + (size_t)nb * 256);
    k_small_sort<<<grid_u, 256, 0, st>>>(g, d_meta, kA, vA, kB, vB);
    if (pair && tp_version() == 3) {
      k_worklist<1><<<1, 1024, 0, st>>>(d_meta, nb, min(grid_u.x, g.tiles1), B.wl + B.wl_list_off, B.wl_count + 1);
      nl++;
    }
    if (pair) {
      uint2 *pa2 = reinterpret_cast<uint2 *>(kA), *pb2 = reinterpret_cast<uint2 *>(kB);
      for (uint32_t p = 0; p < 4; p++) {
        if (launch_list_pass(maxU, nb, st, g, d_meta, pa2, pb2, B.tstat, gb1 + p * 256, 5u * 256u, 8u * p,
                             next_epoch(B, nb, g, st), B.counters + 3, B)) return -1;
        uint2 *t = pa2; pa2 = pb2; pb2 = t;
      }
    } else {
      uint64_t *ks = kA, *kd = kB;
      uint32_t *vs = vA, *vd = vB;
      for (uint32_t p = 0; p < 5; p++) {
        if (launch_radix<uint64_t, 1>(cfg, false, maxU, nb, st, g, d_meta, nullptr, vs, vd, ks, kd, B.tstat, gb1 + p * 256,
                                      5u * 256u, 8u * p, next_epoch(B, nb, g, st), B.counters + 3)) return -1;
        uint64_t *tk = ks; ks = kd; kd = tk;
        uint32_t *tv = vs; vs = vd; vd = tv;
      }
    }
    if (rv1) {
      for (uint32_t sel = 0; sel < 2; sel++) {
        k_round_agg<<<grid_u, 256, 0, st>>>(g, d_meta, kB, agg, h, sel);
        k_round_apply<<<grid_u, 256, 0, st>>>(g, d_meta, B, kB, vB, pA, vA, pB, gB, agg, h, sel);
      }
    } else {
      k_round_agg2<0, 0><<<grid_u, 256, 0, st>>>(g, d_meta, kB, agg, h, 0u);
      if (pair) k_round_agg2<1, 1><<<grid_u, 256, 0, st>>>(g, d_meta, kA, agg, h, 1u);
      else k_round_agg2<0, 1><<<grid_u, 256, 0, st>>>(g, d_meta, kB, agg, h, 1u);
      k_round_apply2<0, 0><<<grid_u, 256, 0, st>>>(g, d_meta, B, kB, vB, pA, gA, vA, pB, gB, agg, h, 0u);
      if (pair) k_round_apply2<1, 1><<<grid_u, 256, 0, st>>>(g, d_meta, B, kA, nullptr, pA, gA, vA, pB, gB, agg, h, 1u);
      else k_round_apply2<0, 1><<<grid_u, 256, 0, st>>>(g, d_meta, B, kB, vB, pA, gA, vA, pB, gB, agg, h, 1u);
    }
    { uint32_t *t = pA; pA = pB; pB = t; }
    { uint32_t *t = gA; gA = gB; gB = t; }
    h *= 2;
    LBZ_CUDA_CHECK(cudaGetLastError());
  }
  if (tm && tm->enabled) cudaEventRecord(tm->stage[3], st);
  k_bwt_final<<<grid_full, 256, 0, st>>>(g, d_meta, B);
  k_primary_index<<<nb, 256, 0, st>>>(g, d_meta, B);
  LBZ_CUDA_CHECK(cudaGetLastError());
  nl += 2;
  if (rounds_out) *rounds_out = rounds;
  if (launches) *launches += nl;
  return 0;
}
