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
  uint32_t *hist;         // [(b*tiles1 + tile)*256 + digit]
  uint32_t *digit_base;   // [b*256 + digit]
  uint32_t *counters;     // [0] = max unsorted over blocks, [1] = total unsorted
  uint8_t *bwt;           // output last column
};

__device__ __forceinline__ uint32_t wrap_add(uint32_t v, uint32_t d, uint32_t n) {
  uint32_t j = v + d;
  if (j >= n) { j -= n; if (j >= n) j %= n; }
  return j;
}

// ---------------------------------------------------------------------------
__global__ void k_sa_init(LbzGeom g, const LbzBlockMeta *__restrict__ meta, uint32_t *__restrict__ sa) {
  const uint32_t b = blockIdx.y;
  const uint32_t n = meta[b].n;
  const uint32_t base = blockIdx.x * LBZ_TILE;
  if (base >= n) return;
  uint32_t *p = sa + lbz_slot_off(g, b);
  for (uint32_t i = base + threadIdx.x; i < min(base + LBZ_TILE, n); i += blockDim.x) p[i] = i;
}

// ---------------------------------------------------------------------------
// Counting sort pass, segmented by block.  MODE 0: digit = text byte
// T[(v + d) mod n], payload = rotation index only.  MODE 1: digit = byte
// `shift/8` of a 64-bit key that travels with the payload.
template <int MODE>
__device__ __forceinline__ uint32_t seg_count(const LbzBlockMeta &m) { return MODE == 0 ? m.n : m.unsorted; }

template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS)
k_hist(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
       const uint32_t *__restrict__ src_val, const uint64_t *__restrict__ src_key,
       uint32_t *__restrict__ hist, uint32_t dsh) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t cnt = seg_count<MODE>(meta[b]);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= cnt) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint32_t dg[SORT_ITEMS];
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++) {
    const uint32_t idx = tbase + warp * (32 * SORT_ITEMS) + it * 32 + lane;
    dg[it] = 0xFFFFFFFFu;
    if (idx < cnt) dg[it] = (MODE == 0) ? src_val[off + idx] : ((uint32_t)(src_key[off + idx] >> dsh) & 0xFFu);
  }
  if (MODE == 0) {
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++)
      if (dg[it] != 0xFFFFFFFFu) dg[it] = T[off + wrap_add(dg[it], dsh, n)];
  }
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++)
    if (dg[it] != 0xFFFFFFFFu) atomicAdd(&sh[dg[it]], 1u);
  __syncthreads();
  hist[((size_t)b * g.tiles1 + tile) * 256 + threadIdx.x] = sh[threadIdx.x];
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_scan(LbzGeom g, const LbzBlockMeta *__restrict__ meta, uint32_t *__restrict__ hist,
       uint32_t *__restrict__ digit_base) {
  const uint32_t b = blockIdx.x;
  const uint32_t cnt = seg_count<MODE>(meta[b]);
  if (cnt == 0) return;
  const uint32_t ntiles = (cnt + LBZ_TILE - 1) / LBZ_TILE;
  uint32_t *h = hist + (size_t)b * g.tiles1 * 256 + threadIdx.x;
  uint32_t run = 0;
  uint32_t t = 0;
  for (; t + 4 <= ntiles; t += 4) {
    const uint32_t a0 = h[(t + 0) * 256], a1 = h[(t + 1) * 256], a2 = h[(t + 2) * 256], a3 = h[(t + 3) * 256];
    h[(t + 0) * 256] = run; run += a0;
    h[(t + 1) * 256] = run; run += a1;
    h[(t + 2) * 256] = run; run += a2;
    h[(t + 3) * 256] = run; run += a3;
  }
  for (; t < ntiles; t++) { const uint32_t a = h[t * 256]; h[t * 256] = run; run += a; }
  __shared__ uint32_t ws[40];
  uint32_t total;
  const uint32_t ex = cta_excl_sum(run, ws, &total);
  digit_base[b * 256 + threadIdx.x] = ex;
}

template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS)
k_scatter(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
          const uint32_t *__restrict__ src_val, uint32_t *__restrict__ dst_val,
          const uint64_t *__restrict__ src_key, uint64_t *__restrict__ dst_key,
          const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_base, uint32_t dsh) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t cnt = seg_count<MODE>(meta[b]);
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= cnt) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
  __shared__ uint32_t wcnt[SORT_WARPS][256];
  for (uint32_t i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t lt = lanemask_lt();
  uint32_t val[SORT_ITEMS];
  uint32_t rd[SORT_ITEMS];     // rank within warp strip | digit << 16 | valid << 31
  uint64_t key[MODE == 1 ? SORT_ITEMS : 1];

  // Phase 1: all independent loads first (16 in flight per thread), then all
  // dependent text gathers, and only then the shared-memory ranking, so that
  // the two global latencies are paid once per tile instead of once per item.
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++) {
    const uint32_t idx = tbase + warp * (32 * SORT_ITEMS) + it * 32 + lane;
    const bool valid = idx < cnt;
    val[it] = valid ? src_val[off + idx] : 0u;
    if (MODE == 1) key[it] = valid ? src_key[off + idx] : 0ull;
    rd[it] = valid ? 0x80000000u : 0u;
  }
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++) {
    uint32_t digit = 0x100u;                   // invalid lanes form their own match group
    if (rd[it]) {
      if (MODE == 0) digit = T[off + wrap_add(val[it], dsh, n)];
      else digit = (uint32_t)(key[it] >> dsh) & 0xFFu;
    }
    rd[it] |= digit << 16;
  }
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++) {
    const bool valid = (rd[it] & 0x80000000u) != 0;
    const uint32_t digit = (rd[it] >> 16) & 0x1FFu;
    const uint32_t mask = __match_any_sync(0xffffffffu, digit);
    uint32_t base = 0;
    if (valid) base = wcnt[warp][digit];
    __syncwarp();
    if (valid && (mask & lt) == 0) wcnt[warp][digit] = base + __popc(mask);   // group leader
    __syncwarp();
    rd[it] |= base + __popc(mask & lt);
  }
  __syncthreads();
  {
    const uint32_t d = threadIdx.x;
    uint32_t run = hist[((size_t)b * g.tiles1 + tile) * 256 + d] + digit_base[b * 256 + d];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
      const uint32_t c = wcnt[w][d];
      wcnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; it++) {
    if (rd[it] & 0x80000000u) {
      const uint32_t digit = (rd[it] >> 16) & 0xFFu;
      const uint32_t dst = wcnt[warp][digit] + (rd[it] & 0xFFFFu);
      dst_val[off + dst] = val[it];
      if (MODE == 1) dst_key[off + dst] = key[it];
    }
  }
}

// ---------------------------------------------------------------------------
// Group heads after the initial sort: head[p] = first BWT_K bytes of rotation
// sa[p] differ from those of sa[p-1].  Keys are only compared for equality, so
// the 8 bytes are fetched as three aligned words and funnel-shifted.
__device__ __forceinline__ uint64_t text_key(const uint8_t *__restrict__ Tb, uint32_t v, uint32_t n) {
  if (v + BWT_K <= n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(Tb + (v & ~3u));
    const uint32_t sh = 8u * (v & 3u);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];      // slot capacity >= n + 64: in bounds
    const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
    return ((uint64_t)x1 << 32) | x0;
  }
  uint64_t k = 0;
  uint32_t j = v;
#pragma unroll
  for (uint32_t d = 0; d < BWT_K; d++) { k |= (uint64_t)Tb[j] << (8 * d); if (++j >= n) j = 0; }
  return k;
}

struct TileAgg { uint32_t count; int last; };

// Tile-parallel pass A: head flags of a tile, number of rotations still tied,
// last head position.  Per-tile aggregates replace a serial scan: pass B sums
// the (<= 220) aggregates of the preceding tiles of its block itself.
__global__ void __launch_bounds__(256)
k_heads_agg(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ T,
            const uint32_t *__restrict__ sa, uint8_t *__restrict__ head, TileAgg *__restrict__ agg) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *Tb = T + off;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  __shared__ __align__(16) uint8_t sflag[LBZ_TILE + 32];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
#pragma unroll 2
  for (uint32_t it = 0; it < LBZ_TILE / 256; it++) {     // uniform trip count: shuffles stay converged
    const uint32_t p = tbase + it * 256 + tid;
    const bool valid = p < n;
    uint64_t k = 0;
    if (valid) k = text_key(Tb, sa[off + p], n);
    uint64_t kprev = __shfl_up_sync(0xffffffffu, k, 1);
    if (valid) {
      if (lane == 0) kprev = p ? text_key(Tb, sa[off + p - 1], n) : ~k;
      sflag[p - tbase] = (p == 0) || (k != kprev);
    } else if (p - tbase <= LBZ_TILE) {
      sflag[p - tbase] = 1;                              // end of block acts as a head
    }
  }
  if (tid == 0) {
    const uint32_t pn = tbase + LBZ_TILE;
    sflag[LBZ_TILE] = (pn < n) ? (text_key(Tb, sa[off + pn], n) != text_key(Tb, sa[off + pn - 1], n)) : 1;
  }
  __syncthreads();
  const bool tracking = BWT_K < n;
  const uint32_t q0 = tid * 16;
  const uint4 fv = *reinterpret_cast<const uint4 *>(&sflag[q0]);
  *reinterpret_cast<uint4 *>(&head[off + tbase + q0]) = fv;
  uint32_t uns = 0;
  int last = -1;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t p = tbase + q0 + j;
    if (p < n) {
      const bool f = sflag[q0 + j], f1 = sflag[q0 + j + 1];
      if (f) last = (int)p;
      uns += !(f && f1);
    }
  }
  uint32_t tot;
  int tmax;
  (void)cta_excl_sum(uns, ws, &tot);
  (void)cta_excl_max(last, -1, wsi, &tmax);
  if (tid == 0) { TileAgg a; a.count = tracking ? tot : 0u; a.last = tmax; agg[(size_t)b * g.tiles1 + tile] = a; }
}

// Tile-parallel pass B: ranks (rank[i] = first position of i's group) and
// compaction of the members of tied groups into the round lists.
__global__ void __launch_bounds__(256)
k_ranks_compact(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B, const TileAgg *__restrict__ agg) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t n = meta[b].n;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= n) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t tid = threadIdx.x;
  const bool tracking = BWT_K < n;
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  // carry from the preceding tiles of this block
  uint32_t c = 0;
  int l = -1;
  for (uint32_t t = tid; t < tile; t += 256) {
    const TileAgg a = agg[(size_t)b * g.tiles1 + t];
    c += a.count; l = max(l, a.last);
  }
  uint32_t carry_cnt;
  int carry_last;
  (void)cta_excl_sum(c, ws, &carry_cnt);
  (void)cta_excl_max(l, -1, wsi, &carry_last);

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
  for (int j = 0; j < 16; j++) if (p0 + j < n && f[j]) last = (int)(p0 + j);
  int tmax;
  int st = cta_excl_max(last, -1, wsi, &tmax);
  st = max(st, carry_last);
  uint32_t unsmask = 0;
  uint32_t starts[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (p0 + j < n) {
      if (f[j]) st = (int)(p0 + j);
      starts[j] = (uint32_t)st;
      if (tracking && !(f[j] && f[j + 1])) unsmask |= 1u << j;
    }
  }
  uint32_t tot;
  uint32_t o = carry_cnt + cta_excl_sum(__popc(unsmask), ws, &tot);
  if (p0 < n) {
    uint32_t v[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 x = *reinterpret_cast<const uint4 *>(&B.sa[off + p0 + 4 * q]);   // slot capacity covers the overread
      v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (p0 + j < n) {
        B.rank[off + v[j]] = starts[j];
        if (unsmask & (1u << j)) {
          B.pos[off + o] = p0 + j;
          B.val[off + o] = v[j];
          B.gs[off + o] = starts[j];
          o++;
        }
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= n) {               // last tile of the block
    const uint32_t U = carry_cnt + tot;
    meta[b].pad_[1] = U;                                  // committed by k_round_commit
    meta[b].depth = BWT_K;
    atomicMax(&B.counters[0], U);
    atomicAdd(&B.counters[1], U);
  }
}

// ---------------------------------------------------------------------------
// Round key: (group start << 20) | rank of rotation (i + h).
__global__ void __launch_bounds__(256)
k_round_keys(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint32_t *__restrict__ val,
             const uint32_t *__restrict__ gs, const uint32_t *__restrict__ rank,
             uint64_t *__restrict__ key, uint32_t h) {
  const uint32_t b = blockIdx.y;
  const uint32_t U = meta[b].unsorted;
  const uint32_t tbase = blockIdx.x * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
  for (uint32_t j = tbase + threadIdx.x; j < min(tbase + LBZ_TILE, U); j += 256) {
    const uint32_t v = val[off + j];
    const uint32_t r = rank[off + wrap_add(v, h, n)];
    key[off + j] = ((uint64_t)gs[off + j] << 20) | r;
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
            TileAgg *__restrict__ agg, uint32_t h) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t U = meta[b].unsorted;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
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
  if (tid == 0) { TileAgg a; a.count = tot; a.last = tmax; agg[(size_t)b * g.tiles1 + tile] = a; }
}

__global__ void __launch_bounds__(256)
k_round_apply(LbzGeom g, LbzBlockMeta *__restrict__ meta, BwtBuffers B,
              const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sval,
              const uint32_t *__restrict__ pos, uint32_t *__restrict__ nval,
              uint32_t *__restrict__ npos, uint32_t *__restrict__ ngs,
              const TileAgg *__restrict__ agg, uint32_t h) {
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t U = meta[b].unsorted;
  const uint32_t tbase = tile * LBZ_TILE;
  if (tbase >= U) return;
  const uint32_t n = meta[b].n;
  const uint32_t off = lbz_slot_off(g, b);
  const uint32_t tid = threadIdx.x;
  const bool more = (2u * h < n);
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  uint32_t c = 0;
  int l = -1;
  for (uint32_t t = tid; t < tile; t += 256) {
    const TileAgg a = agg[(size_t)b * g.tiles1 + t];
    c += a.count; l = max(l, a.last);
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
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const uint32_t j = j0 + q;
    if (j < U) {
      B.sa[off + myp[q]] = myv[q];
      B.rank[off + myv[q]] = mygs[q];
      if (unsmask & (1u << q)) {
        npos[off + o] = myp[q];
        nval[off + o] = myv[q];
        ngs[off + o] = mygs[q];
        o++;
      }
    }
  }
  if (tid == 0 && tbase + LBZ_TILE >= U) {                // last tile of this block's list
    const uint32_t U2 = carry_cnt + tot;
    meta[b].pad_[1] = U2;                                  // committed by k_round_commit
    meta[b].depth = 2u * h;
    atomicMax(&B.counters[0], U2);
    atomicAdd(&B.counters[1], U2);
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
  const uint32_t r0 = B.rank[off];
  uint32_t ties = 0;
  for (uint32_t p = tbase + threadIdx.x; p < min(tbase + LBZ_TILE, n); p += 256) {
    const uint32_t v = B.sa[off + p];
    B.bwt[off + p] = B.T[off + (v ? v - 1 : n - 1)];
    ties += (B.rank[off + v] == r0);
  }
  if (ties) atomicAdd(&meta[b].tie_count, ties);
  if (blockIdx.x == 0 && threadIdx.x == 0) meta[b].bwt_idx = r0;
}

__global__ void k_bwt_prep(LbzBlockMeta *__restrict__ meta, uint32_t nblocks, uint32_t *__restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblocks) { meta[i].tie_count = 0; meta[i].unsorted = 0; meta[i].pad_[1] = 0; }
  if (i == 0) { counters[0] = 0; counters[1] = 0; }
}
// Between rounds: the tied-set size written by the last tile of a block becomes
// the segment size of the next round (kept apart so that no kernel reads and
// writes meta.unsorted at the same time).
__global__ void k_round_commit(LbzBlockMeta *__restrict__ meta, uint32_t nblocks, uint32_t *__restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblocks) meta[i].unsorted = meta[i].pad_[1];
  if (i == 0) { counters[0] = 0; counters[1] = 0; }
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
  k_sa_init<<<grid_full, 256, 0, st>>>(g, d_meta, B.sa);
  nl += 2 + 3 * BWT_K + 2;

  uint32_t *src = B.sa, *dst = B.sa2;
  for (int d = (int)BWT_K - 1; d >= 0; d--) {
    k_hist<0><<<grid_full, SORT_THREADS, 0, st>>>(g, d_meta, B.T, src, nullptr, B.hist, (uint32_t)d);
    k_scan<0><<<nb, 256, 0, st>>>(g, d_meta, B.hist, B.digit_base);
    if (tm && tm->enabled) cudaEventRecord(tm->k0[2 * d], st);
    k_scatter<0><<<grid_full, SORT_THREADS, 0, st>>>(g, d_meta, B.T, src, dst, nullptr, nullptr,
                                                       B.hist, B.digit_base, (uint32_t)d);
    if (tm && tm->enabled) cudaEventRecord(tm->k0[2 * d + 1], st);
    uint32_t *t = src; src = dst; dst = t;
  }
  // BWT_K is even, so the order is back in B.sa
  B.sa = src; B.sa2 = dst;
  TileAgg *agg = reinterpret_cast<TileAgg *>(B.hist);     // hist is free between radix passes
  k_heads_agg<<<grid_full, 256, 0, st>>>(g, d_meta, B.T, B.sa, B.head, agg);
  k_ranks_compact<<<grid_full, 256, 0, st>>>(g, d_meta, B, agg);
  LBZ_CUDA_CHECK(cudaGetLastError());
  if (tm && tm->enabled) cudaEventRecord(tm->stage[2], st);

  uint32_t rounds = 0;
  uint32_t h = BWT_K;
  uint64_t *ksrc = B.key, *kdst = B.key2;
  uint32_t *vsrc = B.val, *vdst = B.val2;
  uint32_t *psrc = B.pos, *pdst = B.pos2;
  uint32_t *gsrc = B.gs, *gdst = B.gs2;
  for (;;) {
    LBZ_CUDA_CHECK(cudaMemcpyAsync(h_counters, B.counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    LBZ_CUDA_CHECK(cudaStreamSynchronize(st));
    const uint32_t maxU = h_counters[0];
    if (maxU == 0) break;
    rounds++;
    nl += 2 + 15 + 2;
    const dim3 grid_u((maxU + LBZ_TILE - 1) / LBZ_TILE, nb);
    k_round_commit<<<(nb + 255) / 256, 256, 0, st>>>(d_meta, nb, B.counters);
    k_round_keys<<<grid_u, 256, 0, st>>>(g, d_meta, vsrc, gsrc, B.rank, ksrc, h);
    for (uint32_t sh = 0; sh < 40; sh += 8) {
      k_hist<1><<<grid_u, SORT_THREADS, 0, st>>>(g, d_meta, nullptr, vsrc, ksrc, B.hist, sh);
      k_scan<1><<<nb, 256, 0, st>>>(g, d_meta, B.hist, B.digit_base);
      k_scatter<1><<<grid_u, SORT_THREADS, 0, st>>>(g, d_meta, nullptr, vsrc, vdst, ksrc, kdst,
                                                      B.hist, B.digit_base, sh);
      uint64_t *tk = ksrc; ksrc = kdst; kdst = tk;
      uint32_t *tv = vsrc; vsrc = vdst; vdst = tv;
    }
    // sorted (key,val) now in ksrc/vsrc; compacted survivors go to vdst/pdst/gdst
    k_round_agg<<<grid_u, 256, 0, st>>>(g, d_meta, ksrc, agg, h);
    k_round_apply<<<grid_u, 256, 0, st>>>(g, d_meta, B, ksrc, vsrc, psrc, vdst, pdst, gdst, agg, h);
    { uint32_t *t = vsrc; vsrc = vdst; vdst = t; }
    { uint32_t *t = psrc; psrc = pdst; pdst = t; }
    { uint32_t *t = gsrc; gsrc = gdst; gdst = t; }
    h *= 2;
    LBZ_CUDA_CHECK(cudaGetLastError());
  }
  if (tm && tm->enabled) cudaEventRecord(tm->stage[3], st);
  k_bwt_final<<<grid_full, 256, 0, st>>>(g, d_meta, B);
  LBZ_CUDA_CHECK(cudaGetLastError());
  nl += 1;
  if (rounds_out) *rounds_out = rounds;
  if (launches) *launches += nl;
  return 0;
}
