// rle1.cu -- initial run-length coding + block split + CRC-32 + used-byte map.
//
// Replaces collect() (reference src/encode.c:135-336) together with the run
// flush at the top of encode() (src/encode.c:443-447) and the chunk -> block
// split performed by the scheduler (src/compress.c:93-110).
//
// One CTA per raw chunk.  The serial state machine of the reference is
// restated position-wise: a maximal run of equal bytes starting at s is cut
// into pieces of 259; input position i with k' = (i - s) mod 259 emits its
// literal iff k' < 4 and a count byte (k' - 3) iff k' == 258 or (k' >= 3 and
// the run ends at i).  Output offsets are a prefix sum, the run start is a
// prefix max, so a tile of 8192 input bytes is processed by 1024 threads with
// two CTA scans and the state carried across tiles in registers.
// The block closes at the first i with cum(i) >= cap, or with k' == 2,
// cum(i) == cap-1 and the run continuing (src/encode.c:218: the 4th literal
// is only written with room for its count byte).  The rest of the chunk is
// re-segmented from the cut into the second block slot.
#include "lbz_common.cuh"

#define RLE_THREADS 1024
#define RLE_PER_THREAD 8
#define RLE_TILE (RLE_THREADS * RLE_PER_THREAD)
#define CRC_POLY 0x04C11DB7u

// (a * b) mod P over GF(2), bit 31 = x^31.
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll 8
  for (int i = 31; i >= 0; --i) {
    r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
    if ((a >> i) & 1u) r ^= b;
  }
  return r;
}

// x^(8*L) mod P
__device__ uint32_t gf_xpow8(uint32_t L) {
  uint32_t res = 1u, base = 0x100u;
  while (L) {
    if (L & 1u) res = gf_mulmod(res, base);
    base = gf_mulmod(base, base);
    L >>= 1;
  }
  return res;
}

__global__ void __launch_bounds__(RLE_THREADS, 1)
k_rle1(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len,
       uint8_t *__restrict__ T, LbzBlockMeta *__restrict__ meta) {
  const uint32_t c = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const uint32_t N = chunk_len[c];
  const uint32_t cap = g.mbs;
  const uint8_t *x = in + (size_t)c * g.mbs;
  uint8_t *Tp0 = T + lbz_slot_off(g, 2 * c);
  uint8_t *Tp1 = T + lbz_slot_off(g, 2 * c + 1);

  __shared__ uint32_t s_ws[40];
  __shared__ int s_wsi[40];
  __shared__ uint8_t s_first[32], s_last[32];
  __shared__ uint8_t s_used[2][256];
  __shared__ uint32_t s_cut, s_ncut;
  __shared__ uint32_t s_crctab[256];
  __shared__ uint32_t s_crcacc[2];

  for (uint32_t i = tid; i < 512; i += RLE_THREADS) (&s_used[0][0])[i] = 0;
  if (tid < 256) {
    uint32_t r = tid << 24;
#pragma unroll
    for (int k = 0; k < 8; k++) r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
    s_crctab[tid] = r;
  }
  if (tid == 0) { s_crcacc[0] = 0; s_crcacc[1] = 0; }

  // CTA-uniform carried state
  uint32_t part = 0, seg_start = 0, m = 0;
  int run_carry = 0;          // start of the run that contains the previous tile's last byte
  uint32_t prev_tile_byte = 0;
  uint32_t n0 = 0, raw0 = N;  // results for part 0 (raw0 = N when no cut happens)

  for (uint32_t tb = 0; tb < N; tb += RLE_TILE) {
    const uint32_t i0 = tb + tid * RLE_PER_THREAD;
    uint32_t b[RLE_PER_THREAD];
    if (i0 + RLE_PER_THREAD <= N) {
      const uint2 w = *reinterpret_cast<const uint2 *>(x + i0);
#pragma unroll
      for (int j = 0; j < 4; j++) { b[j] = (w.x >> (8 * j)) & 0xFFu; b[4 + j] = (w.y >> (8 * j)) & 0xFFu; }
    } else {
#pragma unroll
      for (int j = 0; j < RLE_PER_THREAD; j++) b[j] = (i0 + j < N) ? x[i0 + j] : 0x100u;  // 0x100 = "no byte"
    }
    // neighbours
    uint32_t prevb = __shfl_up_sync(0xffffffffu, b[7], 1);
    uint32_t nextb = __shfl_down_sync(0xffffffffu, b[0], 1);
    __syncthreads();   // s_first/s_last reuse across tiles + redo passes
    if (lane == 0) s_first[warp] = (uint8_t)b[0];
    if (lane == 31) s_last[warp] = (uint8_t)b[7];
    __syncthreads();
    if (lane == 0) prevb = warp ? s_last[warp - 1] : prev_tile_byte;
    if (lane == 31) {
      if (warp < 31) nextb = s_first[warp + 1];
      else nextb = (tb + RLE_TILE < N) ? x[tb + RLE_TILE] : 0x100u;
    }
    // Positions past N inside a warp: the shuffled neighbour of a valid last
    // position may be a "no byte" marker only when i+1 >= N, handled by isend.
    if (lane < 31 && i0 + RLE_PER_THREAD < N) { /* nextb valid */ }
    const uint32_t tile_last_byte = s_last[31];

    bool redo;
    do {
      redo = false;
      // --- run starts -------------------------------------------------------
      int last_start = -1;
#pragma unroll
      for (int j = 0; j < RLE_PER_THREAD; j++) {
        const uint32_t i = i0 + j;
        const uint32_t pb = j ? b[j - 1] : prevb;
        const bool active = (i < N) && (i >= seg_start);
        if (active && (i == seg_start || b[j] != pb)) last_start = (int)i;
      }
      int tile_max;
      int s = cta_excl_max(last_start, -1, s_wsi, &tile_max);
      s = max(s, run_carry);

      // --- emit counts ------------------------------------------------------
      uint32_t e[RLE_PER_THREAD], kp[RLE_PER_THREAD];
      uint32_t endmask = 0, sum_e = 0;
#pragma unroll
      for (int j = 0; j < RLE_PER_THREAD; j++) {
        const uint32_t i = i0 + j;
        const uint32_t pb = j ? b[j - 1] : prevb;
        const uint32_t nb = (j < RLE_PER_THREAD - 1) ? b[j + 1] : nextb;
        const bool active = (i < N) && (i >= seg_start);
        if (active && (i == seg_start || b[j] != pb)) s = (int)i;
        const uint32_t k = active ? (i - (uint32_t)s) % 259u : 0u;
        const bool isend = (i + 1 >= N) || (nb != b[j]);
        const uint32_t lit = k < 4u;
        const uint32_t cnt = (k == 258u) || (k >= 3u && isend);
        kp[j] = k;
        e[j] = active ? lit + cnt : 0u;
        if (isend) endmask |= 1u << j;
        sum_e += e[j];
      }
      uint32_t tile_total;
      const uint32_t o0 = m + cta_excl_sum(sum_e, s_ws, &tile_total);

      // --- block-full detection (first block only) -----------------------------
      uint32_t cut = 0xFFFFFFFFu;
      if (part == 0 && m + tile_total + 1u >= cap) {
        if (tid == 0) s_cut = 0xFFFFFFFFu;
        __syncthreads();
        uint32_t cum = o0;
#pragma unroll
        for (int j = 0; j < RLE_PER_THREAD; j++) {
          const uint32_t i = i0 + j;
          cum += e[j];
          if (i < N && i >= seg_start) {
            const bool isend = (endmask >> j) & 1u;
            if (cum >= cap || (kp[j] == 2u && cum + 1u == cap && !isend)) atomicMin(&s_cut, i);
          }
        }
        __syncthreads();
        cut = s_cut;
      }

      // --- write literals / count bytes ----------------------------------------
      {
        uint8_t *Tp = part ? Tp1 : Tp0;
        uint32_t o = o0;
#pragma unroll
        for (int j = 0; j < RLE_PER_THREAD; j++) {
          const uint32_t i = i0 + j;
          if (e[j] && i <= cut) {
            const uint32_t lit = kp[j] < 4u;
            if (lit) { Tp[o] = (uint8_t)b[j]; s_used[part][b[j]] = 1; }
            if (e[j] > lit) { Tp[o + lit] = (uint8_t)(kp[j] - 3u); s_used[part][kp[j] - 3u] = 1; }
            if (i == cut) s_ncut = o + e[j];
          }
          o += e[j];
        }
      }

      if (cut != 0xFFFFFFFFu) {
        __syncthreads();
        n0 = s_ncut;
        raw0 = cut + 1u;
        part = 1; seg_start = cut + 1u; m = 0; run_carry = (int)seg_start;
        redo = (seg_start < min(N, tb + RLE_TILE));
        __syncthreads();   // s_cut / s_ncut reuse
      } else {
        m += tile_total;
        run_carry = max(run_carry, tile_max);
      }
    } while (redo);
    prev_tile_byte = tile_last_byte;
  }
  __syncthreads();

  // ---- CRC over the raw bytes of each part ----------------------------------
  {
    const uint32_t c0 = raw0;                         // part 0 = [0,c0), part 1 = [c0,N)
    const uint32_t R = (((N + RLE_THREADS - 1) / RLE_THREADS) + 15u) & ~15u;
    const uint32_t lo = min(tid * R, N), hi = min(lo + R, N);
    const uint32_t aEnd = min(hi, c0), bBeg = max(lo, c0);
    if (lo < aEnd) {
      uint32_t r = 0;
      for (uint32_t i = lo; i < aEnd; i++) r = (r << 8) ^ s_crctab[(r >> 24) ^ x[i]];
      atomicXor(&s_crcacc[0], gf_mulmod(r, gf_xpow8(c0 - aEnd)));
    }
    if (bBeg < hi) {
      uint32_t r = 0;
      for (uint32_t i = bBeg; i < hi; i++) r = (r << 8) ^ s_crctab[(r >> 24) ^ x[i]];
      atomicXor(&s_crcacc[1], gf_mulmod(r, gf_xpow8(N - hi)));
    }
    __syncthreads();
    if (tid < 2) {
      const uint32_t len = tid ? (N - c0) : c0;
      const uint32_t crc = s_crcacc[tid] ^ gf_mulmod(0xFFFFFFFFu, gf_xpow8(len));
      LbzBlockMeta *mt = &meta[2 * c + tid];
      const uint32_t nn = (part == 0) ? (tid ? 0u : m) : (tid ? m : n0);
      mt->n = nn;
      mt->raw_len = len;
      mt->crc = crc;
      mt->bwt_idx = 0; mt->tie_count = 1; mt->nmtf = 0; mt->alpha_size = 0;
      mt->num_trees = 0; mt->num_selectors = 0; mt->tree_pad = 0; mt->out_len = 0;
      mt->unsorted = 0; mt->depth = 0; mt->tree_cost = 0;
    }
    if (tid < 16) {
      const uint32_t p = tid >> 3, wq = tid & 7u;
      uint32_t bits = 0;
      for (int k = 0; k < 32; k++) bits |= (uint32_t)(s_used[p][wq * 32 + k] != 0) << k;
      meta[2 * c + p].used[wq] = bits;
    }
  }
}

// ===========================================================================
// Tile-parallel formulation (default; the one-CTA-per-chunk kernel above is kept
// for comparison, LBZ_RLE_V1=1).  The serial dependencies of collect() are three
// prefix quantities over the raw chunk, each a scan over <= 220 tile aggregates:
//   run start   = last position whose byte differs from its predecessor  (prefix max)
//   output size = emitted bytes so far                                  (prefix sum)
//   block cut   = first position where the block is full                (min)
// so the chunk is processed by independent 4 KiB tiles in a few short launches:
//   breaks -> count -> cut -> write          (block 1: the chunk from position 0)
//   breaks -> count -> write                 (block 2: the rest, runs re-segmented from the cut)
//   crc, final                               (CRC-32 of the raw bytes of either block, block records)
// Every launch covers all chunks of the batch (grid = tiles x chunks).
#define RT_THREADS 256
#define RT_PER 16
#define RT_TILE (RT_THREADS * RT_PER)
#define RT_NOCUT 0xFFFFFFFFFFFFFFFFull

// per-chunk scratch (uint32 words): brk[2][T1] | cnt[2][T1] | cut64 (2) | crc[2] | used[2][8]
__host__ __device__ inline uint32_t rt_stride(uint32_t T1) { return 4u * T1 + 20u; }
struct RtView {
  int *brk0, *brk1;
  uint32_t *cnt0, *cnt1;
  unsigned long long *cut;
  uint32_t *crc, *used;
};
__device__ __forceinline__ RtView rt_view(uint32_t *scr, uint32_t T1, uint32_t c) {
  uint32_t *p = scr + (size_t)c * rt_stride(T1);
  RtView v;
  v.brk0 = reinterpret_cast<int *>(p); v.brk1 = reinterpret_cast<int *>(p + T1);
  v.cnt0 = p + 2 * T1; v.cnt1 = p + 3 * T1;
  v.cut = reinterpret_cast<unsigned long long *>(p + 4 * T1);
  v.crc = p + 4 * T1 + 2; v.used = p + 4 * T1 + 4;
  return v;
}

__host__ __device__ inline uint32_t rt_tail_tile(uint32_t cap);
__global__ void k_rt_init(LbzGeom g, uint32_t *__restrict__ scr) {
  const uint32_t c = blockIdx.x;
  RtView v = rt_view(scr, g.tiles1, c);
  if (threadIdx.x == 0) *v.cut = RT_NOCUT;
  if (threadIdx.x < 2) v.crc[threadIdx.x] = 0;
  if (threadIdx.x < 16) v.used[threadIdx.x] = 0;
  for (uint32_t t = threadIdx.x; t < g.tiles1; t += blockDim.x) { v.brk1[t] = -1; v.cnt1[t] = 0; }   // stage-1 kernels skip the head tiles
}

// Bytes of one tile (16 consecutive per thread, 0x100 = past the end) and their neighbours.
__device__ __forceinline__ void rt_load(const uint8_t *__restrict__ x, uint32_t N, uint32_t i0, uint32_t b[RT_PER],
                                        uint32_t &prevb, uint32_t &nextb) {
  if (i0 + RT_PER <= N && ((reinterpret_cast<uintptr_t>(x + i0) & 15u) == 0)) {
    const uint4 w = *reinterpret_cast<const uint4 *>(x + i0);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < RT_PER; j++) b[j] = (ww[j >> 2] >> (8 * (j & 3))) & 0xFFu;
  } else {
#pragma unroll
    for (int j = 0; j < RT_PER; j++) b[j] = (i0 + j < N) ? x[i0 + j] : 0x100u;
  }
  prevb = (i0 > 0 && i0 <= N) ? x[i0 - 1] : 0x100u;
  nextb = (i0 + RT_PER < N) ? x[i0 + RT_PER] : 0x100u;
}

// Last run start inside this thread's 16 positions (-1 if none); a position is a run
// start if it is the first of the segment or its byte differs from its predecessor.
__device__ __forceinline__ int rt_last_break(const uint32_t b[RT_PER], uint32_t prevb, uint32_t i0, uint32_t N, uint32_t seg) {
  int last = -1;
#pragma unroll
  for (int j = 0; j < RT_PER; j++) {
    const uint32_t i = i0 + j;
    const uint32_t pb = j ? b[j - 1] : prevb;
    if (i < N && i >= seg && (i == seg || b[j] != pb)) last = (int)i;
  }
  return last;
}

// Emitted bytes of every position of the tile (literal: phase < 4; count byte: phase
// 258 or the run ends at phase >= 3), given the run start `carry` in front of the tile.
// Returns the exclusive prefix of this thread inside the tile; *total = tile sum.
__device__ __forceinline__ uint32_t rt_emit(const uint32_t b[RT_PER], uint32_t prevb, uint32_t nextb, uint32_t i0,
                                            uint32_t N, uint32_t seg, int carry, uint32_t kp[RT_PER], uint32_t e[RT_PER],
                                            uint32_t &endmask, uint32_t *ws, int *wsi, uint32_t *total) {
  int tile_max;
  int s = cta_excl_max(rt_last_break(b, prevb, i0, N, seg), -1, wsi, &tile_max);
  s = max(s, carry);
  uint32_t sum_e = 0;
  endmask = 0;
#pragma unroll
  for (int j = 0; j < RT_PER; j++) {
    const uint32_t i = i0 + j;
    const uint32_t pb = j ? b[j - 1] : prevb;
    const uint32_t nb = (j < RT_PER - 1) ? b[j + 1] : nextb;
    const bool active = (i < N) && (i >= seg);
    if (active && (i == seg || b[j] != pb)) s = (int)i;
    const uint32_t k = active ? (i - (uint32_t)s) % 259u : 0u;
    const bool isend = (i + 1 >= N) || (nb != b[j]);
    const uint32_t lit = k < 4u;
    const uint32_t cnt = (k == 258u) || (k >= 3u && isend);
    kp[j] = k;
    e[j] = active ? lit + cnt : 0u;
    if (isend) endmask |= 1u << j;
    sum_e += e[j];
  }
  return cta_excl_sum(sum_e, ws, total);
}

// Count-only variant for k_rt_count: no per-position arrays, one warp reduction and one
// shared atomic instead of a CTA scan.  Returns the tile sum to every thread of warp 0.
__device__ __forceinline__ uint32_t rt_emit_total(const uint32_t b[RT_PER], uint32_t prevb, uint32_t nextb, uint32_t i0,
                                                  uint32_t N, uint32_t seg, int carry, uint32_t *s_total, int *wsi) {
  int tile_max;
  int s = cta_excl_max(rt_last_break(b, prevb, i0, N, seg), -1, wsi, &tile_max);
  s = max(s, carry);
  uint32_t sum_e = 0;
#pragma unroll
  for (int j = 0; j < RT_PER; j++) {
    const uint32_t i = i0 + j;
    const uint32_t pb = j ? b[j - 1] : prevb;
    const uint32_t nb = (j < RT_PER - 1) ? b[j + 1] : nextb;
    const bool active = (i < N) && (i >= seg);
    if (active && (i == seg || b[j] != pb)) s = (int)i;
    const uint32_t k = active ? (i - (uint32_t)s) % 259u : 0u;
    const bool isend = (i + 1 >= N) || (nb != b[j]);
    if (active) sum_e += (k < 4u) + ((k == 258u) || (k >= 3u && isend));
  }
  sum_e = __reduce_add_sync(0xffffffffu, sum_e);
  if ((threadIdx.x & 31u) == 0 && sum_e) atomicAdd(s_total, sum_e);
  __syncthreads();
  return *s_total;
}

// Prefix max of the break table / prefix sum of the count table over the tiles before `tile`.
__device__ __forceinline__ int rt_carry_break(const int *brk, uint32_t tile, int *wsi) {
  int l = -1;
  for (uint32_t t = threadIdx.x; t < tile; t += RT_THREADS) l = max(l, brk[t]);
  int tot;
  (void)cta_excl_max(l, -1, wsi, &tot);
  return tot;
}
__device__ __forceinline__ uint32_t rt_carry_count(const uint32_t *cnt, uint32_t tile, uint32_t *ws) {
  uint32_t c = 0;
  for (uint32_t t = threadIdx.x; t < tile; t += RT_THREADS) c += cnt[t];
  uint32_t tot;
  (void)cta_excl_sum(c, ws, &tot);
  return tot;
}

// Segment start of a stage for chunk c: stage 0 = the chunk start; stage 1 = right after
// the cut (N if the chunk has no second block).
__device__ __forceinline__ uint32_t rt_seg(int stage, const RtView &v, uint32_t N) {
  if (stage == 0) return 0u;
  const unsigned long long cut = *v.cut;
  return cut == RT_NOCUT ? N : (uint32_t)(cut >> 32) + 1u;
}

// RLE1 expands by at most 5/4, so the block can only fill -- and a second block can only
// start -- after 0.8 * cap raw bytes: the cut search and all stage-1 kernels skip the
// tiles in front of that point (their table entries are preset by k_rt_init).
__host__ __device__ inline uint32_t rt_tail_tile(uint32_t cap) {
  const uint32_t t = (cap / 5u) * 4u;                        // capacities below 10 (per-block API, src/encode.c:121) have no skippable head
  return t > 8u ? (t - 8u) / RT_TILE : 0u;
}

template <int STAGE>
__global__ void __launch_bounds__(RT_THREADS)
k_rt_breaks(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr) {
  const uint32_t c = blockIdx.y, tile = blockIdx.x + (STAGE ? rt_tail_tile(g.mbs) : 0u), tid = threadIdx.x;
  const uint32_t N = chunk_len[c];
  RtView v = rt_view(scr, g.tiles1, c);
  int *brk = STAGE ? v.brk1 : v.brk0;
  const uint32_t seg = rt_seg(STAGE, v, N);
  const uint32_t lo = tile * RT_TILE;
  __shared__ int wsi[40];
  if (lo >= N || lo + RT_TILE <= seg) { if (tid == 0) brk[tile] = -1; return; }
  const uint8_t *x = in + (size_t)c * g.mbs;
  uint32_t b[RT_PER], prevb, nextb;
  const uint32_t i0 = lo + tid * RT_PER;
  rt_load(x, N, i0, b, prevb, nextb);
  int tmax;
  (void)cta_excl_max(rt_last_break(b, prevb, i0, N, seg), -1, wsi, &tmax);
  if (tid == 0) brk[tile] = tmax;
}

template <int STAGE>
__global__ void __launch_bounds__(RT_THREADS, 6)
k_rt_count(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr) {
  const uint32_t c = blockIdx.y, tile = blockIdx.x + (STAGE ? rt_tail_tile(g.mbs) : 0u), tid = threadIdx.x;
  const uint32_t N = chunk_len[c];
  RtView v = rt_view(scr, g.tiles1, c);
  const uint32_t seg = rt_seg(STAGE, v, N);
  const uint32_t lo = tile * RT_TILE;
  uint32_t *cnt = STAGE ? v.cnt1 : v.cnt0;
  __shared__ int wsi[40];
  if (lo >= N || lo + RT_TILE <= seg) { if (tid == 0) cnt[tile] = 0; return; }
  __shared__ uint32_t s_total;
  if (tid == 0) s_total = 0;
  const uint8_t *x = in + (size_t)c * g.mbs;
  uint32_t b[RT_PER], prevb, nextb;
  const uint32_t i0 = lo + tid * RT_PER;
  rt_load(x, N, i0, b, prevb, nextb);                              // in flight while the break table is scanned
  const int carry = rt_carry_break(STAGE ? v.brk1 : v.brk0, tile, wsi);
  const uint32_t total = rt_emit_total(b, prevb, nextb, i0, N, seg, carry, &s_total, wsi);
  if (tid == 0) cnt[tile] = total;
}

// Block-full position of block 1 (src/encode.c:162,176,202,218,256-264,276): the first
// position whose cumulative output reaches cap, or that writes the 3rd literal of a
// continuing run with one slot left.  Only tiles that can contain it do any work.
__global__ void __launch_bounds__(RT_THREADS, 4)
k_rt_cut(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr) {
  const uint32_t c = blockIdx.y, tile = blockIdx.x + rt_tail_tile(g.mbs), tid = threadIdx.x;
  const uint32_t N = chunk_len[c], cap = g.mbs;
  const uint32_t lo = tile * RT_TILE;
  if (lo >= N) return;
  RtView v = rt_view(scr, g.tiles1, c);
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  const uint32_t m = rt_carry_count(v.cnt0, tile, ws);
  if (m >= cap || m + v.cnt0[tile] + 1u < cap) return;            // block-uniform
  const int carry = rt_carry_break(v.brk0, tile, wsi);
  const uint8_t *x = in + (size_t)c * g.mbs;
  uint32_t b[RT_PER], kp[RT_PER], e[RT_PER], prevb, nextb, endmask, total;
  const uint32_t i0 = lo + tid * RT_PER;
  rt_load(x, N, i0, b, prevb, nextb);
  uint32_t cum = m + rt_emit(b, prevb, nextb, i0, N, 0u, carry, kp, e, endmask, ws, wsi, &total);
#pragma unroll
  for (int j = 0; j < RT_PER; j++) {
    const uint32_t i = i0 + j;
    cum += e[j];
    if (i < N) {
      const bool isend = (endmask >> j) & 1u;
      if (cum >= cap || (kp[j] == 2u && cum + 1u == cap && !isend)) {
        atomicMin(v.cut, ((unsigned long long)i << 32) | cum);
        break;                                                     // later positions of this thread are larger
      }
    }
  }
}

template <int STAGE>
__global__ void __launch_bounds__(RT_THREADS, 4)
k_rt_write(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr,
           uint8_t *__restrict__ T) {
  const uint32_t c = blockIdx.y, tile = blockIdx.x + (STAGE ? rt_tail_tile(g.mbs) : 0u), tid = threadIdx.x;
  const uint32_t N = chunk_len[c];
  RtView v = rt_view(scr, g.tiles1, c);
  const unsigned long long cut64 = *v.cut;
  const uint32_t seg = rt_seg(STAGE, v, N);
  // stage 0 writes positions up to and including the cut, stage 1 everything after it
  const bool hascut = (STAGE == 0 && cut64 != RT_NOCUT);
  const uint32_t last = hascut ? (uint32_t)(cut64 >> 32) : 0xFFFFFFFFu;
  const uint32_t lo = tile * RT_TILE;
  if (lo >= N || lo + RT_TILE <= seg || lo > last) return;
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  __shared__ uint32_t s_used[8];
  // the tile's output, placed so that shared index == (global offset - (m & ~15)): 16-byte
  // groups of the block text line up with 16-byte groups of the staging buffer
  __shared__ __align__(16) uint8_t s_out[RT_TILE + RT_TILE / 4 + 48];
  if (tid < 8) s_used[tid] = 0;
  const uint8_t *x = in + (size_t)c * g.mbs;
  uint32_t b[RT_PER], kp[RT_PER], e[RT_PER], prevb, nextb, endmask, total;
  const uint32_t i0 = lo + tid * RT_PER;
  rt_load(x, N, i0, b, prevb, nextb);                              // in flight while the tables are scanned
  const int carry = rt_carry_break(STAGE ? v.brk1 : v.brk0, tile, wsi);
  const uint32_t m = rt_carry_count(STAGE ? v.cnt1 : v.cnt0, tile, ws);
  const uint32_t shift = m & 15u;
  uint32_t o = shift + rt_emit(b, prevb, nextb, i0, N, seg, carry, kp, e, endmask, ws, wsi, &total);
  uint32_t ub[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < RT_PER; j++) {
    const uint32_t i = i0 + j;
    if (e[j] && i <= last) {
      const uint32_t lit = kp[j] < 4u;
      if (lit) { s_out[o] = (uint8_t)b[j]; ub[b[j] >> 5] |= 1u << (b[j] & 31u); }
      if (e[j] > lit) { const uint32_t cb = kp[j] - 3u; s_out[o + lit] = (uint8_t)cb; ub[cb >> 5] |= 1u << (cb & 31u); }
    }
    o += e[j];
  }
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint32_t r = __reduce_or_sync(0xffffffffu, ub[w]);
    if ((tid & 31u) == 0 && r) atomicOr(&s_used[w], r);
  }
  __syncthreads();
  if (tid < 8 && s_used[tid]) atomicOr(&v.used[STAGE * 8 + tid], s_used[tid]);
  // bytes of this tile that belong to the block: all of them, or up to the cut
  const uint32_t nout = hascut ? min(total, (uint32_t)cut64 - m) : total;
  uint8_t *Tp = T + lbz_slot_off(g, 2 * c + STAGE) + (m - shift);          // 16-byte aligned (slots are)
  const uint32_t beg = shift, end = shift + nout;                            // staged range [beg, end)
  const uint32_t vbeg = (beg + 15u) & ~15u, vend = end & ~15u;
  if (vbeg < vend) {
    for (uint32_t q = vbeg / 16 + tid; q < vend / 16; q += RT_THREADS)
      reinterpret_cast<uint4 *>(Tp)[q] = reinterpret_cast<const uint4 *>(s_out)[q];
    for (uint32_t q = beg + tid; q < vbeg; q += RT_THREADS) Tp[q] = s_out[q];
    for (uint32_t q = vend + tid; q < end; q += RT_THREADS) Tp[q] = s_out[q];
  } else {
    for (uint32_t q = beg + tid; q < end; q += RT_THREADS) Tp[q] = s_out[q];
  }
}

// CRC-32/BZIP2 of the raw bytes of either block.  A CTA takes a 64 KiB span of the chunk,
// stages it with coalesced loads (row of 64 words per thread, padded to 65 so that the
// per-thread serial reads hit 32 different banks), every thread runs the table CRC over
// its 256 bytes and shifts the result to the end of its block by x^(8 * distance) in GF(2).
#define RC_SPAN 65536u
#define RC_ROW 65u
__global__ void __launch_bounds__(256)
k_rt_crc(LbzGeom g, const uint8_t *__restrict__ in, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr) {
  extern __shared__ __align__(16) uint32_t rc_smem[];            // 256 rows x 65 words
  const uint32_t c = blockIdx.y, tid = threadIdx.x;
  const uint32_t N = chunk_len[c];
  const uint32_t span0 = blockIdx.x * RC_SPAN;
  if (span0 >= N) return;
  RtView v = rt_view(scr, g.tiles1, c);
  const unsigned long long cut64 = *v.cut;
  const uint32_t c0 = cut64 == RT_NOCUT ? N : (uint32_t)(cut64 >> 32) + 1u;     // block 1 = [0,c0), block 2 = [c0,N)
  const uint8_t *x = in + (size_t)c * g.mbs;
  __shared__ uint32_t s_crctab[256];
  __shared__ uint32_t s_acc[2];
  {
    uint32_t r = tid << 24;
#pragma unroll
    for (int k = 0; k < 8; k++) r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
    s_crctab[tid] = r;
  }
  if (tid < 2) s_acc[tid] = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(x + span0) & 15u) == 0;
#pragma unroll 4
  for (uint32_t k = 0; k < RC_SPAN / 16 / 256; k++) {
    const uint32_t q = k * 256 + tid;                            // 16-byte group of the span
    const uint32_t p = span0 + q * 16;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (p + 16 <= N && aligned) {
      w = *reinterpret_cast<const uint4 *>(x + p);
    } else if (p < N) {
      uint32_t ww[4] = {0, 0, 0, 0};
      for (uint32_t j = 0; j < 16 && p + j < N; j++) ww[j >> 2] |= (uint32_t)x[p + j] << (8 * (j & 3));
      w = make_uint4(ww[0], ww[1], ww[2], ww[3]);
    }
    uint32_t *row = rc_smem + (q >> 4) * RC_ROW + (q & 15u) * 4;
    row[0] = w.x; row[1] = w.y; row[2] = w.z; row[3] = w.w;
  }
  __syncthreads();
  const uint32_t lo = min(span0 + tid * 256u, N), hi = min(lo + 256u, N);
  const uint32_t aEnd = min(hi, c0), bBeg = max(lo, c0);
  const uint32_t *row = rc_smem + tid * RC_ROW;
  if (lo < aEnd) {
    uint32_t r = 0;
    for (uint32_t i = lo; i < aEnd; i++) {
      const uint32_t byte = (row[(i - lo) >> 2] >> (8 * ((i - lo) & 3u))) & 0xFFu;
      r = (r << 8) ^ s_crctab[(r >> 24) ^ byte];
    }
    atomicXor(&s_acc[0], gf_mulmod(r, gf_xpow8(c0 - aEnd)));
  }
  if (bBeg < hi) {
    uint32_t r = 0;
    for (uint32_t i = bBeg; i < hi; i++) {
      const uint32_t byte = (row[(i - lo) >> 2] >> (8 * ((i - lo) & 3u))) & 0xFFu;
      r = (r << 8) ^ s_crctab[(r >> 24) ^ byte];
    }
    atomicXor(&s_acc[1], gf_mulmod(r, gf_xpow8(N - hi)));
  }
  __syncthreads();
  if (tid < 2 && s_acc[tid]) atomicXor(&v.crc[tid], s_acc[tid]);
}

__global__ void __launch_bounds__(256)
k_rt_final(LbzGeom g, const uint32_t *__restrict__ chunk_len, uint32_t *__restrict__ scr, LbzBlockMeta *__restrict__ meta) {
  const uint32_t c = blockIdx.x, tid = threadIdx.x;
  const uint32_t N = chunk_len[c];
  RtView v = rt_view(scr, g.tiles1, c);
  __shared__ uint32_t ws[40];
  const unsigned long long cut64 = *v.cut;
  const uint32_t ntiles = (N + RT_TILE - 1) / RT_TILE;
  const uint32_t tot0 = rt_carry_count(v.cnt0, ntiles, ws);
  const uint32_t tot1 = rt_carry_count(v.cnt1, ntiles, ws);
  if (tid < 2) {
    const bool hascut = cut64 != RT_NOCUT;
    const uint32_t c0 = hascut ? (uint32_t)(cut64 >> 32) + 1u : N;
    const uint32_t len = tid ? (N - c0) : c0;
    const uint32_t nn = tid ? (hascut ? tot1 : 0u) : (hascut ? (uint32_t)cut64 : tot0);
    LbzBlockMeta *mt = &meta[2 * c + tid];
    mt->n = nn;
    mt->raw_len = len;
    mt->crc = v.crc[tid] ^ gf_mulmod(0xFFFFFFFFu, gf_xpow8(len));
    mt->bwt_idx = 0; mt->tie_count = 1; mt->nmtf = 0; mt->alpha_size = 0;
    mt->num_trees = 0; mt->num_selectors = 0; mt->tree_pad = 0; mt->out_len = 0;
    mt->unsorted = 0; mt->depth = 0; mt->tree_cost = 0;
  }
  if (tid >= 32 && tid < 48) {
    const uint32_t q = tid - 32;
    meta[2 * c + (q >> 3)].used[q & 7u] = v.used[q];
  }
}

extern "C" size_t lbz_rle1_scratch_words(const LbzGeom *g, uint32_t max_chunks) {
  return (size_t)max_chunks * rt_stride(g->tiles1);
}

extern "C" int lbz_launch_rle1_tiles(const LbzGeom *g, const uint8_t *d_in, const uint32_t *d_chunk_len, uint8_t *d_T,
                                     LbzBlockMeta *d_meta, uint32_t *d_scratch, cudaStream_t st) {
  if (g->nchunks == 0) return 0;
  const uint32_t ntiles = (g->mbs + RT_TILE - 1) / RT_TILE, tail0 = rt_tail_tile(g->mbs);
  const dim3 grid(ntiles, g->nchunks), grid_tail(ntiles - tail0, g->nchunks);
  const size_t crc_smem = 256 * RC_ROW * sizeof(uint32_t);
  LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_rt_crc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)crc_smem));
  k_rt_init<<<g->nchunks, 256, 0, st>>>(*g, d_scratch);
  k_rt_breaks<0><<<grid, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_count<0><<<grid, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_cut<<<grid_tail, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_write<0><<<grid, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch, d_T);
  k_rt_breaks<1><<<grid_tail, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_count<1><<<grid_tail, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_write<1><<<grid_tail, RT_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_scratch, d_T);
  k_rt_crc<<<dim3((g->mbs + RC_SPAN - 1) / RC_SPAN, g->nchunks), 256, crc_smem, st>>>(*g, d_in, d_chunk_len, d_scratch);
  k_rt_final<<<g->nchunks, 256, 0, st>>>(*g, d_chunk_len, d_scratch, d_meta);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int lbz_launch_rle1(const LbzGeom *g, const uint8_t *d_in, const uint32_t *d_chunk_len,
                               uint8_t *d_T, LbzBlockMeta *d_meta, cudaStream_t st) {
  if (g->nchunks == 0) return 0;
  k_rle1<<<g->nchunks, RLE_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_T, d_meta);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}
