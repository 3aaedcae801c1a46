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

extern "C" int lbz_launch_rle1(const LbzGeom *g, const uint8_t *d_in, const uint32_t *d_chunk_len,
                               uint8_t *d_T, LbzBlockMeta *d_meta, cudaStream_t st) {
  if (g->nchunks == 0) return 0;
  k_rle1<<<g->nchunks, RLE_THREADS, 0, st>>>(*g, d_in, d_chunk_len, d_T, d_meta);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}
