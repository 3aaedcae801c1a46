// mtf.cu -- move-to-front + zero-run coding + symbol histogram.
//
// Replaces do_mtf() (reference src/encode.c:360-425) and make_map_e()
// (src/encode.c:340-355).  Two kernels, one CTA per block:
//
//  k_mtf_ranks  computes the MTF rank of every BWT position.  The serial
//     dependency (the recency list) is broken per 2048-symbol warp segment:
//     the list in front of a segment is "all symbols ordered by their last
//     occurrence before the segment", so each warp records the last
//     occurrence of every symbol in its segment, a prefix-max over the 32
//     warps (+ the carry of the previous super-tile) gives each warp its own
//     start list (bitonic sort of 256 keys in registers), and the warp then
//     walks its segment with the 256-entry list held as 8 bytes per lane
//     (__vcmpeq4 + ballot to find a symbol, byte shifts + one shuffle to move
//     it to the front).  Positions equal to their predecessor are rank 0 by
//     construction and are skipped.
//
//  k_mtf_emit   turns (rank, zero-run) into the u16 symbol stream: RUNA/RUNB
//     digits of each zero run (bijective base 2, encode.c:381-386), rank+1
//     otherwise, EOB at the end, plus the histogram that seeds the prefix-code
//     clustering.  Run starts are a prefix max, output offsets a prefix sum;
//     one CTA streams the block with carried state, like rle1.cu.
#include "lbz_common.cuh"
#include <climits>

#define MTF_THREADS 1024
#define MTF_WARPS 32
#define MTF_SEG 1024u                       // symbols per warp per super-tile
#define MTF_PARTS 8u                        // CTAs per block (each walks a contiguous range of super-tiles)
#define MTF_SUPER (MTF_SEG * MTF_WARPS)

__device__ __forceinline__ void cmpx_desc(uint32_t &a, uint32_t &b) {   // a >= b afterwards
  const uint32_t hi = max(a, b), lo = min(a, b);
  a = hi; b = lo;
}

__device__ __forceinline__ void mtf_part_range(uint32_t n, uint32_t part, uint32_t &lo, uint32_t &hi) {
  const uint32_t nsuper = (n + MTF_SUPER - 1) / MTF_SUPER;
  const uint32_t spp = (nsuper + MTF_PARTS - 1) / MTF_PARTS;
  lo = min(part * spp * MTF_SUPER, n);
  hi = min((part + 1) * spp * MTF_SUPER, n);
}

__device__ __forceinline__ uint32_t dense_of(const uint32_t *used, uint32_t v) {
  uint32_t cnt = 0;
  for (uint32_t w = 0; w < 8; w++) {
    const uint32_t bits = used[w];
    if (w < (v >> 5)) cnt += __popc(bits);
    else if (w == (v >> 5)) cnt += __popc(bits & ((1u << (v & 31u)) - 1u));
  }
  return cnt;
}

// Last occurrence of every (dense) symbol inside each part of a block, so that
// the MTF walk of part p can start from the recency order left by parts < p.
__global__ void __launch_bounds__(MTF_THREADS)
k_mtf_parttab(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ bwt,
              int *__restrict__ parttab) {
  const uint32_t b = blockIdx.y, part = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  uint32_t lo, hi;
  mtf_part_range(n, part, lo, hi);
  const uint8_t *src = bwt + lbz_slot_off(g, b);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  __shared__ int s_tab[256];
  __shared__ uint8_t s_dense[256];
  if (tid < 256) { s_tab[tid] = INT_MIN; s_dense[tid] = (uint8_t)dense_of(meta[b].used, tid); }
  __syncthreads();
  for (uint32_t base = lo; base < hi; base += MTF_THREADS) {      // uniform trip count per warp
    const uint32_t p = base + tid;
    const bool valid = p < hi;
    const uint32_t c = valid ? s_dense[src[p]] : 0xFFFFu;
    const uint32_t nc = __shfl_down_sync(0xffffffffu, c, 1);
    if (valid && (lane == 31 || c != nc)) atomicMax(&s_tab[c], (int)p);
  }
  __syncthreads();
  if (tid < 256) parttab[((size_t)b * MTF_PARTS + part) * 256 + tid] = s_tab[tid];
}

// One batch of up to 32 events (lane i = i-th event, `valid` lanes only).  All
// ranks of the batch are computed together from the batch-local occurrence
// structure:
//   * symbol seen earlier in the batch (at lane j): rank = number of distinct
//     symbols in lanes (j, i) = lanes there that are the last occurrence before i
//     of their symbol ("alive" mask, a prefix-OR of predecessor bits);
//   * first occurrence in the batch: rank = P[sym] + number of distinct batch
//     symbols seen earlier whose old place was behind it;
// then the position table P (P[sym] = place of sym in the recency list) is advanced
// past the batch in one step: symbols outside the batch slide back by the number of
// batch symbols that were behind them (256-bit occupancy bitmap SB + suffix counts).
__device__ __forceinline__ void mtf_batch(uint8_t *P, uint32_t *SB, uint32_t csym, bool valid, uint32_t pos,
                                          uint8_t *__restrict__ dstr, uint32_t ltm, uint32_t gtm, uint32_t lane) {
  const uint32_t c = valid ? csym : (0x100u + lane);                   // invalid lanes: unique dummies
  const uint32_t m = __match_any_sync(0xffffffffu, c);
  const uint32_t prevm = m & ltm;
  const bool hasprev = prevm != 0;
  const uint32_t j = hasprev ? (31u - __clz(prevm)) : 0u;
  const bool isfirst = valid && !hasprev;
  const bool islast = valid && ((m & gtm) == 0);
  const uint32_t Pc = valid ? (uint32_t)P[c] : 0u;
  uint32_t dead = hasprev ? (1u << j) : 0u;                            // lanes whose symbol occurs again at or before this lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, dead, o);
    if (lane >= (uint32_t)o) dead |= t;
  }
  const uint32_t alive = ltm & ~dead;
  const uint32_t fm = __ballot_sync(0xffffffffu, isfirst);
  uint32_t cnt = 0;
  for (uint32_t mm = fm; mm; mm &= mm - 1) {
    const uint32_t t = __ffs(mm) - 1;
    const uint32_t pt = __shfl_sync(0xffffffffu, Pc, t);
    cnt += (t < lane) && (pt > Pc);
  }
  const uint32_t rank = hasprev ? __popc((alive >> j) >> 1) : (Pc + cnt);
  if (valid) dstr[pos] = (uint8_t)rank;

  const uint32_t lastm = __ballot_sync(0xffffffffu, islast);
  uint32_t Bw[8];
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint32_t contrib = (isfirst && (Pc >> 5) == (uint32_t)w) ? (1u << (Pc & 31u)) : 0u;
    Bw[w] = __reduce_or_sync(0xffffffffu, contrib);
  }
  if (lane < 8) {
    uint32_t mine = 0, suffix = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      if ((uint32_t)w == lane) mine = Bw[w];
      if ((uint32_t)w > lane) suffix += __popc(Bw[w]);
    }
    SB[lane] = mine;
    SB[8 + lane] = suffix;
  }
  __syncwarp();
  {
    uint2 pv = *reinterpret_cast<uint2 *>(&P[lane * 8]);
    uint32_t wv[2] = {pv.x, pv.y};
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      uint32_t out = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t pk = (wv[h2] >> (8 * k)) & 0xFFu;
        const uint32_t w = pk >> 5;
        const uint32_t above = __popc((SB[w] >> (pk & 31u)) >> 1) + SB[8 + w];
        out |= ((pk + above) & 0xFFu) << (8 * k);
      }
      wv[h2] = out;
    }
    pv.x = wv[0]; pv.y = wv[1];
    *reinterpret_cast<uint2 *>(&P[lane * 8]) = pv;
  }
  __syncwarp();
  if (islast) P[c] = (uint8_t)__popc(lastm & gtm);
  __syncwarp();
}

struct MtfSmem {
  int tab[MTF_WARPS][256];
  int carry[256];
  uint32_t B[MTF_WARPS][16];
  uint32_t qpos[MTF_WARPS][64];
  uint16_t qsym[MTF_WARPS][64];
  __align__(8) uint8_t P[MTF_WARPS][256];
  uint8_t dense[256];
};

__global__ void __launch_bounds__(MTF_THREADS, 2)
k_mtf_ranks(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ bwt,
            uint8_t *__restrict__ mtfrank, const int *__restrict__ parttab) {
  const uint32_t b = blockIdx.y, part = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  uint32_t part_lo, part_hi;
  mtf_part_range(n, part, part_lo, part_hi);
  if (part_lo >= part_hi) return;
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *src = bwt + off;
  uint8_t *dstr = mtfrank + off;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  extern __shared__ __align__(16) unsigned char mtf_smem_raw[];
  MtfSmem &SM = *reinterpret_cast<MtfSmem *>(mtf_smem_raw);
  int (*s_tab)[256] = SM.tab;
  int *s_carry = SM.carry;
  uint8_t *s_dense = SM.dense;
  uint8_t (*s_P)[256] = SM.P;
  uint32_t (*s_B)[16] = SM.B;
  uint16_t (*s_qsym)[64] = SM.qsym;
  uint32_t (*s_qpos)[64] = SM.qpos;

  if (tid < 256) {
    // dense renumbering of the used byte values (encode.c:340-355)
    uint32_t cnt = 0;
    for (uint32_t w = 0; w < 8; w++) {
      const uint32_t bits = meta[b].used[w];
      if (w < (tid >> 5)) cnt += __popc(bits);
      else if (w == (tid >> 5)) cnt += __popc(bits & ((1u << (tid & 31u)) - 1u));
    }
    s_dense[tid] = (uint8_t)cnt;
    int cr = -1 - (int)tid;                // never seen: identity order
    for (uint32_t q = 0; q < part; q++) cr = max(cr, parttab[((size_t)b * MTF_PARTS + q) * 256 + tid]);
    s_carry[tid] = cr;
  }

  for (uint32_t sbase = part_lo; sbase < part_hi; sbase += MTF_SUPER) {
    __syncthreads();
    for (uint32_t i = tid; i < MTF_WARPS * 256; i += MTF_THREADS) (&s_tab[0][0])[i] = INT_MIN;
    __syncthreads();
    const uint32_t segbase = sbase + warp * MTF_SEG;

    // (A) last occurrence of every symbol inside this warp's segment
    if (segbase < part_hi) {
      for (uint32_t q = 0; q < MTF_SEG / 32; q++) {
        const uint32_t p = segbase + q * 32 + lane;
        if (segbase + q * 32 >= part_hi) break;
        const bool valid = p < part_hi;
        const uint32_t c = valid ? s_dense[src[p]] : 0xFFFFu;
        const uint32_t nc = __shfl_down_sync(0xffffffffu, c, 1);
        if (valid && (lane == 31 || c != nc)) atomicMax(&s_tab[warp][c], (int)p);
      }
    }
    __syncthreads();
    // (B) exclusive prefix max over warps, seeded with the carry
    if (tid < 256) {
      int run = s_carry[tid];
#pragma unroll 4
      for (int w = 0; w < MTF_WARPS; w++) {
        const int t = s_tab[w][tid];
        s_tab[w][tid] = run;
        run = max(run, t);
      }
      s_carry[tid] = run;
    }
    __syncthreads();
    if (segbase >= part_hi) continue;

    // (C) start list = symbols by descending last occurrence
    uint32_t e[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t c = lane * 8 + r;
      e[r] = ((uint32_t)(s_tab[warp][c] + 512) << 8) | c;
    }
    // bitonic sort of 256 elements, blocked layout (index = lane*8 + r), descending
#pragma unroll
    for (uint32_t k = 2; k <= 256; k <<= 1) {
#pragma unroll
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        if (j >= 8) {
          const uint32_t lj = j >> 3;
          const bool upper = (lane & lj) != 0;                       // partner has the lower index
          const bool desc = ((lane * 8) & k) == 0;                   // this k-block sorts descending
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, e[r], lj);
            const bool keep_max = (desc != upper);
            e[r] = keep_max ? max(e[r], o) : min(e[r], o);
          }
        } else {
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int pr = r ^ (int)j;
            if (pr > r) {
              const bool desc = (((lane * 8 + r) & k) == 0);
              if (desc) cmpx_desc(e[r], e[pr]); else cmpx_desc(e[pr], e[r]);
            }
          }
        }
      }
    }
    // position table P[sym] = place of sym in the recency list at the segment start
    uint8_t *P = s_P[warp];
#pragma unroll
    for (int r = 0; r < 8; r++) P[e[r] & 0xFFu] = (uint8_t)(lane * 8 + r);
    __syncwarp();

    // Walk the segment.  Positions equal to their predecessor have rank 0 and leave
    // the list alone, so only the others ("events") are queued (per warp, in shared
    // memory) and processed 32 at a time by mtf_batch().
    const uint32_t ltm = lanemask_lt();
    const uint32_t gtm = ~ltm & ~(1u << lane);
    uint16_t *qsym = s_qsym[warp];
    uint32_t *qpos = s_qpos[warp];
    uint32_t qn = 0;                                                   // queued events (warp-uniform)
    uint32_t prev_sym = segbase ? s_dense[src[segbase - 1]] : 0u;    // list front before the segment
    for (uint32_t q = 0; q < MTF_SEG / 32; q++) {
      if (segbase + q * 32 >= part_hi) break;
      const uint32_t p = segbase + q * 32 + lane;
      const bool valid = p < part_hi;
      const uint32_t c = valid ? s_dense[src[p]] : 0xFFFFu;
      uint32_t pc = __shfl_up_sync(0xffffffffu, c, 1);
      if (lane == 0) pc = prev_sym;
      const bool ev = valid && (c != pc);
      const uint32_t evmask = __ballot_sync(0xffffffffu, ev);
      prev_sym = __shfl_sync(0xffffffffu, c, 31);
      if (valid && !ev) dstr[p] = 0;
      if (ev) {
        const uint32_t slot = qn + __popc(evmask & ltm);
        qsym[slot] = (uint16_t)c;
        qpos[slot] = p;
      }
      qn += __popc(evmask);
      __syncwarp();
      if (qn >= 32) {
        mtf_batch(P, s_B[warp], qsym[lane], true, qpos[lane], dstr, ltm, gtm, lane);
        const uint32_t rest = qn - 32;                               // < 32: move the tail to the front
        const uint16_t ts = (lane < rest) ? qsym[32 + lane] : (uint16_t)0;
        const uint32_t tp = (lane < rest) ? qpos[32 + lane] : 0u;
        __syncwarp();
        if (lane < rest) { qsym[lane] = ts; qpos[lane] = tp; }
        qn = rest;
        __syncwarp();
      }
    }
    if (qn) mtf_batch(P, s_B[warp], (lane < qn) ? (uint32_t)qsym[lane] : 0u, lane < qn, qpos[lane < qn ? lane : 0], dstr, ltm, gtm, lane);
  }
}

// ---------------------------------------------------------------------------
#define EMIT_THREADS 256
#define EMIT_PER 8
#define EMIT_TILE (EMIT_THREADS * EMIT_PER)

#define EMIT_PARTS 16u

__device__ __forceinline__ void emit_part_range(uint32_t n, uint32_t part, uint32_t &lo, uint32_t &hi) {
  const uint32_t ntiles = (n + EMIT_TILE - 1) / EMIT_TILE;
  const uint32_t tpp = (ntiles + EMIT_PARTS - 1) / EMIT_PARTS;
  lo = min(part * tpp * EMIT_TILE, n);
  hi = min((part + 1) * tpp * EMIT_TILE, n);
}

// Walk positions [lo, hi) of a block: zero runs -> RUNA/RUNB digits, other
// positions -> rank+1.  WRITE = false only counts the emitted symbols (first
// pass, to learn where each part starts in the output); WRITE = true emits them
// and accumulates the histogram.  `carry_nz` = last position before lo whose
// rank is not zero (-1 if none), `m` = output index at lo.
template <bool WRITE>
__device__ uint32_t emit_walk(const uint8_t *__restrict__ src, const uint8_t *__restrict__ rk, uint16_t *__restrict__ out,
                              uint32_t n, uint32_t lo, uint32_t hi, bool first_is_zero, int carry_nz, uint32_t m,
                              uint32_t *s_freq, uint32_t *ws, int *wsi) {
  const uint32_t tid = threadIdx.x;
  for (uint32_t tb = lo; tb < hi; tb += EMIT_TILE) {
    const uint32_t p0 = tb + tid * EMIT_PER;
    uint32_t c[EMIT_PER + 2];   // bytes p0-1 .. p0+8
#pragma unroll
    for (int j = 0; j < EMIT_PER + 2; j++) {
      const int64_t p = (int64_t)p0 + j - 1;
      c[j] = (p >= 0 && p < (int64_t)n) ? src[p] : 0x100u;
    }
    uint32_t zmask = 0;
    int last = -1;
#pragma unroll
    for (int j = 0; j < EMIT_PER; j++) {
      const uint32_t p = p0 + j;
      if (p < hi) {
        const bool z = p ? (c[j + 1] == c[j]) : first_is_zero;
        if (z) zmask |= 1u << j; else last = (int)p;
      }
    }
    int tmax;
    int lnz = cta_excl_max(last, -1, wsi, &tmax);
    lnz = max(lnz, carry_nz);
    uint32_t ec[EMIT_PER], kk[EMIT_PER];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < EMIT_PER; j++) {
      const uint32_t p = p0 + j;
      ec[j] = 0; kk[j] = 0;
      if (p < hi) {
        if (zmask & (1u << j)) {
          const bool runend = (p + 1 >= n) || (c[j + 2] != c[j + 1]);   // a run is emitted where it ends
          if (runend) {
            const uint32_t k = (uint32_t)((int)p - lnz);
            kk[j] = k;
            ec[j] = 31u - __clz(k + 1u);
          }
        } else {
          lnz = (int)p;
          ec[j] = 1;
        }
      }
      sum += ec[j];
    }
    uint32_t tot;
    uint32_t o = m + cta_excl_sum(sum, ws, &tot);
    if (WRITE) {
      uint32_t runa = 0, runb = 0;
#pragma unroll
      for (int j = 0; j < EMIT_PER; j++) {
        const uint32_t p = p0 + j;
        uint32_t sym = 0xFFFFu;                           // histogram key of this lane (none)
        if (p < hi && ec[j]) {
          if (zmask & (1u << j)) {
            const uint32_t v = kk[j] + 1u, nd = ec[j];
            for (uint32_t d = 0; d < nd; d++) out[o + d] = (uint16_t)((v >> d) & 1u);
            const uint32_t ones = __popc(v & ((1u << nd) - 1u));
            runb += ones; runa += nd - ones;
          } else {
            sym = (uint32_t)rk[p] + 1u;
            out[o] = (uint16_t)sym;
          }
          o += ec[j];
        }
        // warp-aggregated histogram update: MTF ranks are heavily skewed, so many
        // lanes hit the same bin
        const uint32_t mm = __match_any_sync(0xffffffffu, sym);
        if (sym != 0xFFFFu && (mm & lanemask_lt()) == 0) atomicAdd(&s_freq[sym], (uint32_t)__popc(mm));
      }
      runa = __reduce_add_sync(0xffffffffu, runa);
      runb = __reduce_add_sync(0xffffffffu, runb);
      if ((threadIdx.x & 31u) == 0) {
        if (runa) atomicAdd(&s_freq[0], runa);
        if (runb) atomicAdd(&s_freq[1], runb);
      }
    }
    m += tot;
    carry_nz = max(carry_nz, tmax);
  }
  return m;
}

// Last position before `lo` whose symbol differs from its predecessor (-1 if none):
// the start of the zero run that may reach into this part.
__device__ int emit_find_carry(const uint8_t *__restrict__ src, uint32_t lo, bool first_is_zero, int *wsi) {
  int found = -1;
  for (int64_t top = (int64_t)lo; top > 0 && found < 0; top -= EMIT_THREADS) {
    const int64_t p = top - 1 - threadIdx.x;
    int cand = -1;
    if (p >= 0) {
      const bool z = p ? (src[p] == src[p - 1]) : first_is_zero;
      if (!z) cand = (int)p;
    }
    int tmax;
    (void)cta_excl_max(cand, -1, wsi, &tmax);
    found = tmax;
  }
  return found;
}

template <bool WRITE>
__global__ void __launch_bounds__(EMIT_THREADS)
k_mtf_emit(LbzGeom g, LbzBlockMeta *__restrict__ meta, const uint8_t *__restrict__ bwt,
           const uint8_t *__restrict__ mtfrank, uint16_t *__restrict__ mtfv,
           uint32_t *__restrict__ freq_out, uint32_t *__restrict__ part_count) {
  const uint32_t b = blockIdx.y, part = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  uint32_t lo, hi;
  emit_part_range(n, part, lo, hi);
  const uint32_t off = lbz_slot_off(g, b);
  const uint8_t *src = bwt + off;
  uint16_t *out = mtfv + off;
  const uint32_t tid = threadIdx.x;

  __shared__ uint32_t s_freq[LBZ_MAX_ALPHA + 2];
  __shared__ uint32_t ws[40];
  __shared__ int wsi[40];
  __shared__ uint32_t s_first_dense;
  for (uint32_t i = tid; i < LBZ_MAX_ALPHA + 2; i += EMIT_THREADS) s_freq[i] = 0;
  if (!WRITE && part == 0)
    for (uint32_t i = tid; i < 260; i += EMIT_THREADS) freq_out[b * 260 + i] = 0;
  if (tid == 0) s_first_dense = dense_of(meta[b].used, src[0]);   // position 0 is a zero iff it is symbol 0
  __syncthreads();
  const bool first_is_zero = (s_first_dense == 0);
  if (lo >= hi) {
    if (!WRITE && tid == 0) part_count[b * EMIT_PARTS + part] = 0;
    return;
  }
  const int carry_nz = emit_find_carry(src, lo, first_is_zero, wsi);
  uint32_t m0 = 0;
  if (WRITE) for (uint32_t q = 0; q < part; q++) m0 += part_count[b * EMIT_PARTS + q];
  const uint32_t m = emit_walk<WRITE>(src, mtfrank + off, out, n, lo, hi, first_is_zero, carry_nz, m0, s_freq, ws, wsi);
  if (!WRITE) {
    if (tid == 0) part_count[b * EMIT_PARTS + part] = m;
    return;
  }
  __syncthreads();
  for (uint32_t i = tid; i < LBZ_MAX_ALPHA + 1; i += EMIT_THREADS)
    if (s_freq[i]) atomicAdd(&freq_out[b * 260 + i], s_freq[i]);
  if (hi == n) {                                          // the part that holds the end of the block
    uint32_t ninuse = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) ninuse += __popc(meta[b].used[w]);
    const uint32_t eob = ninuse + 1u, as = ninuse + 2u;
    const uint32_t nm = m + 1u;
    const uint32_t padded = ((nm + LBZ_GROUP - 1) / LBZ_GROUP) * LBZ_GROUP;
    if (tid == 0) { out[m] = (uint16_t)eob; atomicAdd(&freq_out[b * 260 + eob], 1u); }
    if (tid >= 1 && m + tid < padded) out[m + tid] = (uint16_t)as;      // group padding (encode.c:1034)
    if (tid == 0) { meta[b].nmtf = nm; meta[b].alpha_size = as; }
  }
}

extern "C" uint32_t lbz_mtf_parts() { return MTF_PARTS; }

extern "C" int lbz_launch_mtf(const LbzGeom *g, LbzBlockMeta *d_meta, const uint8_t *d_bwt, uint8_t *d_mtfrank,
                              uint16_t *d_mtfv, uint32_t *d_freq, int *d_parttab, uint32_t *d_emitcnt, cudaStream_t st) {
  const uint32_t nb = 2 * g->nchunks;
  if (nb == 0) return 0;
  k_mtf_parttab<<<dim3(MTF_PARTS, nb), MTF_THREADS, 0, st>>>(*g, d_meta, d_bwt, d_parttab);
  LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_mtf_ranks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MtfSmem)));
  k_mtf_ranks<<<dim3(MTF_PARTS, nb), MTF_THREADS, sizeof(MtfSmem), st>>>(*g, d_meta, d_bwt, d_mtfrank, d_parttab);
  k_mtf_emit<false><<<dim3(EMIT_PARTS, nb), EMIT_THREADS, 0, st>>>(*g, d_meta, d_bwt, d_mtfrank, d_mtfv, d_freq, d_emitcnt);
  k_mtf_emit<true><<<dim3(EMIT_PARTS, nb), EMIT_THREADS, 0, st>>>(*g, d_meta, d_bwt, d_mtfrank, d_mtfv, d_freq, d_emitcnt);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}
