// huffman.cu -- multi-table prefix-code construction, one CTA per block.
//
// Replaces generate_prefix_code() (reference src/encode.c:1005-1137) with its
// helpers generate_initial_trees (:779-841), find_best_tree (:846-877),
// make_code_lengths (:713-766), assign_codes (:882-987), package_merge
// (:660-710), and the selector MTF / padding / size arithmetic of encode()
// (:460-544).
//
// Parallel structure inside the CTA (512 threads):
//   * E-step: one thread per 50-symbol group; the six candidate costs are
//     accumulated in one u64 with 10-bit fields exactly like the reference
//     (field overflow carries into the next field and must be reproduced);
//     histograms by shared-memory atomics.
//   * M-step: symbols are ordered by parallel rank counting, the two-queue
//     Huffman merge is run by one thread per tree (six trees in six warps).
//   * Final codes: package-merge in its list form.  List N_1 is the sorted
//     leaves, N_{k+1} = merge(leaves, pairs of N_k).  One warp per tree merges
//     each level with binary searches (merge path), records a leaf/package
//     bit per item, and every lane then evaluates one height limit h by
//     walking down the lists with popcounts.  On equal weight a leaf sorts
//     before a package (the reference encodes this in bits 24..31 of its keys,
//     encode.c:652-654).  Package weights saturate at 2^32-1, which keeps
//     every leaf/package comparison exact because no leaf reaches that value.
//   * Selector MTF: 6-symbol recency lists composed with a warp scan.
#include "lbz_common.cuh"

#define HUF_THREADS 512
#define HUF_WARPS 16
#define HUF_MAXH 20


struct HufSmem {
  unsigned long long len_pack[260];
  uint32_t freq[LBZ_MAX_TREES][260];
  uint32_t mtffreq[260];
  uint32_t hw[LBZ_MAX_TREES][520];         // Huffman node weights / PM merged list
  uint32_t pm_leaf[LBZ_MAX_TREES][260];
  uint32_t pm_pkg[LBZ_MAX_TREES][260];
  uint32_t pm_bits[LBZ_MAX_TREES][HUF_MAXH + 1][17];
  uint32_t hcnt[LBZ_MAX_TREES][32];
  uint32_t tree_cost[LBZ_MAX_TREES];
  uint32_t first_use[LBZ_MAX_TREES];
  uint32_t old2new[LBZ_MAX_TREES], new2old[LBZ_MAX_TREES];
  uint32_t ws[40];
  uint32_t misc[8];
  uint16_t ord[LBZ_MAX_TREES][260];
  uint16_t hparent[LBZ_MAX_TREES][520];
  uint16_t pm_ge[LBZ_MAX_TREES][HUF_MAXH + 1][HUF_MAXH + 2];
  uint8_t length[LBZ_MAX_TREES][260];
  uint8_t hdepth[LBZ_MAX_TREES][520];
  uint8_t pm_len[LBZ_MAX_TREES][HUF_MAXH + 1][260];
  uint8_t selector[18008];
  // E-step staging: every warp copies the 32 groups it is about to score (32 x 50 u16 = 800
  // words, contiguous in HBM) with coalesced 128-bit loads; lane l then reads its group at a
  // stride of 25 words, which hits 32 different banks (25 is odd).
  __align__(16) uint32_t stage[HUF_THREADS / 32][800];
};

// Sort the alphabet of every tree: heaviest first, ties by lower symbol
// (order produced by sort_alphabet() on keys freq<<32 | 1<<16 | 258-sym,
// encode.c:553-567,739-740,901-902).  floor1: use max(freq,1) (EM-loop keys).
__device__ void rank_symbols(HufSmem &S, uint32_t ntrees, uint32_t as, bool floor1) {
  for (uint32_t idx = threadIdx.x; idx < ntrees * as; idx += HUF_THREADS) {
    const uint32_t t = idx / as, i = idx - t * as;
    uint32_t fi = S.freq[t][i];
    if (floor1 && fi == 0) fi = 1;
    uint32_t r = 0;
    for (uint32_t j = 0; j < as; j++) {
      uint32_t fj = S.freq[t][j];
      if (floor1 && fj == 0) fj = 1;
      r += (fj > fi) || (fj == fi && j < i);
    }
    S.ord[t][r] = (uint16_t)i;
  }
}

// Two-queue Huffman for tree t by a single thread; writes S.length[t][0..as).
__device__ void huffman_serial(HufSmem &S, uint32_t t, uint32_t as) {
  uint32_t *w = S.hw[t];
  uint16_t *par = S.hparent[t];
  uint8_t *dep = S.hdepth[t];
  for (uint32_t k = 0; k < as; k++) {
    const uint32_t f = S.freq[t][S.ord[t][as - 1 - k]];
    w[k] = f ? f : 1u;                       // ascending weights
  }
  uint32_t li = 0, ii = as, ni = as;
  while (ni < 2 * as - 1) {
    uint32_t pick[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      bool leaf;
      if (li >= as) leaf = false;
      else if (ii >= ni) leaf = true;
      else leaf = w[li] <= w[ii];            // tie: leaf before internal node
      pick[q] = leaf ? li++ : ii++;
    }
    w[ni] = w[pick[0]] + w[pick[1]];
    par[pick[0]] = (uint16_t)ni;
    par[pick[1]] = (uint16_t)ni;
    ni++;
  }
  uint32_t *cnt = S.hcnt[t];
  for (int d = 0; d < 32; d++) cnt[d] = 0;
  dep[ni - 1] = 0;
  for (int node = (int)ni - 2; node >= 0; node--) {
    const uint32_t d = dep[par[node]] + 1u;
    dep[node] = (uint8_t)d;
    if (node < (int)as) cnt[d < 31 ? d : 31]++;
  }
  // heaviest symbol gets the shallowest leaf (encode.c:750-763): deal out the
  // multiset of leaf depths in increasing order over the sorted symbols.
  uint8_t *len = S.length[t];
  uint32_t s = 0;
  for (uint32_t d = 1; d < 32; d++)
    for (uint32_t c = cnt[d]; c > 0; c--) len[S.ord[t][s++]] = (uint8_t)d;
}

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {
  const uint32_t s = a + b;
  return s < a ? 0xFFFFFFFFu : s;
}

// Length-limited codes for tree t by one warp (package-merge, list form).
// Returns the cost (payload + tree transmission) in S.tree_cost[t] and leaves
// the final lengths in S.length[t] and the canonical codes in code_out.
__device__ void limited_codes_warp(HufSmem &S, uint32_t t, uint32_t as, uint32_t *code_out) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t *leaf = S.pm_leaf[t], *pkg = S.pm_pkg[t], *merged = S.hw[t];
  for (uint32_t a = lane; a < as; a += 32) {
    const uint32_t f = S.freq[t][S.ord[t][as - 1 - a]];
    leaf[a] = f;
    merged[a] = f;
  }
  for (uint32_t i = lane; i < (HUF_MAXH + 1) * 17; i += 32) (&S.pm_bits[t][0][0])[i] = 0;
  __syncwarp();
  uint32_t cnt = as;
  for (uint32_t k = 2; k <= HUF_MAXH; k++) {
    const uint32_t np = cnt >> 1;
    for (uint32_t bI = lane; bI < np; bI += 32) pkg[bI] = sat_add(merged[2 * bI], merged[2 * bI + 1]);
    __syncwarp();
    // leaves: position = a + #packages strictly lighter
    for (uint32_t a = lane; a < as; a += 32) {
      const uint32_t v = leaf[a];
      uint32_t lo = 0, hi = np;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (pkg[mid] < v) lo = mid + 1; else hi = mid; }
      merged[a + lo] = v;
    }
    // packages: position = b + #leaves not heavier
    for (uint32_t bI = lane; bI < np; bI += 32) {
      const uint32_t v = pkg[bI];
      uint32_t lo = 0, hi = as;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (leaf[mid] <= v) lo = mid + 1; else hi = mid; }
      const uint32_t pos = bI + lo;
      merged[pos] = v;
      atomicOr(&S.pm_bits[t][k][pos >> 5], 1u << (pos & 31u));
    }
    cnt = as + np;
    __syncwarp();
  }

  // each lane evaluates one height limit
  const uint32_t h = lane;
  const bool feasible = (h >= 2 && h <= HUF_MAXH && (1u << h) >= as);
  uint32_t ge_top = 1;
  if (feasible) {
    uint32_t take = 2 * as - 2;
    for (uint32_t k = h; k >= 1; k--) {
      uint32_t pk = 0;
      const uint32_t *bits = S.pm_bits[t][k];
      const uint32_t full = take >> 5, rem = take & 31u;
      for (uint32_t wq = 0; wq < full; wq++) pk += __popc(bits[wq]);
      if (rem) pk += __popc(bits[full] & ((1u << rem) - 1u));
      const uint32_t leaves = take - pk;
      S.pm_ge[t][h][h - k + 1] = (uint16_t)leaves;
      take = 2 * pk;
    }
    S.pm_ge[t][h][h + 1] = 0;
    ge_top = S.pm_ge[t][h][h];
  }
  // first feasible height whose longest length is unused ends the search (encode.c:916-920)
  const uint32_t stopmask = __ballot_sync(0xffffffffu, feasible && ge_top == 0);
  const uint32_t hstop = stopmask ? (uint32_t)(__ffs(stopmask) - 1) : 32u;
  const bool evaluate = feasible && h < hstop;
  uint32_t cost = 0xFFFFFFFFu;
  if (evaluate) {
    uint8_t *len = S.pm_len[t][h];
    uint32_t c = 0, s = 0;
    for (uint32_t d = 1; d <= h; d++) {
      const uint32_t nd = (uint32_t)S.pm_ge[t][h][d] - (uint32_t)S.pm_ge[t][h][d + 1];
      for (uint32_t q = 0; q < nd; q++) {
        const uint32_t sym = S.ord[t][s++];
        len[sym] = (uint8_t)d;
        c += S.freq[t][sym] * d;
      }
    }
    for (uint32_t v = 1; v < as; v++) {
      const int dl = (int)len[v] - (int)len[v - 1];
      c += 2u * (uint32_t)(dl < 0 ? -dl : dl);
    }
    cost = c + 5u + as;
  }
  // min cost, lowest height on ties
  uint32_t best = evaluate ? ((cost << 5) | h) : 0xFFFFFFFFu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  const uint32_t bh = best & 31u;
  __syncwarp();
  for (uint32_t v = lane; v < as; v += 32) S.length[t][v] = S.pm_len[t][bh][v];
  __syncwarp();
  if (lane == 0) {
    S.tree_cost[t] = best >> 5;
    uint32_t nlen[HUF_MAXH + 2], next[HUF_MAXH + 2];
#pragma unroll
    for (int d = 0; d < HUF_MAXH + 2; d++) nlen[d] = 0;
    for (uint32_t v = 0; v < as; v++) nlen[S.length[t][v]]++;
    uint32_t c = 0;
    for (uint32_t d = 1; d <= HUF_MAXH; d++) { next[d] = c; c = (c + nlen[d]) << 1; }
    for (uint32_t v = 0; v < as; v++) code_out[v] = next[S.length[t][v]]++;
    code_out[as] = 0;
    S.length[t][as] = 0;
  }
  __syncwarp();
}

// --- selector MTF over a 6-symbol alphabet -----------------------------------
// A recency summary is a nibble-packed list of distinct symbols, most recent
// first, with the count in bits 28..31.
__device__ __forceinline__ uint32_t rec_push(uint32_t s, uint32_t c) {   // c becomes most recent
  const uint32_t cnt = s >> 28;
  uint32_t out = c, k = 1;
  for (uint32_t i = 0; i < cnt; i++) {
    const uint32_t x = (s >> (4 * i)) & 0xFu;
    if (x != c) { out |= x << (4 * k); k++; }
  }
  return out | (k << 28);
}
__device__ __forceinline__ uint32_t rec_combine(uint32_t first, uint32_t then) {  // apply `then` after `first`
  const uint32_t c2 = then >> 28;
  uint32_t out = then & 0x0FFFFFFFu, k = c2;
  uint32_t seen = 0;
  for (uint32_t i = 0; i < c2; i++) seen |= 1u << ((then >> (4 * i)) & 0xFu);
  const uint32_t c1 = first >> 28;
  for (uint32_t i = 0; i < c1; i++) {
    const uint32_t x = (first >> (4 * i)) & 0xFu;
    if (!(seen & (1u << x))) { out |= x << (4 * k); k++; }
  }
  return out | (k << 28);
}

__global__ void __launch_bounds__(HUF_THREADS, 1)
k_huffman(LbzGeom g, LbzBlockMeta *__restrict__ meta, uint16_t *__restrict__ mtfv_all,
          const uint32_t *__restrict__ freq_all, LbzCoding *__restrict__ coding_all,
          uint32_t cluster_factor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HufSmem &S = *reinterpret_cast<HufSmem *>(smem_raw);
  const uint32_t b = blockIdx.x;
  const uint32_t n = meta[b].n;
  if (n == 0) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t nm = meta[b].nmtf, as = meta[b].alpha_size;
  const uint16_t *mtfv = mtfv_all + lbz_slot_off(g, b);
  LbzCoding &C = coding_all[b];
  const uint32_t ng = (nm + LBZ_GROUP - 1) / LBZ_GROUP;
  uint32_t nt = nm > 2400 ? 6 : nm > 1200 ? 5 : nm > 600 ? 4 : nm > 300 ? 3 : nm > 150 ? 2 : 1;

  for (uint32_t i = tid; i < 260; i += HUF_THREADS) S.mtffreq[i] = freq_all[b * 260 + i];
  for (uint32_t i = tid; i < LBZ_MAX_TREES * 260; i += HUF_THREADS) (&S.length[0][0])[i] = 1;
  __syncthreads();

  // ---- initial classes (encode.c:779-841), serial: <= 258 steps ------------
  if (tid == 0) {
    uint32_t live = 0, cum = 0, a, rem = nm, ntl = nt;
    for (a = 0; cum < nm; a++) { cum += S.mtffreq[a]; live += S.mtffreq[a] != 0; }
    if (ntl > live) ntl = live;
    a = 0;
    for (uint32_t t = 0; ntl > 0; t++, ntl--) {
      uint32_t f = S.mtffreq[a], e = a + 1;
      cum = f;
      live -= f != 0;
      while (live > ntl - 1 && cum * ntl < rem) { f = S.mtffreq[e]; cum += f; live -= f != 0; e++; }
      if (cum > f && (2 * cum - f) * ntl > 2 * rem) { cum -= f; live += f != 0; e--; }
      for (uint32_t v = a; v < e; v++) S.length[t][v] = 0;
      a = e;
      rem -= cum;
    }
  }
  __syncthreads();

  // ---- EM iterations (encode.c:1043-1084) ------------------------------------
  for (uint32_t iter = 0; iter < cluster_factor; iter++) {
    for (uint32_t v = tid; v <= as; v += HUF_THREADS) {
      unsigned long long p = 0;
      if (v < as)
        for (uint32_t t = 0; t < LBZ_MAX_TREES; t++) p += (unsigned long long)S.length[t][v] << (10 * t);
      S.len_pack[v] = p;
    }
    for (uint32_t i = tid; i < LBZ_MAX_TREES * 260; i += HUF_THREADS) (&S.freq[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t g0 = 0; g0 < ng; g0 += HUF_THREADS) {            // warp-uniform trip count
      const uint32_t gw = g0 + warp * 32;                          // first group of this warp
      const uint32_t gi = gw + lane;
      uint32_t *stg = S.stage[warp];
      if (gw < ng) {
        // only the block's own (padded) groups are read: the slot ends at most 64 symbols later
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(mtfv + (size_t)gw * LBZ_GROUP);
        const uint32_t lim = (ng - gw) * LBZ_GROUP;                // symbols of this warp's valid groups
#pragma unroll
        for (int q = 0; q < 7; q++) {
          const uint32_t i4 = q * 32 + lane;
          if (i4 < 200 && i4 * 8u < lim) reinterpret_cast<uint4 *>(stg)[i4] = gsrc[i4];
        }
      }
      __syncwarp();
      if (gi < ng) {
        const uint32_t *gp = stg + lane * (LBZ_GROUP / 2);
        uint32_t wv[LBZ_GROUP / 2];
        unsigned long long cp = 0;
#pragma unroll
        for (int i = 0; i < LBZ_GROUP / 2; i++) {
          wv[i] = gp[i];
          cp += S.len_pack[wv[i] & 0xFFFFu];
          cp += S.len_pack[wv[i] >> 16];
        }
        uint32_t bc = (uint32_t)cp & 0x3FFu, bt = 0;
        for (uint32_t t = 1; t < nt; t++) {
          cp >>= 10;
          const uint32_t c = (uint32_t)cp & 0x3FFu;
          if (c < bc) { bc = c; bt = t; }
        }
        S.selector[gi] = (uint8_t)bt;
#pragma unroll
        for (int i = 0; i < LBZ_GROUP / 2; i++) {
          atomicAdd(&S.freq[bt][wv[i] & 0xFFFFu], 1u);
          atomicAdd(&S.freq[bt][wv[i] >> 16], 1u);
        }
      }
      __syncwarp();                                                // the staging words are reused next trip
    }
    __syncthreads();
    rank_symbols(S, nt, as, true);
    __syncthreads();
    if (lane == 0 && warp < nt) huffman_serial(S, warp, as);
    __syncthreads();
  }

  // ---- tree order by first use (encode.c:1088-1111) -----------------------------
  if (tid < LBZ_MAX_TREES) { S.first_use[tid] = 0xFFFFFFFFu; S.tree_cost[tid] = 0; }
  __syncthreads();
  for (uint32_t gi = tid; gi < ng; gi += HUF_THREADS) atomicMin(&S.first_use[S.selector[gi]], gi);
  __syncthreads();
  if (tid == 0) {
    uint32_t nn = 0;
    for (uint32_t t = 0; t < LBZ_MAX_TREES; t++) S.old2new[t] = 0xFFu;
    for (;;) {
      uint32_t bestt = 0xFFu, bestg = 0xFFFFFFFFu;
      for (uint32_t t = 0; t < nt; t++)
        if (S.old2new[t] == 0xFFu && S.first_use[t] < bestg) { bestg = S.first_use[t]; bestt = t; }
      if (bestt == 0xFFu) break;
      S.old2new[bestt] = nn; S.new2old[nn] = bestt; nn++;
    }
    S.misc[0] = nn;
  }
  __syncthreads();
  uint32_t nn = S.misc[0];
  rank_symbols(S, nt, as, false);
  __syncthreads();
  // one warp per used tree; codes go straight to the NEW slot of the output
  if (warp < nt && S.old2new[warp] != 0xFFu)
    limited_codes_warp(S, warp, as, C.code[S.old2new[warp]]);
  __syncthreads();

  uint32_t cost = 0;
  for (uint32_t t = 0; t < nt; t++) if (S.old2new[t] != 0xFFu) cost += S.tree_cost[t];
  if (nn == 1) {                                    // dummy second tree (encode.c:1117-1132)
    const uint32_t t = S.new2old[0] ^ 1u;
    uint32_t cl0 = 0;
    while ((2u << cl0) <= as) cl0++;
    const uint32_t nshort = (2u << cl0) - as;
    for (uint32_t v = tid; v < as; v += HUF_THREADS) S.length[t][v] = (uint8_t)(v < nshort ? cl0 : cl0 + 1);
    for (uint32_t v = tid; v <= as; v += HUF_THREADS) C.code[1][v] = 0;
    if (tid == 0) { S.old2new[t] = 1; S.new2old[1] = t; S.length[t][as] = 0; }
    if (nshort < as) cost += 2;
    cost += as + 5;
    nn = 2;
  }
  __syncthreads();
  for (uint32_t i = tid; i < nn * 260; i += HUF_THREADS) {
    const uint32_t t = i / 260, v = i - t * 260;
    C.length[t][v] = (v <= as) ? S.length[S.new2old[t]][v] : 0;
  }

  // ---- selectors: renumber, MTF, cost (encode.c:473-512) --------------------------
  const uint32_t per = (ng + HUF_THREADS - 1) / HUF_THREADS;
  const uint32_t g0 = min(tid * per, ng), g1 = min(g0 + per, ng);
  uint32_t summ = 0;
  for (uint32_t gi = g0; gi < g1; gi++) {
    const uint32_t c = S.old2new[S.selector[gi]];
    S.selector[gi] = (uint8_t)c;
    summ = rec_push(summ, c);
  }
  // inclusive scan of summaries across the CTA
  uint32_t inc = summ;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t prev = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc = rec_combine(prev, inc);
  }
  __syncthreads();
  if (lane == 31) S.ws[warp] = inc;
  __syncthreads();
  uint32_t before = 0;                               // everything before this thread's segment
  for (uint32_t w = 0; w < warp; w++) before = rec_combine(before, S.ws[w]);
  {
    uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane > 0) before = rec_combine(before, prev);
  }
  // start list = recency summary, then the untouched symbols in increasing order (initial list 0..5)
  uint32_t list[LBZ_MAX_TREES];
  {
    uint32_t k = 0, seen = 0;
    const uint32_t cb = before >> 28;
    for (uint32_t i = 0; i < cb; i++) { const uint32_t x = (before >> (4 * i)) & 0xFu; list[k++] = x; seen |= 1u << x; }
    for (uint32_t x = 0; x < LBZ_MAX_TREES; x++) if (!(seen & (1u << x))) list[k++] = x;
  }
  uint32_t selbits = 0;
  for (uint32_t gi = g0; gi < g1; gi++) {
    const uint32_t c = S.selector[gi];
    uint32_t j = 0;
#pragma unroll
    for (int q = 0; q < LBZ_MAX_TREES; q++) if (list[q] == c) j = q;
#pragma unroll
    for (int q = LBZ_MAX_TREES - 1; q > 0; q--) if ((uint32_t)q <= j) list[q] = list[q - 1];
    list[0] = c;
    C.selector[gi] = (uint8_t)c;
    C.selector_mtf[gi] = (uint8_t)j;
    selbits += j + 1;
  }
  uint32_t tot;
  (void)cta_excl_sum(selbits, S.ws, &tot);

  if (tid == 0) {
    uint32_t bits = 48 + 32 + 1 + 24 + 3 + 15 + cost + tot;
    const uint32_t pad = (8u - (bits & 7u)) & 7u;
    bits += pad;
    uint32_t rows = 0;
    for (uint32_t r = 0; r < 16; r++) {
      const uint32_t wd = meta[b].used[r >> 1];
      rows += ((wd >> (16 * (r & 1u))) & 0xFFFFu) != 0;
    }
    bits += 16 + 16 * rows;
    meta[b].num_trees = nn;
    meta[b].tree_pad = pad >> 1;
    meta[b].num_selectors = ng + (pad & 1u);
    meta[b].tree_cost = cost;
    meta[b].out_len = bits >> 3;
    if (pad & 1u) C.selector_mtf[ng] = 0;
  }
}

extern "C" size_t lbz_coding_bytes() { return sizeof(LbzCoding); }

extern "C" int lbz_launch_huffman(const LbzGeom *g, LbzBlockMeta *d_meta, uint16_t *d_mtfv, const uint32_t *d_freq,
                                  void *d_coding, uint32_t cluster_factor, cudaStream_t st) {
  const uint32_t nb = 2 * g->nchunks;
  if (nb == 0) return 0;
  LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_huffman, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HufSmem)));
  k_huffman<<<nb, HUF_THREADS, sizeof(HufSmem), st>>>(*g, d_meta, d_mtfv, d_freq,
                                                      reinterpret_cast<LbzCoding *>(d_coding), cluster_factor);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}
