// pack.cu -- bit serialisation of one block per CTA, then gather into stream order.
//
// Replaces transmit() (reference src/encode.c:1152-1281): block magic, ~CRC,
// rand bit, primary index, used-byte bitmap, tree count, selector count,
// unary selector MTF values, delta-coded code lengths, then 50 codes per group.
// Everything is expressed as a stream of (length <= 32, value) items; item
// offsets are a prefix sum over lengths, each item is OR-ed into a shared-
// memory staging window at its bit position (MSB first) and complete 32-bit
// words are flushed, byte-swapped, with coalesced stores.  Blocks end on a byte
// boundary by construction (encode.c:514-525), so k_gather can concatenate
// them as bytes in stream order -- the only cross-block step of the path.
#include "lbz_common.cuh"

#define PACK_THREADS 1024
#define PACK_PER 4
#define PACK_TILE (PACK_THREADS * PACK_PER)
#define PACK_WORDS (PACK_TILE + 8)


struct PackSmem {
  uint32_t buf[PACK_WORDS];
  uint32_t code[LBZ_MAX_TREES][260];
  uint32_t ws[40];
  uint32_t rowmap[17];             // [0] = 16-bit row presence, [1..16] = rows
  uint8_t length[LBZ_MAX_TREES][260];
};

struct PackState {
  uint32_t bitpos;                 // bits emitted so far (CTA-uniform)
  uint32_t *out32;
};

// OR `len` bits of `val` (right-aligned) into the staging window at absolute
// bit position P; `wbase` is the absolute word index of buf[0].
__device__ __forceinline__ void put_bits(uint32_t *buf, uint32_t wbase, uint32_t P, uint32_t len, uint32_t val) {
  if (len == 0) return;
  const uint32_t w = (P >> 5) - wbase, q = P & 31u;
  const int sh = 32 - (int)q - (int)len;
  if (sh >= 0) {
    atomicOr(&buf[w], val << sh);
  } else {
    atomicOr(&buf[w], val >> (-sh));
    atomicOr(&buf[w + 1], val << (32 + sh));
  }
}

// One phase: `count` items produced by gen(i, len&, val&), PACK_PER consecutive
// items per thread per tile.
template <class Gen>
__device__ void emit_phase(PackSmem &S, PackState &st, uint32_t count, Gen gen) {
  const uint32_t tid = threadIdx.x;
  for (uint32_t base = 0; base < count; base += PACK_TILE) {
    uint32_t len[PACK_PER], val[PACK_PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PACK_PER; q++) {
      const uint32_t i = base + tid * PACK_PER + q;
      len[q] = 0; val[q] = 0;
      if (i < count) gen(i, len[q], val[q]);
      sum += len[q];
    }
    uint32_t tot;
    uint32_t P = st.bitpos + cta_excl_sum(sum, S.ws, &tot);
    const uint32_t wbase = st.bitpos >> 5;
#pragma unroll
    for (int q = 0; q < PACK_PER; q++) {
      put_bits(S.buf, wbase, P, len[q], val[q]);
      P += len[q];
    }
    __syncthreads();
    const uint32_t newpos = st.bitpos + tot;
    const uint32_t nfull = (newpos >> 5) - wbase;        // complete words
    for (uint32_t w = tid; w < nfull; w += PACK_THREADS) st.out32[wbase + w] = __byte_perm(S.buf[w], 0, 0x0123);
    const uint32_t carry = S.buf[nfull];
    __syncthreads();
    for (uint32_t w = tid; w <= nfull + 1 && w < PACK_WORDS; w += PACK_THREADS) S.buf[w] = (w == 0) ? carry : 0u;
    __syncthreads();
    st.bitpos = newpos;
  }
}

__global__ void __launch_bounds__(PACK_THREADS, 1)
k_pack(LbzGeom g, LbzBlockMeta *__restrict__ meta, const uint16_t *__restrict__ mtfv_all,
       const LbzCoding *__restrict__ coding_all, uint8_t *__restrict__ out_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PackSmem &S = *reinterpret_cast<PackSmem *>(smem_raw);
  const uint32_t b = blockIdx.x;
  const LbzBlockMeta mt = meta[b];
  if (mt.n == 0) return;
  const uint32_t tid = threadIdx.x;
  const LbzCoding &C = coding_all[b];
  const uint16_t *mtfv = mtfv_all + lbz_slot_off(g, b);
  const uint32_t as = mt.alpha_size, nt = mt.num_trees;
  const uint32_t ng = (mt.nmtf + LBZ_GROUP - 1) / LBZ_GROUP;

  for (uint32_t i = tid; i < PACK_WORDS; i += PACK_THREADS) S.buf[i] = 0;
  for (uint32_t i = tid; i < LBZ_MAX_TREES * 260; i += PACK_THREADS) {
    (&S.code[0][0])[i] = (&C.code[0][0])[i];
    (&S.length[0][0])[i] = (&C.length[0][0])[i];
  }
  if (tid < 16) {
    const uint32_t wd = mt.used[tid >> 1];
    const uint32_t half = (wd >> (16 * (tid & 1u))) & 0xFFFFu;   // bit v = byte 16*tid+v used
    S.rowmap[1 + tid] = __brev(half) >> 16;                      // MSB = byte 16*tid (encode.c:1205-1207)
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t big = 0;
    for (int r = 0; r < 16; r++) big = (big << 1) | (S.rowmap[1 + r] != 0);
    S.rowmap[0] = big;
  }
  __syncthreads();

  PackState st;
  st.bitpos = 0;
  st.out32 = reinterpret_cast<uint32_t *>(out_all + (size_t)b * g.out_cap);

  // ---- header (encode.c:1185-1223) ----------------------------------------------
  {
    const uint32_t crc = mt.crc ^ 0xFFFFFFFFu, idx = mt.bwt_idx, nsel = mt.num_selectors;
    emit_phase(S, st, 28, [&](uint32_t i, uint32_t &len, uint32_t &val) {
      switch (i) {
        case 0: len = 24; val = 0x314159u; break;
        case 1: len = 24; val = 0x265359u; break;
        case 2: len = 16; val = crc >> 16; break;
        case 3: len = 16; val = crc & 0xFFFFu; break;
        case 4: len = 1; val = 0; break;
        case 5: len = 24; val = idx; break;
        case 6: len = 16; val = S.rowmap[0]; break;
        case 23: len = 3; val = nt; break;
        case 24: len = 15; val = nsel; break;
        default:
          if (i >= 7 && i < 23) { val = S.rowmap[1 + (i - 7)]; len = val ? 16 : 0; }
          break;
      }
    });
  }
  // ---- selectors: unary MTF values (encode.c:1224-1228) ------------------------------
  emit_phase(S, st, mt.num_selectors, [&](uint32_t i, uint32_t &len, uint32_t &val) {
    const uint32_t j = C.selector_mtf[i];
    len = j + 1; val = (1u << (j + 1)) - 2u;
  });
  // ---- trees: 5-bit start, then per symbol (10)* / (11)* and a 0 (encode.c:1231-1255) ---
  {
    const uint32_t tree_pad = mt.tree_pad;
    emit_phase(S, st, nt * as * 3, [&](uint32_t i, uint32_t &len, uint32_t &val) {
      const uint32_t part = i % 3, tv = i / 3;
      const uint32_t t = tv / as, v = tv - t * as;
      const int cur = S.length[t][v];
      int prev;
      if (v == 0) {
        prev = cur;
        if (t == 0) prev += (cur < 4) ? (int)tree_pad : -(int)tree_pad;
        if (part == 0) { len = 5; val = (uint32_t)prev; return; }
      } else {
        if (part == 0) return;
        prev = S.length[t][v - 1];
      }
      const bool inc = prev < cur;
      const uint32_t pairs = (uint32_t)(inc ? cur - prev : prev - cur);
      const uint32_t pa = min(pairs, 15u), pb = pairs - pa;
      // p pairs of '10' = 0xAAAAAAAA >> (32-2p); of '11' = 0xFFFFFFFF >> (32-2p)
      const uint32_t pat = inc ? 0xAAAAAAAAu : 0xFFFFFFFFu;
      if (part == 1) { len = 2 * pa; val = pa ? (pat >> (32 - 2 * pa)) : 0u; }
      else { len = 2 * pb + 1; val = pb ? ((pat >> (32 - 2 * pb)) << 1) : 0u; }
    });
  }
  // ---- symbols (encode.c:1258-1272) --------------------------------------------------
  emit_phase(S, st, ng * LBZ_GROUP, [&](uint32_t i, uint32_t &len, uint32_t &val) {
    const uint32_t sym = mtfv[i];
    const uint32_t t = C.selector[i / LBZ_GROUP];
    len = S.length[t][sym]; val = S.code[t][sym];
  });
  // ---- flush the last partial word ------------------------------------------------------
  if (tid == 0) {
    if (st.bitpos & 31u) st.out32[st.bitpos >> 5] = __byte_perm(S.buf[0], 0, 0x0123);
    meta[b].pad_[0] = st.bitpos;      // host checks bitpos == 8 * out_len
  }
}

// ---------------------------------------------------------------------------
// Stream-order offsets of the packed blocks, then a byte gather.
__global__ void __launch_bounds__(1024)
k_out_offsets(uint32_t nblocks, const LbzBlockMeta *__restrict__ meta, uint32_t *__restrict__ out_off,
              uint32_t *__restrict__ total) {
  __shared__ uint32_t ws[40];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < nblocks && meta[i].n) ? meta[i].out_len : 0u;
    uint32_t tot;
    const uint32_t ex = cta_excl_sum(v, ws, &tot);
    if (i < nblocks) out_off[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256)
k_gather(LbzGeom g, const LbzBlockMeta *__restrict__ meta, const uint32_t *__restrict__ out_off,
         const uint8_t *__restrict__ out_all, uint8_t *__restrict__ packed) {
  const uint32_t b = blockIdx.y;
  if (meta[b].n == 0) return;
  const uint32_t len = meta[b].out_len;
  const uint8_t *src = out_all + (size_t)b * g.out_cap;
  uint8_t *dst = packed + out_off[b];
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < len; i += gridDim.x * 256) dst[i] = src[i];
}

extern "C" int lbz_launch_pack(const LbzGeom *g, LbzBlockMeta *d_meta, const uint16_t *d_mtfv, const void *d_coding,
                               uint8_t *d_out, uint32_t *d_out_off, uint32_t *d_total, uint8_t *d_packed,
                               cudaStream_t st) {
  const uint32_t nb = 2 * g->nchunks;
  if (nb == 0) return 0;
  LBZ_CUDA_CHECK(cudaFuncSetAttribute(k_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PackSmem)));
  k_pack<<<nb, PACK_THREADS, sizeof(PackSmem), st>>>(*g, d_meta, d_mtfv, reinterpret_cast<const LbzCoding *>(d_coding), d_out);
  k_out_offsets<<<1, 1024, 0, st>>>(nb, d_meta, d_out_off, d_total);
  k_gather<<<dim3(64, nb), 256, 0, st>>>(*g, d_meta, d_out_off, d_out, d_packed);
  LBZ_CUDA_CHECK(cudaGetLastError());
  return 0;
}
