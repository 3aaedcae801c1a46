// lbz_common.cuh -- shared definitions for the B200 bzip2 block-compression engine.
//
// Data layout in HBM (see DESIGN.md "Data layout"):
//   A batch holds up to `nchunks` raw input chunks of `mbs` bytes
//   (mbs = level*100000, the scheduler's in_granul, reference process.c:631).
//   Each chunk yields 1..2 bzip2 blocks (reference compress.c:93-110), so a
//   chunk owns two fixed "block slots": slot 2c (capacity S1 >= mbs+64) and
//   slot 2c+1 (capacity S2 >= mbs/4+64; the spill block can never be larger
//   because RLE1 expands by at most 5/4).  All per-element arrays (text,
//   rotation order, ranks, sort keys, BWT, MTF symbols) are indexed by
//   element offset  c*(S1+S2) + part*S1 + i , so tile -> block mapping is pure
//   arithmetic and every kernel can be launched before block sizes are known
//   on the host.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define LBZ_TILE 4096u          // elements per sort tile (256 threads x 16)
#define LBZ_MAX_ALPHA 258
#define LBZ_GROUP 50
#define LBZ_MAX_TREES 6
#define LBZ_MAX_SEL 18002

struct LbzGeom {
  uint32_t mbs;       // max block size = raw chunk size
  uint32_t S1, S2;    // slot capacities (multiples of LBZ_TILE)
  uint32_t stride;    // S1 + S2
  uint32_t tiles1;    // S1 / LBZ_TILE
  uint32_t nchunks;   // chunks in this batch
  uint32_t out_cap;   // per-slot capacity of the packed bitstream (bytes)
};

__host__ __device__ inline uint32_t lbz_slot_off(const LbzGeom &g, uint32_t b) {
  return (b >> 1) * g.stride + (b & 1) * g.S1;
}
__host__ __device__ inline uint32_t lbz_slot_cap(const LbzGeom &g, uint32_t b) {
  return (b & 1) ? g.S2 : g.S1;
}

// Per-block record, written by the kernels, read back by the host.
struct LbzBlockMeta {
  uint32_t n;            // n' = RLE1 output length (0 = slot unused)
  uint32_t raw_len;      // raw bytes consumed by this block
  uint32_t crc;          // un-inverted CRC-32/BZIP2 of the raw bytes (encode.c:542)
  uint32_t bwt_idx;      // primary index (first position of the tie group)
  uint32_t tie_count;    // > 1 iff the block is exactly periodic
  uint32_t nmtf;
  uint32_t alpha_size;   // EOB + 1
  uint32_t num_trees;
  uint32_t num_selectors;
  uint32_t tree_pad;
  uint32_t out_len;      // bytes of the packed block
  uint32_t unsorted;     // rotations still in tied groups (sort bookkeeping)
  uint32_t depth;        // prefix length the current order is valid for
  uint32_t tree_cost;    // bits: sum over trees of payload + tree transmission
  uint32_t used[8];      // 256-bit used-byte map, bit v of word v/32
  uint32_t pad_[2];      // [0] = bits written by k_pack (host cross-check)
  // tied-rotation lists of the refinement rounds: list S holds groups of <= 32
  // members (sorted locally), list L the larger ones (radix passes); S lives at
  // list index [0, us), L at [lbase, lbase + ul); *_next are committed between rounds
  uint32_t us, ul, lbase, us_next, ul_next, lbase_next;
};

// Prefix-code description of one block (global memory, one per block slot).
struct LbzCoding {
  uint8_t length[LBZ_MAX_TREES][260];      // NEW tree order (after renumbering by first use)
  uint32_t code[LBZ_MAX_TREES][260];
  uint8_t selector[18008];                 // NEW tree numbers, one per 50-symbol group
  uint8_t selector_mtf[18008];             // incl. the optional padding selector
};

// Device-side timers (CUDA events on the engine's stream).
#define LBZ_NSTAGE 8     // rle1, sort8 (initial radix), refine (doubling rounds), bwt_final, mtf, huffman, pack, copy
#define LBZ_NK0 8        // launches of the dominant kernel (k_scatter<0>) per batch
struct LbzTimers {
  cudaEvent_t stage[LBZ_NSTAGE + 1];
  cudaEvent_t k0[2 * LBZ_NK0];
  int enabled;
};

#define LBZ_CUDA_CHECK(x)                                                        \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "lbzip2_b200: CUDA error %s at %s:%d: %s\n",               \
              cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

// ---- small device helpers -------------------------------------------------

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming 128-bit load that does not allocate in L1.
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// L2 residency hints.  Randomly scattered 4-byte updates (the rank array) should
// stay in L2 until their sector is complete; data that is streamed through once
// should not push them out.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_u32_hint(uint32_t *p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_u32_hint(const uint32_t *p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// Inclusive warp scan (sum).
__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (uint32_t)o) v += t;
  }
  return v;
}
__device__ __forceinline__ int warp_incl_max(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (uint32_t)o) v = max(v, t);
  }
  return v;
}

// CTA-wide exclusive sum over one value per thread.  `ws` = 33+ words of
// shared scratch.  Returns the exclusive prefix; *total gets the CTA sum.
// Contains two __syncthreads(); safe to call repeatedly with the same scratch.
__device__ __forceinline__ uint32_t cta_excl_sum(uint32_t v, uint32_t *ws, uint32_t *total) {
  const uint32_t w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint32_t inc = warp_incl_sum(v);
  __syncthreads();                       // protect ws from a previous call
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane_id() < nw ? ws[lane_id()] : 0;
    uint32_t xi = warp_incl_sum(x);
    ws[lane_id()] = xi - x;              // exclusive warp bases
    if (lane_id() == 31) ws[32] = xi;
  }
  __syncthreads();
  *total = ws[32];
  return ws[w] + inc - v;
}

// CTA-wide exclusive max over one int per thread (identity INT_MIN given by caller as `ident`).
__device__ __forceinline__ int cta_excl_max(int v, int ident, int *ws, int *total) {
  const uint32_t w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = warp_incl_max(v);
  int exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane_id() == 0) exc = ident;
  __syncthreads();
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane_id() < nw ? ws[lane_id()] : ident;
    int xi = warp_incl_max(x);
    int xe = __shfl_up_sync(0xffffffffu, xi, 1);
    if (lane_id() == 0) xe = ident;
    ws[lane_id()] = xe;
    if (lane_id() == 31) ws[32] = xi;
  }
  __syncthreads();
  *total = ws[32];
  return max(ws[w], exc);
}
