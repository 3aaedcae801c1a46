// unbz_kernels.cuh -- device code of the bzip2 block DECOMPRESSOR.
//
// Replaces, for a whole batch of blocks at once, the reference's
//   scan()      src/parse.c:281-342   -> k_ub_scan        (48-bit block magic at any bit offset)
//   retrieve()  src/decode.c:518-791  -> k_ub_header, k_ub_tree_*, k_ub_chain, k_ub_symbols (prefix decoding)
//                                        k_ub_tok_sum/scan/write (zero-run arithmetic, status)
//                                        k_ub_mtf_tile/scan/fill (inverse MTF, run expansion)
//   decode()    src/decode.c:840-917  -> k_ub_lf_*, k_ub_walk1/rank/walk2, k_ub_period, k_ub_derand
//   emit()      src/decode.c:936-1143 -> k_ub_rl_sum/scan/emit, k_ub_crc_fin
//
// Shape of the work.  Where a prefix code starts depends on every code before it, so ONE thread per
// block (k_ub_chain) walks the code lengths -- up to four codes per table look-up -- and does nothing
// else; throughput comes from many blocks in flight.  Everything else is re-stated so that a block
// is worked on by thousands of threads:
//   * symbol values are decoded per 50-code group from the positions the walk left (symbols);
//   * the zero-run digits become prefix sums over symbol tiles, which also reproduces the
//     reference's overflow tests and their precedence over later errors (tok_sum/scan/write);
//   * the inverse MTF is a product of list permutations: tiles of 1024 tokens compute their own
//     permutation (mtf_tile), one thread per block chains them into the list at every tile start
//     (mtf_scan), tiles then replay their tokens and fill the runs of the last column (mtf_fill);
//   * the successor table of the inverse BWT is a stable counting sort by byte value
//     (tile histograms -> per-value scan over tiles -> tile scatter);
//   * the 900 k-step pointer chase is cut at ~3500 splitter nodes: every splitter walks to the next
//     one (walk1), one thread ranks the splitters (rank), every splitter walks again and writes its
//     piece of the text at its final position (walk2)  [Helman-JaJa list ranking];
//   * undoing the initial run-length coding is a 5-state automaton; tiles publish their
//     state-to-state maps and output sizes (rl_sum), one thread chains them (rl_scan), tiles then
//     write their output and fold their CRC, shifted to the end of the block in GF(2), into one
//     word per block (rl_emit, crc_fin).
//
// Every kernel here is written in a restricted style: no __syncthreads, no warp intrinsics, shared
// memory only where a single thread of the CTA uses it.  That keeps the arithmetic identical when
// the same source is compiled for the host by tests/simt_emul (UB_EMUL), which is how the logic is
// checked on machines without a GPU.  The emulation is test infrastructure; the product library is
// built from this file by nvcc only and has no host path.
#ifndef UNBZ_KERNELS_CUH
#define UNBZ_KERNELS_CUH

#include <stdint.h>
#include <stddef.h>

#ifdef UB_EMUL
#define UB_KERNEL static void
#define UB_DEVICE static inline
#define UB_SHARED static
struct UbEmuIdx { unsigned gid, bid, tid; };
extern UbEmuIdx ub_emu_idx;
#define UB_GID ((uint64_t)ub_emu_idx.gid)
#define UB_BID (ub_emu_idx.bid)
#define UB_TID (ub_emu_idx.tid)
static inline uint32_t ub_atomic_xor(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o ^ v; return o; }
static inline uint32_t ub_atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
struct uint2 { uint32_t x, y; };
static inline uint32_t ub_bswap32(uint32_t v) { return __builtin_bswap32(v); }
static inline uint32_t ub_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31u)) >> 32); }
#else
#define UB_KERNEL __global__ void
#define UB_DEVICE __device__ __forceinline__
#define UB_SHARED __shared__
#define UB_GID ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x)
#define UB_BID (blockIdx.x)
#define UB_TID (threadIdx.x)
__device__ __forceinline__ uint32_t ub_atomic_xor(uint32_t *p, uint32_t v) { return atomicXor(p, v); }
__device__ __forceinline__ uint32_t ub_atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
__device__ __forceinline__ uint32_t ub_bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint32_t ub_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_l(lo, hi, sh); }
#endif

// ---- geometry ---------------------------------------------------------------------------------
#define UB_MAXBLK 900000u                 // MAX_BLOCK_SIZE, src/common.h:50
#define UB_STRIDE 900096u                 // per-slot stride of the byte / node arrays (multiple of 128)
#define UB_SYMSTRIDE 900096u              // symbols kept per block (18001 groups of 50 at most)
#define UB_SELCAP 18016u                  // selectors kept per block (18001 used, src/decode.c:631)
#define UB_TL 2048u                       // counting-sort tile (bytes)
#define UB_NTL ((UB_MAXBLK + UB_TL - 1) / UB_TL)          // 440
#define UB_SPL_SHIFT 8u                   // a splitter every 256 nodes ...
#define UB_KS ((UB_MAXBLK >> UB_SPL_SHIFT) + 2u)         // ... 3516 of them + 1 spare + the start node
#define UB_TU 1024u                       // run-expansion tile (bytes of coded text)
#define UB_NTU ((UB_MAXBLK + UB_TU - 1) / UB_TU)          // 879
#define UB_UNSET 0xFFFFFFFFu
#define UB_NOEMIT 0xFFFFFFFFFFFFFFFFull

// status values = the reference's `enum error` (src/common.h:54-76), same order
enum {
  UB_OK = 0, UB_MORE, UB_FINISH, UB_ERR_MAGIC, UB_ERR_HEADER, UB_ERR_BITMAP, UB_ERR_TREES,
  UB_ERR_GROUPS, UB_ERR_SELECTOR, UB_ERR_DELTA, UB_ERR_PREFIX, UB_ERR_INCOMPLT, UB_ERR_EMPTY,
  UB_ERR_UNTERM, UB_ERR_RUNLEN, UB_ERR_BLKCRC, UB_ERR_STRMCRC, UB_ERR_OVERFLOW, UB_ERR_BWTIDX,
  UB_ERR_EOF
};

// One block slot of a wave (plain data, shared by host and device).
struct UbBlock {
  uint64_t pos;          // bit position of the block magic in the input
  uint64_t end_bit;      // first bit after the end-of-block symbol
  uint64_t out_len;      // decoded bytes of the block
  uint64_t out_off;      // where they go in the wave's output buffer; UB_NOEMIT = not emitted
  uint32_t status;       // retrieve() status
  uint32_t rand, bwt_idx, block_size;
  uint32_t alpha_size, num_trees, num_selectors;
  uint32_t period;       // length of the successor cycle through the primary index
  uint32_t rl_state;     // run-expansion state after the last byte (4 = missing run length)
  uint32_t crc_acc;      // XOR of the tiles' shifted CRC contributions
  uint32_t crc;          // final block CRC
  uint32_t ntok;         // list moves (k_ub_tok_scan)
  uint32_t nsym;         // symbols walked by k_ub_chain (incl. the end-of-block symbol)
  uint32_t ngrp;         // 50-symbol groups k_ub_chain started
  uint64_t sym_bit;      // first bit of the symbol data (k_ub_header)
};

// ---- bit reader (big-endian 32-bit words, as src/decode.c:372-426) -----------------------------
struct UbBits {
  const uint32_t *words;
  uint64_t nwords;       // whole words of input, the tail zero-filled
  uint64_t wi;           // next word to append
  uint64_t v;            // live bits, left-justified
  uint32_t w;            // number of live bits
  uint32_t ahead;        // words[wi], loaded one refill early so that its latency is hidden
};

// Position the reader on absolute bit p.  Returns false if p lies beyond the input.
UB_DEVICE bool ub_bits_seek(UbBits &b, uint64_t p) {
  b.wi = p >> 5;
  b.v = 0; b.w = 0; b.ahead = 0;
  if (b.wi >= b.nwords) return false;
  uint32_t off = (uint32_t)(p & 31u);
  b.v = (uint64_t)ub_bswap32(b.words[b.wi]) << (32u + off);
  b.w = 32u - off;
  b.wi++;
  b.ahead = b.wi < b.nwords ? b.words[b.wi] : 0u;
  return true;
}
// NEED(): at least 32 live bits, or report the end of input (src/decode.c:387-407).
UB_DEVICE bool ub_bits_need(UbBits &b) {
  if (b.w < 32u) {
    if (b.wi >= b.nwords) return false;
    b.v |= (uint64_t)ub_bswap32(b.ahead) << (32u - b.w);
    b.w += 32u;
    b.wi++;
    b.ahead = b.wi < b.nwords ? b.words[b.wi] : 0u;
  }
  return true;
}
UB_DEVICE uint32_t ub_bits_peek(const UbBits &b, uint32_t k) { return (uint32_t)(b.v >> (64u - k)); }
UB_DEVICE void ub_bits_dump(UbBits &b, uint32_t k) { b.v <<= k; b.w -= k; }
UB_DEVICE uint32_t ub_bits_take(UbBits &b, uint32_t k) { uint32_t x = ub_bits_peek(b, k); ub_bits_dump(b, k); return x; }
UB_DEVICE uint64_t ub_bits_pos(const UbBits &b) { return (b.wi << 5) - b.w; }

// ---- k_ub_scan: block magics at every bit offset (src/parse.c:281-342) --------------------------
// One thread per input word: the 32 alignments that start inside it.
UB_KERNEL k_ub_scan(const uint32_t *words, uint64_t nwords, uint64_t *hits, uint32_t *nhits, uint32_t hit_cap) {
  uint64_t wi = UB_GID;
  if (wi >= nwords) return;
  uint64_t w0 = ub_bswap32(words[wi]);
  uint64_t w1 = wi + 1 < nwords ? ub_bswap32(words[wi + 1]) : 0;
  uint64_t w2 = wi + 2 < nwords ? ub_bswap32(words[wi + 2]) : 0;
  uint64_t hi = (w0 << 32) | w1, lo = w2 << 32;
  for (uint32_t s = 0; s < 32u; s++) {
    uint64_t x = s ? ((hi << s) | (lo >> (64u - s))) : hi;
    if ((x >> 16) == 0x314159265359ull) {
      uint64_t p = (wi << 5) + s;
      if (p + 48u <= (nwords << 5)) {
        uint32_t k = ub_atomic_add(nhits, 1u);
        if (k < hit_cap) hits[k] = p;
      }
    }
  }
}

// ---- prefix decoding ----------------------------------------------------------------------------
// retrieve() (src/decode.c:518-791) is split so that the one thing that is serial inside a block --
// finding where every code starts -- is all the serial thread does:
//   k_ub_header      1 thread/block   header fields, byte map, selectors, code lengths
//   k_ub_tree_canon  thread/tree      Kraft check, canonical code tables (make_tree, decode.c:181-305)
//   k_ub_tree_l1     thread/symbol    12-bit table: window -> (symbol, length)
//   k_ub_tree_multi  thread/window    12-bit table: window -> cumulative lengths of up to 4 whole codes
//   k_ub_chain       1 thread/block   walks the code LENGTHS, up to 4 codes per table step, and leaves
//                                     the bit position and tree of every 50-code group
//   k_ub_symbols     thread/group     decodes the group's symbol values
#define UB_WBITS 12u
#define UB_WSIZE (1u << UB_WBITS)
#define UB_PENDING 0xFFFFFFFFu            // header parsed, symbols not walked yet
#define UB_MAXGRP 18001u                  // src/decode.c:631

struct UbTreeG {                          // one prefix code (global memory)
  uint32_t first[22];                     // first code of each length (right-justified)
  uint32_t count[22];
  uint32_t offset[22];                    // index into perm of the first symbol of each length
  uint16_t perm[258];                     // symbols in canonical order
  uint16_t status;                        // tree number, or UB_ERR_PREFIX / UB_ERR_INCOMPLT
  uint8_t len[260];                       // code lengths as transmitted
};

// Move the byte at position r of a 256-entry list (packed little-endian in 64 words) to the front.
UB_DEVICE uint32_t ub_mtf_front(uint32_t *lw, uint32_t r) {
  uint32_t wr = r >> 2, sh = (r & 3u) * 8u;
  uint32_t c = (lw[wr] >> sh) & 0xFFu;
  uint32_t m = sh == 24u ? 0xFFFFFFFFu : ((1u << (sh + 8u)) - 1u);
  uint32_t cur = lw[wr];
  for (uint32_t k = wr; k > 0; k--) {
    uint32_t below = lw[k - 1];
    uint32_t shifted = (cur << 8) | (below >> 24);
    lw[k] = (k == wr) ? ((shifted & m) | (cur & ~m)) : shifted;
    cur = below;
  }
  {
    uint32_t shifted = (cur << 8) | c;
    lw[0] = (wr == 0) ? ((shifted & m) | (cur & ~m)) : shifted;
  }
  return c;
}

// One CTA of 32 threads per block slot; thread 0 does the work.  Leaves status = UB_PENDING and
// sym_bit = first bit of the symbol data, or the error retrieve() would report.
UB_KERNEL k_ub_header(const uint32_t *words, uint64_t nwords, UbBlock *blk, uint32_t nblk,
                      uint32_t *list0_all, uint8_t *sel_all, UbTreeG *tree_all) {
  if (UB_TID != 0) return;
  const uint32_t b = UB_BID;
  if (b >= nblk) return;
  UbBlock &B = blk[b];
  uint8_t *sel = sel_all + (size_t)b * UB_SELCAP;
  uint32_t *listw = list0_all + (size_t)b * 256u;
  uint32_t status = UB_PENDING;
  UbBits br;
  br.words = words; br.nwords = nwords;

  B.block_size = 0; B.end_bit = 0; B.rand = 0; B.bwt_idx = 0; B.alpha_size = 0;
  B.num_trees = 0; B.num_selectors = 0; B.period = 0; B.rl_state = 0; B.crc_acc = 0; B.crc = 0;
  B.out_len = 0; B.out_off = UB_NOEMIT; B.ntok = 0; B.nsym = 0; B.ngrp = 0; B.sym_bit = 0;

#define UB_FAIL(code) do { status = (code); goto finish; } while (0)
#define UB_NEED() do { if (!ub_bits_need(br)) UB_FAIL(UB_ERR_EOF); } while (0)
  {
    uint32_t alpha, ntrees, nsel, nsym = 0;
    if (!ub_bits_seek(br, B.pos + 80u)) UB_FAIL(UB_ERR_EOF);
    UB_NEED();
    B.rand = ub_bits_take(br, 1);
    B.bwt_idx = ub_bits_take(br, 24);

    // byte map: 16 row flags, 16 bits per used row (src/decode.c:533-553)
    UB_NEED();
    {
      uint32_t big = ub_bits_take(br, 16);
      uint32_t acc = 0;
      for (uint32_t i = 0; i < 64; i++) listw[i] = 0;
      for (uint32_t i = 0; i < 16; i++) {
        if (big & (0x8000u >> i)) {
          uint32_t small = ub_bits_take(br, 16);
          UB_NEED();
          for (uint32_t j = 0; j < 16; j++)
            if (small & (0x8000u >> j)) {
              acc |= (16u * i + j) << ((nsym & 3u) * 8u);
              nsym++;
              if ((nsym & 3u) == 0) { listw[(nsym >> 2) - 1u] = acc; acc = 0; }
            }
        }
      }
      if (nsym & 3u) listw[nsym >> 2] = acc;
    }
    if (nsym == 0) UB_FAIL(UB_ERR_BITMAP);
    alpha = nsym + 2u;
    B.alpha_size = alpha;

    ntrees = ub_bits_take(br, 3);
    B.num_trees = ntrees;
    if (ntrees < 2u || ntrees > 6u) UB_FAIL(UB_ERR_TREES);
    nsel = ub_bits_take(br, 15);
    B.num_selectors = nsel;
    if (nsel == 0) UB_FAIL(UB_ERR_GROUPS);

    // unary selector ranks, 6-bit look-ahead (src/decode.c:566-575)
    for (uint32_t i = 0; i < nsel; i++) {
      uint32_t x = ub_bits_peek(br, 6);
      uint32_t k = 1;
      while (k <= 6u && (x & (0x40u >> k))) k++;
      if (k > ntrees) UB_FAIL(UB_ERR_SELECTOR);
      if (i < UB_SELCAP) sel[i] = (uint8_t)(k - 1u);
      ub_bits_dump(br, k);
      UB_NEED();
    }

    // delta-coded lengths: up to three +-1 steps per 6-bit window, range checked once per
    // window (src/decode.c:577-601 with its L[]/R[] tables)
    for (uint32_t t = 0; t < ntrees; t++) {
      uint8_t *len = tree_all[(size_t)b * 6u + t].len;
      int cur = (int)ub_bits_take(br, 5);
      uint32_t j = 0;
      while (j < alpha) {
        uint32_t x = ub_bits_peek(br, 6);
        uint32_t used = 0;
        bool done = false;
        while (used + 2u <= 6u && (x & (0x20u >> used))) {
          cur += (x & (0x20u >> (used + 1u))) ? -1 : 1;
          used += 2u;
        }
        if (used < 6u) { used += 1u; done = true; }
        if (cur < 1 || cur > 20) UB_FAIL(UB_ERR_DELTA);
        if (done) len[j++] = (uint8_t)cur;
        ub_bits_dump(br, used);
        UB_NEED();
      }
    }
  }
finish:
#undef UB_FAIL
#undef UB_NEED
  B.end_bit = ub_bits_pos(br);
  B.sym_bit = B.end_bit;
  B.status = status;
}

// thread per (slot, tree): make_tree(), src/decode.c:181-305.  A bad tree only matters if a group
// selects it, so its kind is kept as the tree's status.
UB_KERNEL k_ub_tree_canon(const UbBlock *blk, uint32_t nblk, UbTreeG *tree_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / 6u), t = (uint32_t)(g % 6u);
  if (b >= nblk || blk[b].status != UB_PENDING || t >= blk[b].num_trees) return;
  UbTreeG &T = tree_all[(size_t)b * 6u + t];
  const uint32_t alpha = blk[b].alpha_size;
  uint32_t count[22];
  for (uint32_t k = 0; k < 22; k++) count[k] = 0;
  for (uint32_t s = 0; s < alpha; s++) count[T.len[s]]++;
  uint64_t kraft = 0;
  for (uint32_t k = 1; k <= 20; k++) kraft += (uint64_t)count[k] << (20u - k);
  if (kraft != (1u << 20)) {
    T.status = (uint16_t)(kraft < (1u << 20) ? UB_ERR_INCOMPLT : UB_ERR_PREFIX);
    return;
  }
  T.status = (uint16_t)t;
  uint32_t code = 0, off = 0;
  uint32_t fill[22];
  for (uint32_t k = 1; k <= 20; k++) {
    T.first[k] = code; T.offset[k] = off; T.count[k] = count[k]; fill[k] = off;
    code = (code + count[k]) << 1;
    off += count[k];
  }
  T.first[0] = 0; T.count[0] = 0; T.offset[0] = 0; T.first[21] = 0; T.count[21] = 0; T.offset[21] = 0;
  for (uint32_t s = 0; s < alpha; s++) T.perm[fill[T.len[s]]++] = (uint16_t)s;
}

// thread per (slot, tree, symbol): the symbol's code, left-justified in 12 bits, owns a span of the
// window table.  Entries no short code reaches stay 0 (the table is cleared before every wave).
UB_KERNEL k_ub_tree_l1(const UbBlock *blk, uint32_t nblk, const UbTreeG *tree_all, uint16_t *l1_all) {
  uint64_t g = UB_GID;
  uint32_t s = (uint32_t)(g % 258u), t = (uint32_t)((g / 258u) % 6u), b = (uint32_t)(g / (258u * 6u));
  if (b >= nblk || blk[b].status != UB_PENDING || t >= blk[b].num_trees || s >= blk[b].alpha_size) return;
  const UbTreeG &T = tree_all[(size_t)b * 6u + t];
  if (T.status >= 6u) return;
  uint32_t len = T.len[s];
  if (len > UB_WBITS) return;
  uint32_t q = 0;                                   // rank of s among the symbols of its length
  for (uint32_t x = 0; x < s; x++) q += (T.len[x] == len);
  uint32_t base = (T.first[len] + q) << (UB_WBITS - len), span = 1u << (UB_WBITS - len);
  uint16_t e = (uint16_t)((s << 5) | len);
  uint16_t *l1 = l1_all + ((size_t)b * 6u + t) * UB_WSIZE;
  for (uint32_t z = 0; z < span; z++) l1[base + z] = e;
}

// thread per (slot, tree, window): how far up to four whole codes reach into the window.
//   bits  0..3   length of all of them together      } all the walk needs in the common case
//   bits  4..6   number of codes (0: the first code is longer than the window)
//   bit   7      set if the entry needs care: no code, or the end-of-block symbol is among them
//   bits  8..23  cumulative length after the 1st..4th code (4 bits each)
//   bits 24..26  1-based index of the end-of-block symbol among them, 0 if absent
// The low byte of every entry is also kept in a byte table of its own (mq): it is all the common case
// of the walk reads, and at 4 KB per tree the tables of several blocks stay in one SM's L1.
UB_KERNEL k_ub_tree_multi(const UbBlock *blk, uint32_t nblk, const UbTreeG *tree_all, const uint16_t *l1_all,
                          uint32_t *ml_all, uint8_t *mq_all) {
  uint64_t g = UB_GID;
  uint32_t idx = (uint32_t)(g % UB_WSIZE), t = (uint32_t)((g / UB_WSIZE) % 6u), b = (uint32_t)(g / (UB_WSIZE * 6u));
  if (b >= nblk || blk[b].status != UB_PENDING || t >= blk[b].num_trees) return;
  if (tree_all[(size_t)b * 6u + t].status >= 6u) return;
  const uint16_t *l1 = l1_all + ((size_t)b * 6u + t) * UB_WSIZE;
  const uint32_t eob = blk[b].alpha_size - 1u;
  uint32_t cum = 0, cnt = 0, lens = 0, eobk = 0, cur = idx;
  for (uint32_t c = 0; c < 4u; c++) {
    uint32_t x = l1[cur];
    uint32_t len = x & 31u;
    if (x == 0 || cum + len > UB_WBITS) break;   // the next code is not wholly inside the known bits
    cum += len;
    lens |= cum << (4u * c);
    cnt++;
    if ((x >> 5) == eob) { eobk = cnt; break; }
    cur = (idx << cum) & (UB_WSIZE - 1u);
  }
  uint32_t care = (cnt == 0 || eobk != 0) ? 0x80u : 0u;
  ml_all[((size_t)b * 6u + t) * UB_WSIZE + idx] = cum | (cnt << 4) | care | (lens << 8) | (eobk << 24);
  mq_all[((size_t)b * 6u + t) * UB_WSIZE + idx] = (uint8_t)(cum | (cnt << 4) | care);
}

// One code by its canonical tables, for windows the 12-bit tables do not resolve.  c20 = the next
// 20 bits.  Returns the symbol, *k = its length.
UB_DEVICE uint32_t ub_canon_decode(const UbTreeG &T, uint32_t c20, uint32_t *k) {
  for (uint32_t len = 1; len <= 20u; len++) {
    uint32_t c = c20 >> (20u - len);
    if (c - T.first[len] < T.count[len]) { *k = len; return T.perm[T.offset[len] + c - T.first[len]]; }
  }
  *k = 20u;                                          // unreachable for a complete code
  return 0;
}

// One CTA of 32 threads per block slot; thread 0 walks the code lengths.  Leaves the SERIAL status
// (what stopped the walk: the end-of-block symbol, the end of the input, a bad tree, no end-of-block
// in the last group); k_ub_tok_scan turns it into retrieve()'s status.
UB_KERNEL k_ub_chain(const uint32_t *words, uint64_t nwords, UbBlock *blk, uint32_t nblk, const uint8_t *sel_all,
                     const UbTreeG *tree_all, const uint16_t *l1_all, const uint32_t *ml_all, const uint8_t *mq_all,
                     uint64_t *gpos_all, uint8_t *gtree_all) {
  if (UB_TID != 0) return;
  const uint32_t b = UB_BID;
  if (b >= nblk) return;
  UbBlock &B = blk[b];
  if (B.status != UB_PENDING) return;
  const uint8_t *sel = sel_all + (size_t)b * UB_SELCAP;
  uint64_t *gpos = gpos_all + (size_t)b * (UB_MAXGRP + 1u);
  uint8_t *gtree = gtree_all + (size_t)b * (UB_MAXGRP + 1u);
  const uint32_t eob = B.alpha_size - 1u;
  uint32_t nsel = B.num_selectors > UB_MAXGRP ? UB_MAXGRP : B.num_selectors;   // src/decode.c:631-632
  uint32_t sl[6];
  for (uint32_t t = 0; t < 6u; t++) sl[t] = t < B.num_trees ? tree_all[(size_t)b * 6u + t].status : 0u;
  uint64_t pos = B.sym_bit;
  uint32_t status = UB_ERR_UNTERM, nsym = 0, g = 0;
  UbBits br;
  br.words = words; br.nwords = nwords;

  for (; g < nsel; g++) {
    uint32_t r = sel[g];
    uint32_t t = sl[r];
    if (t >= 6u) { status = t; break; }              // a bad tree is selected (src/decode.c:640-642)
    for (; r > 0; r--) sl[r] = sl[r - 1];
    sl[0] = t;
    gpos[g] = pos;
    gtree[g] = (uint8_t)t;
    const UbTreeG &T = tree_all[(size_t)b * 6u + t];
    const uint32_t *ml = ml_all + ((size_t)b * 6u + t) * UB_WSIZE;
    const uint8_t *mq = mq_all + ((size_t)b * 6u + t) * UB_WSIZE;
    bool done = false;

    // Two code paths, as in the reference (src/decode.c:644-661): while a whole group (50 codes of
    // at most 20 bits = 32 words) is certain to lie inside the input, a lean window reader without
    // end-of-input tests is used; the last groups go through the exact reader.
    if ((pos >> 5) + 36u <= nwords) {
      const uint32_t *wp = words + (pos >> 5);         // hi = wp[0], lo = wp[1], ahead = wp[2]
      uint32_t bp = (uint32_t)(pos & 31u);
      uint32_t hi = ub_bswap32(wp[0]), lo = ub_bswap32(wp[1]), ahead = wp[2];
      uint32_t rem = 50u;
      while (rem) {
        uint32_t win = ub_funnel_l(lo, hi, bp);        // the 32 bits that start at bit bp of hi
        uint32_t q = mq[win >> (32u - UB_WBITS)];
        uint32_t len;
        if (rem >= 4u && !(q & 0x80u)) {               // common case: all codes of the entry, no end of block
          len = q & 15u;
          rem -= q >> 4;
        } else {
          uint32_t e = ml[win >> (32u - UB_WBITS)];
          uint32_t cnt = (e >> 4) & 7u;
          if (cnt) {
            uint32_t take = cnt < rem ? cnt : rem;
            uint32_t eobk = e >> 24;
            if (eobk && eobk <= take) { take = eobk; done = true; }
            len = (e >> (4u + 4u * take)) & 15u;
            rem -= take;
          } else {
            uint32_t s = ub_canon_decode(T, win >> 12, &len);
            rem -= 1u;
            if (s == eob) done = true;
          }
        }
        bp += len;
        if (bp >= 32u) {
          bp -= 32u; wp++;
          hi = lo; lo = ub_bswap32(ahead); ahead = wp[2];
        }
        if (done) break;
      }
      nsym += 50u - rem;
      pos = ((uint64_t)(wp - words) << 5) + bp;
    } else {
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
  B.nsym = nsym;
  B.ngrp = g;
  B.end_bit = pos;
  B.status = status;
}

// thread per (slot, group): the group's symbol values (0 = RUNA, 1 = RUNB, 2..alpha-2 = list rank
// + 1, alpha-1 = end of block), decoded from the position k_ub_chain left
UB_KERNEL k_ub_symbols(const uint32_t *words, uint64_t nwords, const UbBlock *blk, uint32_t nblk,
                       const UbTreeG *tree_all, const uint16_t *l1_all, const uint64_t *gpos_all,
                       const uint8_t *gtree_all, uint16_t *sym_all) {
  uint64_t gid = UB_GID;
  uint32_t b = (uint32_t)(gid / (UB_MAXGRP + 1u)), g = (uint32_t)(gid % (UB_MAXGRP + 1u));
  if (b >= nblk || g >= blk[b].ngrp) return;
  uint32_t nsym = blk[b].nsym, lo_s = g * 50u;
  if (lo_s >= nsym) return;
  uint32_t cnt = nsym - lo_s < 50u ? nsym - lo_s : 50u;
  uint32_t t = gtree_all[(size_t)b * (UB_MAXGRP + 1u) + g];
  const UbTreeG &T = tree_all[(size_t)b * 6u + t];
  const uint16_t *l1 = l1_all + ((size_t)b * 6u + t) * UB_WSIZE;
  uint16_t *sym = sym_all + (size_t)b * UB_SYMSTRIDE + lo_s;
  uint64_t pos = gpos_all[(size_t)b * (UB_MAXGRP + 1u) + g];
  uint64_t wi = pos >> 5;
  uint32_t bp = (uint32_t)(pos & 31u);
  uint32_t hi = wi < nwords ? ub_bswap32(words[wi]) : 0u;
  uint32_t lo = wi + 1 < nwords ? ub_bswap32(words[wi + 1]) : 0u;
  for (uint32_t j = 0; j < cnt; j++) {
    uint32_t win = ub_funnel_l(lo, hi, bp);
    uint32_t x = l1[win >> (32u - UB_WBITS)];
    uint32_t s, k;
    if (x) { s = x >> 5; k = x & 31u; } else s = ub_canon_decode(T, win >> 12, &k);
    sym[j] = (uint16_t)s;
    bp += k;
    if (bp >= 32u) {
      bp -= 32u; wi++;
      hi = lo;
      lo = wi + 1 < nwords ? ub_bswap32(words[wi + 1]) : 0u;
    }
  }
}

// ---- zero-run arithmetic (src/decode.c:756-775) as prefix sums over the symbols -----------------
// Every symbol adds to the length of the last column: a list-move symbol 1 (the first byte of its
// run), the k-th digit d of a zero run (d+1) << k.  P(i) = sum of what symbols 0..i-1 add.  The
// reference's overflow tests become: a digit is refused when the run it extends already exceeds
// 900000; a list-move / end-of-block symbol at index i overflows when P(i) > 900000.  The run a
// digit extends is found by looking back over at most 21 symbols, so tiles need no carry.
#define UB_TS 1024u                                        // symbols per tile
#define UB_NTS ((UB_SYMSTRIDE + UB_TS - 1) / UB_TS)       // 879
#define UB_RUN_INF 0x7FFFFFFFu

struct UbRunState { uint32_t run, k; };                    // pending run and number of its digits

// State just before symbol i (what the reference's `run` / `shift` hold there).
UB_DEVICE UbRunState ub_run_state_at(const uint16_t *sym, uint32_t i) {
  UbRunState st;
  uint32_t k = 0;
  while (k < i && k < 22u && sym[i - 1u - k] < 2u) k++;
  if (k >= 22u) { st.run = UB_RUN_INF; st.k = 22u; return st; }
  uint32_t run = (i - k) > 0 ? 1u : 0u;               // a list-move symbol precedes the digits, or the block starts
  for (uint32_t j = 0; j < k; j++) {
    if (run > UB_MAXBLK) { run = UB_RUN_INF; break; }  // the digit would have been refused
    run += ((uint32_t)sym[i - k + j] + 1u) << j;
  }
  st.run = run; st.k = k;
  return st;
}
// Advance over symbol s; returns what it adds to the column length.
UB_DEVICE uint32_t ub_run_step(UbRunState &st, uint32_t s) {
  if (s < 2u) {
    if (st.run > UB_MAXBLK) { st.run = UB_RUN_INF; return 0; }
    uint32_t add = (s + 1u) << st.k;
    st.run += add; st.k++;
    return add;
  }
  st.run = 1u; st.k = 0;
  return 1u;
}

struct UbTokTile { uint32_t add, moves; };                 // column bytes added / list moves in the tile

// thread per (slot, tile)
UB_KERNEL k_ub_tok_sum(const UbBlock *blk, uint32_t nblk, const uint16_t *sym_all, UbTokTile *tt) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTS), tile = (uint32_t)(g % UB_NTS);
  if (b >= nblk) return;
  uint32_t nsym = blk[b].nsym, alpha = blk[b].alpha_size, lo = tile * UB_TS;
  if (lo >= nsym) return;
  uint32_t hi = lo + UB_TS < nsym ? lo + UB_TS : nsym;
  const uint16_t *sym = sym_all + (size_t)b * UB_SYMSTRIDE;
  UbRunState st = ub_run_state_at(sym, lo);
  uint32_t add = 0, moves = 0;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t s = sym[i];
    if (s == alpha - 1u) break;                        // end of block: adds nothing
    uint32_t a = ub_run_step(st, s);
    add = add + a < add ? 0xFFFFFFFFu : add + a;       // saturate
    moves += s >= 2u;
  }
  tt[(size_t)b * UB_NTS + tile].add = add;
  tt[(size_t)b * UB_NTS + tile].moves = moves;
}

// thread per slot: tile offsets, the first overflow (if any), retrieve()'s final status
UB_KERNEL k_ub_tok_scan(UbBlock *blk, uint32_t nblk, const uint16_t *sym_all, const UbTokTile *tt,
                        uint32_t *tile_pos, uint32_t *tile_tok) {
  uint64_t b = UB_GID;
  if (b >= nblk) return;
  UbBlock &B = blk[b];
  uint32_t nsym = B.nsym, alpha = B.alpha_size;
  uint32_t ntile = (nsym + UB_TS - 1u) / UB_TS;
  const uint16_t *sym = sym_all + (size_t)b * UB_SYMSTRIDE;
  uint64_t P = 0;
  uint32_t M = 0;
  bool overflow = false;
  for (uint32_t t = 0; t < ntile && !overflow; t++) {
    tile_pos[(size_t)b * UB_NTS + t] = (uint32_t)P;
    tile_tok[(size_t)b * UB_NTS + t] = M;
    uint64_t after = P + tt[(size_t)b * UB_NTS + t].add;
    if (after > UB_MAXBLK) {
      // P(i) passes 900000 inside this tile: replay from the tile start until the reference
      // would have raised the overflow, or the symbols end
      UbRunState st = ub_run_state_at(sym, t * UB_TS);
      uint64_t Pi = P;
      for (uint32_t i = t * UB_TS; i < nsym; i++) {
        uint32_t s = sym[i];
        if (s < 2u) {
          if (st.run > UB_MAXBLK) { overflow = true; break; }
        } else if (Pi > UB_MAXBLK) { overflow = true; break; }
        if (s == alpha - 1u) break;
        Pi += ub_run_step(st, s);
      }
      if (!overflow) { P = UB_MAXBLK + 1u; break; }    // symbols ended first: the serial status stands
    }
    P = after;
    M += tt[(size_t)b * UB_NTS + t].moves;
  }
  uint32_t status = B.status;
  if (overflow) status = UB_ERR_OVERFLOW;
  else if (status == UB_OK) {
    if (P == 0) status = UB_ERR_EMPTY;
    else if (B.bwt_idx >= P) status = UB_ERR_BWTIDX;
  }
  uint32_t n = P > UB_MAXBLK ? 0u : (uint32_t)P;
  B.status = status;
  B.block_size = n;
  B.period = n;
  B.ntok = M;
}

// thread per (slot, tile): the tile's list moves as tokens (rank, start of the new byte's run)
UB_KERNEL k_ub_tok_write(const UbBlock *blk, uint32_t nblk, const uint16_t *sym_all, const uint32_t *tile_pos,
                         const uint32_t *tile_tok, uint8_t *rank_all, uint32_t *pos_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTS), tile = (uint32_t)(g % UB_NTS);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t nsym = blk[b].nsym, alpha = blk[b].alpha_size, lo = tile * UB_TS;
  if (lo >= nsym) return;
  uint32_t hi = lo + UB_TS < nsym ? lo + UB_TS : nsym;
  const uint16_t *sym = sym_all + (size_t)b * UB_SYMSTRIDE;
  uint8_t *tok_rank = rank_all + (size_t)b * UB_STRIDE;
  uint32_t *tok_pos = pos_all + (size_t)b * UB_STRIDE;
  UbRunState st = ub_run_state_at(sym, lo);
  uint32_t P = tile_pos[(size_t)b * UB_NTS + tile], m = tile_tok[(size_t)b * UB_NTS + tile];
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t s = sym[i];
    if (s == alpha - 1u) break;
    if (s >= 2u) { tok_rank[m] = (uint8_t)(s - 1u); tok_pos[m] = P; m++; }
    P += ub_run_step(st, s);
  }
}

// ---- inverse MTF + run expansion (src/decode.c:428-516 mtf_one, :766-775) -----------------------
#define UB_TM 1024u                                        // tokens per tile
#define UB_NTM ((UB_MAXBLK + UB_TM) / UB_TM)              // 879 (a block has at most 900001 tokens)
UB_DEVICE uint32_t ub_mtf_ntile(uint32_t ntok) { return ntok ? (ntok + UB_TM - 1u) / UB_TM : 1u; }

// Per slot and tile, 128 words of scratch: [0,64) the tile's move product as a position map
// ("what stood at position p before the tile stands at inv[p] after it", one byte per position),
// [64,128) the list at the tile's start (filled by k_ub_mtf_scan).
#define UB_TPW 128u

// thread per (slot, tile): product of the tile's moves
UB_KERNEL k_ub_mtf_tile(const UbBlock *blk, uint32_t nblk, const uint8_t *rank_all, uint32_t *tileperm) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTM), tile = (uint32_t)(g % UB_NTM);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t ntok = blk[b].ntok;
  if (tile >= ub_mtf_ntile(ntok)) return;
  uint32_t lo = tile * UB_TM, hi = lo + UB_TM < ntok ? lo + UB_TM : ntok;
  const uint8_t *rk = rank_all + (size_t)b * UB_STRIDE;
  uint32_t lw[64];
  for (uint32_t i = 0; i < 64; i++) lw[i] = (4u * i) | ((4u * i + 1u) << 8) | ((4u * i + 2u) << 16) | ((4u * i + 3u) << 24);
  for (uint32_t k = lo; k < hi; k++) ub_mtf_front(lw, rk[k]);
  // lw: position i now holds what stood at position lw[i]; store the inverse map
  uint8_t *inv = (uint8_t *)(tileperm + ((size_t)b * UB_NTM + tile) * UB_TPW);
  for (uint32_t i = 0; i < 256; i++) inv[(lw[i >> 2] >> ((i & 3u) * 8u)) & 0xFFu] = (uint8_t)i;
}

// thread per (slot, list element): follow the element through the tiles and write it into the
// start list of every tile at the position it has there
UB_KERNEL k_ub_mtf_scan(const UbBlock *blk, uint32_t nblk, const uint32_t *list0_all, uint32_t *tileperm) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g >> 8), e = (uint32_t)(g & 255u);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t ntile = ub_mtf_ntile(blk[b].ntok);
  uint8_t val = (uint8_t)((list0_all[(size_t)b * 256u + (e >> 2)] >> ((e & 3u) * 8u)) & 0xFFu);
  uint32_t p = e;
  for (uint32_t t = 0; t < ntile; t++) {
    uint8_t *base = (uint8_t *)(tileperm + ((size_t)b * UB_NTM + t) * UB_TPW);
    base[256u + p] = val;
    p = base[p];
  }
}

// thread per (slot, tile): replay the tile's moves from its start list and fill the runs
UB_KERNEL k_ub_mtf_fill(const UbBlock *blk, uint32_t nblk, const uint8_t *rank_all, const uint32_t *pos_all,
                        const uint32_t *tileperm, uint8_t *bwt_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTM), tile = (uint32_t)(g % UB_NTM);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t ntok = blk[b].ntok, n = blk[b].block_size;
  if (tile >= ub_mtf_ntile(ntok)) return;
  uint32_t lo = tile * UB_TM, hi = lo + UB_TM < ntok ? lo + UB_TM : ntok;
  const uint8_t *rk = rank_all + (size_t)b * UB_STRIDE;
  const uint32_t *ps = pos_all + (size_t)b * UB_STRIDE;
  uint8_t *bwt = bwt_all + (size_t)b * UB_STRIDE;
  const uint32_t *start = tileperm + ((size_t)b * UB_NTM + tile) * UB_TPW + 64u;
  uint32_t lw[64];
  for (uint32_t i = 0; i < 64; i++) lw[i] = start[i];
  if (tile == 0) {
    uint8_t c0 = (uint8_t)(lw[0] & 0xFFu);
    uint32_t e = ntok ? ps[0] : n;
    for (uint32_t j = 0; j < e; j++) bwt[j] = c0;
  }
  for (uint32_t k = lo; k < hi; k++) {
    uint8_t c = (uint8_t)ub_mtf_front(lw, rk[k]);
    uint32_t s0 = ps[k], e = (k + 1u < ntok) ? ps[k + 1u] : n;
    for (uint32_t j = s0; j < e; j++) bwt[j] = c;
  }
}

// ---- successor table of the inverse BWT (src/decode.c:851-870) ---------------------------------
// node[j] = (i << 8) | c  where i is the position of the q-th occurrence of byte c in the last
// column and j = (#bytes < c) + q.  Walking x -> node[x] >> 8 from the primary index and reading
// the low bytes spells the text forward.

// thread per (slot, tile): byte histogram of the tile
UB_KERNEL k_ub_lf_hist(const UbBlock *blk, uint32_t nblk, const uint8_t *bwt_all, uint32_t *tilehist) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTL), tile = (uint32_t)(g % UB_NTL);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, lo = tile * UB_TL;
  if (lo >= n) return;
  uint32_t hi = lo + UB_TL < n ? lo + UB_TL : n;
  const uint8_t *bwt = bwt_all + (size_t)b * UB_STRIDE;
  uint32_t h[256];
  for (uint32_t i = 0; i < 256; i++) h[i] = 0;
  for (uint32_t i = lo; i < hi; i++) h[bwt[i]]++;
  uint32_t *row = tilehist + ((size_t)b * UB_NTL + tile) * 256u;
  for (uint32_t i = 0; i < 256; i++) row[i] = h[i];
}

// thread per (slot, byte value): occurrences of the value in the block
UB_KERNEL k_ub_lf_total(const UbBlock *blk, uint32_t nblk, const uint32_t *tilehist, uint32_t *ftab_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g >> 8), v = (uint32_t)(g & 255u);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, ntile = (n + UB_TL - 1u) / UB_TL;
  const uint32_t *col = tilehist + (size_t)b * UB_NTL * 256u + v;
  uint32_t tot = 0;
  for (uint32_t t = 0; t < ntile; t++) tot += col[(size_t)t * 256u];
  ftab_all[(size_t)b * 256u + v] = tot;
}

// thread per (slot, byte value): running start of this value's bucket through the tiles
UB_KERNEL k_ub_lf_scan(const UbBlock *blk, uint32_t nblk, const uint32_t *ftab_all, uint32_t *tilehist) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g >> 8), v = (uint32_t)(g & 255u);
  if (b >= nblk || blk[b].status != UB_OK) return;
  const uint32_t *f = ftab_all + (size_t)b * 256u;
  uint32_t run = 0;
  for (uint32_t i = 0; i < v; i++) run += f[i];
  uint32_t n = blk[b].block_size, ntile = (n + UB_TL - 1u) / UB_TL;
  uint32_t *col = tilehist + (size_t)b * UB_NTL * 256u + v;
  for (uint32_t t = 0; t < ntile; t++) {
    uint32_t c = col[(size_t)t * 256u];
    col[(size_t)t * 256u] = run;
    run += c;
  }
}

// thread per (slot, tile): stable scatter
UB_KERNEL k_ub_lf_scatter(const UbBlock *blk, uint32_t nblk, const uint8_t *bwt_all, const uint32_t *tilehist,
                          uint32_t *node_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTL), tile = (uint32_t)(g % UB_NTL);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, lo = tile * UB_TL;
  if (lo >= n) return;
  uint32_t hi = lo + UB_TL < n ? lo + UB_TL : n;
  const uint8_t *bwt = bwt_all + (size_t)b * UB_STRIDE;
  uint32_t *node = node_all + (size_t)b * UB_STRIDE;
  const uint32_t *row = tilehist + ((size_t)b * UB_NTL + tile) * 256u;
  uint32_t cnt[256];
  for (uint32_t i = 0; i < 256; i++) cnt[i] = row[i];
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t c = bwt[i];
    node[cnt[c]++] = (i << 8) | c;
  }
}

// ---- the chase, cut at splitters ----------------------------------------------------------------
// Splitter nodes: every node whose index is a multiple of 256, plus the primary index.  Slot k of a
// block's splitter table stands for node k*256; the last slot (UB_KS-1) stands for the primary
// index when that is not a multiple of 256.
UB_DEVICE bool ub_is_split(uint32_t x, uint32_t idx) { return (x & ((1u << UB_SPL_SHIFT) - 1u)) == 0 || x == idx; }
UB_DEVICE uint32_t ub_split_slot(uint32_t x, uint32_t idx) {
  return (x == idx && (idx & ((1u << UB_SPL_SHIFT) - 1u)) != 0) ? UB_KS - 1u : (x >> UB_SPL_SHIFT);
}
UB_DEVICE bool ub_slot_node(uint32_t k, uint32_t n, uint32_t idx, uint32_t *x) {
  if (k == UB_KS - 1u) {
    if ((idx & ((1u << UB_SPL_SHIFT) - 1u)) == 0) return false;
    *x = idx;
    return true;
  }
  uint32_t node = k << UB_SPL_SHIFT;
  if (node >= n) return false;
  *x = node;
  return true;
}

// thread per (slot, splitter): length of the piece that starts here and the splitter that ends it
UB_KERNEL k_ub_walk1(const UbBlock *blk, uint32_t nblk, const uint32_t *node_all, uint2 *seg) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_KS), k = (uint32_t)(g % UB_KS);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, idx = blk[b].bwt_idx, x;
  if (!ub_slot_node(k, n, idx, &x)) return;
  const uint32_t *node = node_all + (size_t)b * UB_STRIDE;
  uint32_t len = 0;
  do {
    x = node[x] >> 8;
    len++;
  } while (!ub_is_split(x, idx));
  uint2 v;
  v.x = ub_split_slot(x, idx);                         // the splitter that ends the piece
  v.y = len;
  seg[(size_t)b * UB_KS + k] = v;
}

// thread per slot: text position of every splitter on the cycle through the primary index (one
// 8-byte load per step).  If the cycle closes before block_size steps the text is periodic
// (src/decode.c:866-868) and `period` says how much of it the second walk produces.
UB_KERNEL k_ub_rank(UbBlock *blk, uint32_t nblk, const uint2 *seg, uint32_t *segpos) {
  uint64_t b = UB_GID;
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, idx = blk[b].bwt_idx;
  const uint32_t s0 = ub_split_slot(idx, idx);
  uint32_t s = s0, pos = 0;
  const size_t base = (size_t)b * UB_KS;
  do {
    uint2 v = seg[base + s];
    segpos[base + s] = pos;
    pos += v.y;
    s = v.x;
  } while (pos < n && s != s0);
  blk[b].period = pos < n ? pos : n;
}

// thread per (slot, splitter): write the piece
UB_KERNEL k_ub_walk2(const UbBlock *blk, uint32_t nblk, const uint32_t *node_all, const uint2 *seg,
                     const uint32_t *segpos, uint8_t *txt_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_KS), k = (uint32_t)(g % UB_KS);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t p = segpos[(size_t)b * UB_KS + k];
  if (p == UB_UNSET) return;
  uint32_t n = blk[b].block_size, idx = blk[b].bwt_idx, x;
  if (!ub_slot_node(k, n, idx, &x)) return;
  const uint32_t *node = node_all + (size_t)b * UB_STRIDE;
  uint8_t *txt = txt_all + (size_t)b * UB_STRIDE;
  uint32_t len = seg[(size_t)b * UB_KS + k].y;
  for (uint32_t t = 0; t < len && p + t < n; t++) {
    uint32_t v = node[x];
    txt[p + t] = (uint8_t)v;
    x = v >> 8;
  }
}

// thread per (slot, 1024 text positions): repeat the period
UB_KERNEL k_ub_period(const UbBlock *blk, uint32_t nblk, uint8_t *txt_all) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTU), tile = (uint32_t)(g % UB_NTU);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, p = blk[b].period;
  if (p >= n) return;
  uint32_t lo = tile * UB_TU, hi = lo + UB_TU < n ? lo + UB_TU : n;
  if (lo < p) lo = p;
  uint8_t *txt = txt_all + (size_t)b * UB_STRIDE;
  for (uint32_t i = lo; i < hi; i++) txt[i] = txt[i % p];
}

// thread per slot: undo the block randomisation of old bzip2 files (src/decode.c:893-899)
UB_KERNEL k_ub_derand(const UbBlock *blk, uint32_t nblk, uint8_t *txt_all, const uint16_t *rtab) {
  uint64_t b = UB_GID;
  if (b >= nblk || blk[b].status != UB_OK || !blk[b].rand) return;
  uint32_t n = blk[b].block_size;
  uint8_t *txt = txt_all + (size_t)b * UB_STRIDE;
  uint32_t k = 0, j = 617u;
  while (j < n) {
    txt[j] ^= 1u;
    k = (k + 1u) & 511u;
    j += rtab[k];
  }
}

// ---- run expansion + CRC (src/decode.c:936-1143) ------------------------------------------------
// State = number of equal literal bytes just seen (1..3), 4 = "the next byte is a repeat count",
// 0 = "the previous byte was a count" (also the start state).
struct UbTileSum {
  uint32_t out[5];       // output bytes of the tile for each entry state
  uint32_t end;          // exit state for each entry state, 3 bits each
};

UB_KERNEL k_ub_rl_sum(const UbBlock *blk, uint32_t nblk, const uint8_t *txt_all, UbTileSum *tsum) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTU), tile = (uint32_t)(g % UB_NTU);
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, lo = tile * UB_TU;
  if (lo >= n) return;
  uint32_t hi = lo + UB_TU < n ? lo + UB_TU : n;
  const uint8_t *txt = txt_all + (size_t)b * UB_STRIDE;
  uint32_t st[5], out[5];
  for (uint32_t s = 0; s < 5; s++) { st[s] = s; out[s] = 0; }
  uint32_t prev = lo ? txt[lo - 1] : 0x100u;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t c = txt[i];
    bool eq = (c == prev);
    for (uint32_t s = 0; s < 5; s++) {
      if (st[s] == 4u) { out[s] += c; st[s] = 0; }
      else { out[s] += 1u; st[s] = (st[s] >= 1u && eq) ? st[s] + 1u : 1u; }
    }
    prev = c;
  }
  UbTileSum &T = tsum[(size_t)b * UB_NTU + tile];
  uint32_t e = 0;
  for (uint32_t s = 0; s < 5; s++) { T.out[s] = out[s]; e |= st[s] << (3u * s); }
  T.end = e;
}

// thread per slot: entry state and output offset of every tile
UB_KERNEL k_ub_rl_scan(UbBlock *blk, uint32_t nblk, const UbTileSum *tsum, uint32_t *tstate, uint64_t *toff) {
  uint64_t b = UB_GID;
  if (b >= nblk || blk[b].status != UB_OK) return;
  uint32_t n = blk[b].block_size, ntile = (n + UB_TU - 1u) / UB_TU;
  uint32_t st = 0;
  uint64_t off = 0;
  for (uint32_t t = 0; t < ntile; t++) {
    const UbTileSum &T = tsum[(size_t)b * UB_NTU + t];
    tstate[(size_t)b * UB_NTU + t] = st;
    toff[(size_t)b * UB_NTU + t] = off;
    off += T.out[st];
    st = (T.end >> (3u * st)) & 7u;
  }
  blk[b].out_len = off;
  blk[b].rl_state = st;
}

// a * b in GF(2)[x] / (CRC-32/BZIP2 polynomial); bit k of a word = coefficient of x^k
UB_DEVICE uint32_t ub_gf_mul(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int i = 31; i >= 0; i--) {
    r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u);
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}
// a * x^(8*nbytes); pw[k] = x^(8 * 2^k)
UB_DEVICE uint32_t ub_gf_shift(uint32_t a, uint64_t nbytes, const uint32_t *pw) {
  for (uint32_t k = 0; nbytes; k++, nbytes >>= 1)
    if (nbytes & 1u) a = ub_gf_mul(a, pw[k]);
  return a;
}

// thread per (slot, tile): write the tile's output, fold its CRC into the block's accumulator
UB_KERNEL k_ub_rl_emit(UbBlock *blk, uint32_t nblk, const uint8_t *txt_all, const UbTileSum *tsum,
                       const uint32_t *tstate, const uint64_t *toff, uint8_t *out, const uint32_t *crctab,
                       const uint32_t *pw) {
  uint64_t g = UB_GID;
  uint32_t b = (uint32_t)(g / UB_NTU), tile = (uint32_t)(g % UB_NTU);
  if (b >= nblk || blk[b].status != UB_OK || blk[b].out_off == UB_NOEMIT) return;
  uint32_t n = blk[b].block_size, lo = tile * UB_TU;
  if (lo >= n) return;
  uint32_t hi = lo + UB_TU < n ? lo + UB_TU : n;
  const uint8_t *txt = txt_all + (size_t)b * UB_STRIDE;
  uint32_t st = tstate[(size_t)b * UB_NTU + tile];
  uint64_t off = toff[(size_t)b * UB_NTU + tile];
  uint8_t *o = out + blk[b].out_off + off;
  uint64_t w = 0;
  uint32_t crc = 0;
  uint32_t prev = lo ? txt[lo - 1] : 0x100u;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t c = txt[i];
    if (st == 4u) {
      for (uint32_t k = 0; k < c; k++) {
        o[w++] = (uint8_t)prev;
        crc = (crc << 8) ^ crctab[(crc >> 24) ^ prev];
      }
      st = 0;
    } else {
      o[w++] = (uint8_t)c;
      crc = (crc << 8) ^ crctab[(crc >> 24) ^ c];
      st = (st >= 1u && c == prev) ? st + 1u : 1u;
    }
    prev = c;
  }
  uint64_t after = blk[b].out_len - (off + w);
  ub_atomic_xor(&blk[b].crc_acc, ub_gf_shift(crc, after, pw));
}

// thread per slot: add the shifted initial register value, invert
UB_KERNEL k_ub_crc_fin(UbBlock *blk, uint32_t nblk, const uint32_t *pw) {
  uint64_t b = UB_GID;
  if (b >= nblk || blk[b].status != UB_OK || blk[b].out_off == UB_NOEMIT) return;
  blk[b].crc = ~(blk[b].crc_acc ^ ub_gf_shift(0xFFFFFFFFu, blk[b].out_len, pw));
}

#endif  // UNBZ_KERNELS_CUH
