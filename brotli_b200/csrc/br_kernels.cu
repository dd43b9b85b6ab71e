// br_kernels.cu -- CUDA kernels (sm_100a) and the per-stream pipeline that drives them.
//
// Stage map (reference function -> kernel):
//   hash keys / bucket rings (hash_longest_match64_inc.h Store*)  -> k_hash_keys, k_radix_* (the last pass writes rank[]), k_seg
//   CreateBackwardReferences + FindLongestMatch                   -> k_walk      (one warp per input block)
//   EncodeData glue (encode.c:985)                                -> k_chain     (one warp per stream)
//   BrotliBuildMetaBlockGreedy + BrotliStoreMetaBlock             -> k_encode_mb (one warp per metablock)
//   metablock concatenation / uncompressed fallback               -> k_assemble_scan, k_assemble_copy
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "br_params.h"
#include "br_lz77.h"
#include "br_chain.h"
#include "br_entropy.h"
#include "br_entropy2.h"
#include "br_assemble.h"
#include "br_entropy_flat.h"
#include "br_pipeline.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "brotli_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
  return 0; } } while (0)

// ---------------------------------------------------------------------------- hashing + sort
__global__ void k_hash_keys(BrParams P, const u8* __restrict__ data, u16* __restrict__ keys) {
  u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n) return;
  u32 hashable = P.n >= P.htl ? P.n - P.htl + 1 : 0;
  keys[p] = (u16)(p < hashable ? br_hash_key(P, data, p) : P.nbuckets);
}

// qualities 2..4: the key is the table slot of the one-position-per-slot hashers (br_lz77.h br_quick_slot), up to 20 bits
// (+ the overflow key of the unhashable tail): 32-bit keys, three radix passes
// One CTA per input block: the slot depends on the position counted from the first byte of the block's stream (BrBlk::base).
__global__ void k_slot_keys(BrParams P, const u8* __restrict__ data, const BrBlk* __restrict__ blk, u32* __restrict__ keys) {
  const BrBlk B = blk[blockIdx.x];
  const u32 hashable = P.n >= P.htl ? P.n - P.htl + 1 : 0;
  for (u32 p = B.start + threadIdx.x; p < B.end; p += blockDim.x)
    keys[p] = p < hashable ? br_quick_slot(P, br_ld64u(data, p), p - B.base) : P.nbuckets;
}

// qualities 2..4: where the search at position p enters each of its slots' segments (br_lz77.h br_quick_pred_fill); one CTA
// per input block, a thread per position
__global__ void k_slot_pred(BrStream s, u32* __restrict__ qpred) {
  const BrBlk B = s.blk[blockIdx.x];
  for (u32 p = B.start + threadIdx.x; p < B.end; p += blockDim.x)
    br_quick_pred_fill(s, p, B.base, qpred + ((size_t)p << s.P.qk_sweep_bits));
}

#define RADIX_TILE 4096
// digit histogram of one tile -> hist[d * ntiles + tile]   (KT: u16 bucket keys, u32 slot keys)
template <int SHIFT, class KT = u16>
__global__ void __launch_bounds__(256) k_radix_count(const KT* __restrict__ keys, u32 n, u32* __restrict__ hist, u32 ntiles) {
  __shared__ u32 cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  u32 base = blockIdx.x * RADIX_TILE;
  for (u32 i = threadIdx.x; i < RADIX_TILE; i += 256) {
    u32 j = base + i;
    if (j < n) atomicAdd(&cnt[(keys[j] >> SHIFT) & 0xFF], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}
// stable scatter: warp w of the CTA owns elements [w*512, w*512+512) of the tile, row by row
// `inv` (nullable): the final pass also writes the inverse permutation inv[position] = index in the sorted order (BrStream::rank)
template <int SHIFT, bool HAS_VALS, class KT = u16>
__global__ void __launch_bounds__(256) k_radix_scatter(const KT* __restrict__ keys, const u32* __restrict__ vals, u32 n,
    const u32* __restrict__ hist_scanned, u32 ntiles, KT* __restrict__ out_keys, u32* __restrict__ out_vals, u32* __restrict__ inv) {
  __shared__ u32 wcnt[8][256];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (u32 i = threadIdx.x; i < 8 * 256; i += 256) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const u32 wbase = blockIdx.x * RADIX_TILE + warp * 512;
  for (u32 r = 0; r < 16; ++r) {
    u32 j = wbase + r * 32 + lane;
    u32 d = j < n ? ((keys[j] >> SHIFT) & 0xFFu) : 0xFFFFFFFFu;
    u32 m = __match_any_sync(0xffffffffu, d);
    if (d != 0xFFFFFFFFu && (m & ((1u << lane) - 1)) == 0) wcnt[warp][d] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  {
    u32 d = threadIdx.x;
    u32 running = hist_scanned[d * ntiles + blockIdx.x];
    for (u32 w = 0; w < 8; ++w) { u32 t = wcnt[w][d]; wcnt[w][d] = running; running += t; }
  }
  __syncthreads();
  for (u32 r = 0; r < 16; ++r) {
    u32 j = wbase + r * 32 + lane;
    u32 k = j < n ? keys[j] : 0;
    u32 d = j < n ? ((k >> SHIFT) & 0xFFu) : 0xFFFFFFFFu;
    u32 m = __match_any_sync(0xffffffffu, d);
    u32 rank = __popc(m & ((1u << lane) - 1));
    if (j < n) {
      u32 dst = wcnt[warp][d] + rank;
      out_keys[dst] = (KT)k;
      const u32 v = HAS_VALS ? vals[j] : j;
      out_vals[dst] = v;
      if (inv) inv[v] = dst;
    }
    __syncwarp();
    if (d != 0xFFFFFFFFu && rank == 0) wcnt[warp][d] += __popc(m);
    __syncwarp();
  }
}
// chunk table of one input block (what br_build_blocks would fill on the host): one CTA per block
__global__ void k_build_chunks(const BrBlk* __restrict__ blk, BrBlockIn* __restrict__ bin, u32 ch) {
  const BrBlk B = blk[blockIdx.x];
  for (u32 c = threadIdx.x; c < B.nchunks; c += blockDim.x) {
    BrBlockIn k; memset(&k, 0, sizeof(k));
    k.pos = B.start + c * ch; k.end = k.pos + ch < B.end ? k.pos + ch : B.end; k.blk_start = B.start; k.blk_end = B.end;
    k.first = c == 0; k.last = k.end == B.end; k.is_last = B.is_last; k.force_flush = B.force_flush; k.blk = blockIdx.x; k.base = B.base;
    bin[B.first_chunk + c] = k;
  }
}
// tagS[j] = tag of the position S[j] (br_lz77.h br_tag4)
__global__ void k_tags(const u8* __restrict__ data, const u32* __restrict__ S, u32 n, u16* __restrict__ tagS) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) tagS[j] = (u16)br_tag4(br_ld32u(data, S[j]));
}
// seg[k] = first index in S whose key is >= k, for k in [0, nbuckets + 1]
template <class KT>
__global__ void k_seg(const KT* __restrict__ sorted_keys, u32 n, u32 nbuckets, u32* __restrict__ seg) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  int k0 = j == 0 ? -1 : (int)sorted_keys[j - 1];
  int k1 = j == n ? (int)nbuckets + 1 : (int)sorted_keys[j];
  for (int k = k0 + 1; k <= k1; ++k) seg[k] = j;
}

// ---------------------------------------------------------------------------- exclusive scan
#define SCAN_CHUNK 2048
__global__ void __launch_bounds__(256) k_scan_chunks(u32* __restrict__ a, u32 n, u32* __restrict__ sums) {
  __shared__ u32 part[256];
  u32 base = blockIdx.x * SCAN_CHUNK + threadIdx.x * 8;
  u32 v[8], acc = 0;
  for (int i = 0; i < 8; ++i) { v[i] = base + i < n ? a[base + i] : 0; acc += v[i]; }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    u32 t = threadIdx.x >= (u32)o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  u32 run = part[threadIdx.x] - acc;
  for (int i = 0; i < 8; ++i) { if (base + i < n) a[base + i] = run; run += v[i]; }
  if (threadIdx.x == 255) sums[blockIdx.x] = part[255];
}
__global__ void k_scan_add(u32* __restrict__ a, u32 n, const u32* __restrict__ sums) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += sums[i / SCAN_CHUNK];
}
// in-place exclusive scan of a[0..n); tmp must hold >= n/SCAN_CHUNK + n/SCAN_CHUNK^2 + 8 words
static void scan_exclusive(u32* a, u32 n, u32* tmp, cudaStream_t st) {
  if (n == 0) return;
  u32 chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  k_scan_chunks<<<chunks, 256, 0, st>>>(a, n, tmp);
  if (chunks > 1) {
    scan_exclusive(tmp, chunks, tmp + chunks, st);
    k_scan_add<<<(n + 255) / 256, 256, 0, st>>>(a, n, tmp);
  }
}

// ---------------------------------------------------------------------------- stored bits
// Initial speculation: every hashable position of a block is stored (true for > 99.8 % of
// text positions).  One CTA per input block (blocks cut by a FLUSH are irregular); words shared
// by two blocks are ORed into a zeroed bitmap.
__global__ void k_init_bits(BrStream s) {
  const BrBlk B = s.blk[blockIdx.x];
  const u32 htl = s.P.htl;
  // positions of this block that the parse stores: those with a full hash load inside the block ...
  u32 lo = B.start, hi = B.end - B.start >= htl ? B.end - htl + 1 : B.start;
  // ... and the last three, stored by StitchToPreviousBlock of the next block (hash_longest_match64_inc.h:127)
  u32 s_lo = 0, s_hi = 0;
  if (blockIdx.x + 1 < s.nblk) {
    const BrBlk N = s.blk[blockIdx.x + 1];
    if (N.end - N.start >= htl - 1 && N.start >= 3) { s_lo = N.start - 3 > B.start ? N.start - 3 : B.start; s_hi = N.start; }
  }
  const u32 w0 = B.start >> 5, w1 = (B.end - 1) >> 5;
  for (u32 wi = w0 + threadIdx.x; wi <= w1; wi += blockDim.x) {
    u32 v = 0;
    for (u32 b = 0; b < 32; ++b) {
      const u32 p = wi * 32 + b;
      if ((p >= lo && p < hi) || (p >= s_lo && p < s_hi)) v |= 1u << b;
    }
    if (v) atomicOr(s.bits_latest + wi, v);
  }
}
// bits_latest permuted into S order + per-1024 popcounts
__global__ void __launch_bounds__(1024) k_build_storedS(BrStream s, u32* __restrict__ storedS, u32* __restrict__ blockcnt) {
  __shared__ u32 wsum[32];
  u32 j = blockIdx.x * 1024 + threadIdx.x;
  int bit = 0;
  if (j < s.P.n) { u32 q = s.S[j]; bit = (s.bits_latest[q >> 5] >> (q & 31)) & 1; }
  u32 m = __ballot_sync(0xffffffffu, bit);
  if ((threadIdx.x & 31) == 0) { storedS[j >> 5] = m; wsum[threadIdx.x >> 5] = __popc(m); }
  __syncthreads();
  if (threadIdx.x < 32) {
    u32 v = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) blockcnt[blockIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------- warp-task kernels
// G = bucket-ring rows fetched together (br_lz77.h): 1 for the 16/32-entry rings of quality 5-6 (throughput bound,
// 8 CTAs per SM), 4 for the 64-256-entry rings of quality 7-9 (more registers, fewer resident warps).
#ifndef BR_WALK1_MINB
#define BR_WALK1_MINB 8
#endif
#ifndef BR_WALK_G_DEEP
#define BR_WALK_G_DEEP 4    /* rows fetched together for the 256-entry rings of quality 9: 8 (the whole ring, 125 registers,
                               16 warps per SM) walks one chunk faster, 4 (96 registers, 20 warps per SM) wins on the whole
                               job -- config C4 5.37 s against 5.90 s (profiles/r02p_variants.log) */
#endif
#ifndef BR_WALK_G_SMALL
#define BR_WALK_G_SMALL 1   /* rows fetched together for the 16/32-entry rings */
#endif
template <int G, bool M>
__global__ void __launch_bounds__(128, G <= 2 ? BR_WALK1_MINB : G == 4 ? 5 : 4) k_walk(BrStream s) {
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= s.counters[5]) return;
  const u32 b = br_sched_entry(s, t);
  br_walk_block<G, M>(s, b, s.forced && b == s.counters[6]);
}
// positions covered by the runs of this launch (br_chain.h br_cover_run); before k_commit
__global__ void k_cover(BrStream s) {
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= s.counters[4]) return;
  br_cover_run(s, s.ran_list[t]);
}
__global__ void k_commit(BrStream s) {
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= s.counters[4]) return;
  br_commit_bits(s, s.ran_list[t]);
}
// qualities 2..4: every run against the committed stored-bits (br_chain.h br_verify_run); warp per chunk
__global__ void k_verify(BrStream s) {
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= s.P.nblocks) return;
  br_verify_run(s, t);
}
__global__ void k_chain_a(BrStream s) {
  u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi < s.nblk) br_chain_a(s, bi);
}
__global__ void __launch_bounds__(32) k_chain_b(BrStream s) { br_chain_b(s); }
__global__ void __launch_bounds__(128) k_chain_b1(BrStream s) {   // batch of streams: a warp per stream
  const u32 st = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (st < s.P.multi) br_chain_b1(s, st);
}
__global__ void __launch_bounds__(32) k_chain_b2(BrStream s) { br_chain_b2(s); }
__global__ void k_chain_b3(BrStream s) {
  const u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi < s.nblk) br_chain_b3(s, bi);
}
__global__ void k_chain_c(BrStream s) {
  u32 bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi < s.nblk) br_chain_c(s, bi);
}
__global__ void k_chain_d(BrStream s) {
  u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < s.P.nblocks) br_chain_d(s, k);
}
__global__ void k_compact(BrStream s, BrCmd* cmds_all, const u32* __restrict__ block_mb) {
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= s.P.nblocks) return;
  br_compact_block(s, t, cmds_all, block_mb);
}
// ---- data-parallel entropy stage (br_entropy2.h)
__global__ void k_cmd_scan_inputs(const BrCmd* __restrict__ cmds, u32 C, u32* ins, u32* span, u32* hasd) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > C) return;
  u32 a = 0, b = 0, d = 0;
  if (i < C) br_cmd_scan_inputs(cmds[i], &a, &b, &d);
  ins[i] = a; span[i] = b; hasd[i] = d;
}
__global__ void k_cmd_mb(BrStream s, u32* cmd_mb, u32 n_mbs) {
  // (grid.y is capped at 65535: 16 KiB input blocks of noise at quality 2, 3 make up to 65536 metablocks per GiB)
  for (u32 m = blockIdx.y; m < n_mbs; m += gridDim.y) {
    const BrMetaBlock mb = s.mbs[m];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < mb.ncmd; i += gridDim.x * blockDim.x) cmd_mb[mb.cmd_off + i] = m;
  }
}
__global__ void k_expand(BrEnt e) {
  u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w * 32 >= e.total_cmds) return;
  br_expand_cmds(e, w * 32);
}
// one warp per metablock: literal context model + bookkeeping
__global__ void __launch_bounds__(32) k_mb_setup(BrStream s, BrEnt e) {
  const u32 i = blockIdx.x;
  const BrMetaBlock mb = s.mbs[i];
  if (!mb.compress) return;
  u8* sc = e.scratch + e.scratch_off[i];
  BrMbMem* M = (BrMbMem*)sc;
  u32 which = s.P.disable_ctx ? 1u : br_decide_context_modeling(s, mb.start, mb.end - mb.start, M->sc.rle_syms);
  if (threadIdx.x == 0) {
    BrMbAux a; memset(&a, 0, sizeof(a));
    a.which = which;
    a.lit_base = e.lit_ord[mb.cmd_off]; a.dist_base = e.dist_ord[mb.cmd_off];
    a.nsym[0] = mb.nlit; a.nsym[1] = mb.ncmd; a.nsym[2] = e.dist_ord[mb.cmd_off + mb.ncmd] - a.dist_base;
    u32 off = br_align8((u32)sizeof(BrMbMem));
    for (int cat = 0; cat < 3; ++cat) { a.var_off[cat] = off; off += br_mb2_var_bytes(br_mb2_nblk(cat, mb.nlit, mb.ncmd)); }
    e.aux[i] = a;
  }
}
__global__ void __launch_bounds__(256) k_split(BrStream s, BrEnt e) {
  extern __shared__ u32 sm[];
  const u32 i = blockIdx.x; const int cat = (int)blockIdx.y;
  const BrMetaBlock mb = s.mbs[i];
  if (!mb.compress) return;
  BrMbAux& a = e.aux[i];
  u8* sc = e.scratch + e.scratch_off[i];
  BrMbMem* M = (BrMbMem*)sc;
  const u32 A = cat == 0 ? 256u : cat == 1 ? 704u : 64u, nc = cat == 0 ? a.which : 1u;
  const u32 minb = cat == 1 ? 1024u : 512u;
  const double thr = cat == 0 ? 400.0 : cat == 1 ? 500.0 : 100.0;
  u32* H = cat == 0 ? M->lit_H : cat == 1 ? M->cmd_H : M->dist_H;
  u32 nblk = br_mb2_nblk(cat, mb.nlit, mb.ncmd);
  br_split_cta(s, e, mb, a, cat, A, nc, minb, thr, a.nsym[cat], br_mb2_types(sc, a, cat, nblk),
               br_mb2_lengths(sc, a, cat, nblk), H, sm);
}
__global__ void __launch_bounds__(64) k_prep(BrStream s, BrEnt e) {
  const u32 i = blockIdx.x;
  const BrMetaBlock mb = s.mbs[i];
  if (!mb.compress) return;
  br_prep_codes(s, mb, e.aux[i], e.scratch + e.scratch_off[i], e.outbits + e.out_off[i]);
}
// qualities 2, 3: histograms + the three codes of a metablock without block splits (br_entropy_flat.h)
__global__ void __launch_bounds__(64) k_prep_flat(BrStream s, BrEnt e) {
  const u32 i = blockIdx.x;
  const BrMetaBlock mb = s.mbs[i];
  if (!mb.compress) return;
  br_prep_flat(s, e, mb, e.aux[i], e.scratch + e.scratch_off[i], e.outbits + e.out_off[i]);
}
__global__ void k_lit_bits(BrStream s, BrEnt e) {
  u32 o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < e.total_lits) br_lit_bits(s, e, o);
  else if (o == e.total_lits) e.lit_len[o] = 0;
}
__global__ void k_cmd_bits(BrStream s, BrEnt e) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e.total_cmds) br_cmd_bits(s, e, i);
  else if (i == e.total_cmds) e.cmd_len[i] = 0;
}
__global__ void k_emit_cmd(BrStream s, BrEnt e) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e.total_cmds) br_emit_cmd(s, e, i);
}
__global__ void k_emit_lit(BrStream s, BrEnt e) {
  u32 o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < e.total_lits) br_emit_lit(s, e, o);
}

// ---------------------------------------------------------------------------- stream assembly
__global__ void k_assemble_scan(BrStream s, const u64* __restrict__ out_off, u32* out, BrCopyDesc* desc, u32* res,
                                int with_header, const u32* __restrict__ cut_kind, u64* cut_end_bit, u64* stream_end) {
  if (threadIdx.x == 0) br_assemble_scan(s, out_off, out, desc, res, with_header, cut_kind, cut_end_bit, stream_end);
}
// grid.y = metablock, grid.x strides over its words / bytes
__global__ void k_assemble_copy(BrStream s, const BrCopyDesc* __restrict__ desc, const u32* __restrict__ outbits, u32* out, u32 n_mbs) {
  for (u32 m = blockIdx.y; m < n_mbs; m += gridDim.y) {   // (grid.y is capped at 65535, see k_cmd_mb)
  const BrCopyDesc d = desc[m];
  if (d.kind == 0) {
    u32 nwords = (d.nbits + 31) / 32;
    const u32* src = outbits + d.src_off;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
      u32 nb = (i + 1) * 32 <= d.nbits ? 32 : d.nbits - i * 32;
      u32 v = src[i];
      if (nb < 32) v &= (1u << nb) - 1;
      u64 db = d.dst_bit + (u64)i * 32;
      br_put_bits_at(out + (db >> 5), (u32)(db & 31), nb, v);
    }
  } else {
    // raw bytes: destination is byte aligned; go through 32-bit words of the destination
    u64 dst_byte = d.dst_bit >> 3;
    u32 nbytes = d.nbits;
    u64 first_word = dst_byte >> 2, last_word = (dst_byte + nbytes - 1) >> 2;
    u32 nw = (u32)(last_word - first_word + 1);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) {
      u64 wbyte = (first_word + i) << 2;
      u32 v = 0;
      for (u32 b = 0; b < 4; ++b) {
        u64 ob = wbyte + b;
        if (ob >= dst_byte && ob < dst_byte + nbytes) v |= (u32)s.data[d.src_off + (ob - dst_byte)] << (8 * b);
      }
      if (v) atomicOr(out + first_word + i, v);
    }
  }
}
}

// ============================================================================ host pipeline
struct BrDeviceTables {
  int device = -1;
  u8* blob = nullptr;       // brotli_tables.bin on the device
  double* log2tab = nullptr;
  u32 log2tab_n = 0;
};
static BrDeviceTables g_tables[16];
static std::mutex g_tables_mu;

extern "C" const unsigned char br_tables_blob[];
extern "C" const unsigned int br_tables_blob_len;
extern "C" const double* br_host_log2_table(u32* n);   // br_host.cc

// Per-device constant tables, uploaded once.  Jobs of several host threads (BrotliB200CompressBatch) come here
// concurrently: the init is serialised, the uploads are synchronous copies that have completed before `device` is
// published, so kernels on any (non-blocking) stream launched afterwards see the finished tables.
static BrDeviceTables* get_tables() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  std::lock_guard<std::mutex> lock(g_tables_mu);
  BrDeviceTables& t = g_tables[dev];
  if (t.device == dev) return &t;
  u8* blob = nullptr; double* l2 = nullptr;
  u32 n = 0; const double* h = br_host_log2_table(&n);
  if (cudaMalloc(&blob, br_tables_blob_len) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaMalloc(&l2, (size_t)n * 8) != cudaSuccess) { cudaGetLastError(); cudaFree(blob); return nullptr; }
  if (cudaMemcpy(blob, br_tables_blob, br_tables_blob_len, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(l2, h, (size_t)n * 8, cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError(); cudaFree(blob); cudaFree(l2); return nullptr;
  }
  t.blob = blob; t.log2tab = l2; t.log2tab_n = n;
  t.device = dev;
  return &t;
}

// Grow-only device arena owned by a job; all per-stream arrays are carved out of it.
struct BrArena {
  u8* base = nullptr; size_t cap = 0, used = 0;
  bool reserve(size_t bytes) {
    if (bytes <= cap) { used = 0; return true; }
    if (base) cudaFree(base);
    base = nullptr; cap = 0;
    if (cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = bytes; used = 0; return true;
  }
  template <class T> T* take(size_t count) {
    size_t off = (used + 255) & ~(size_t)255;
    used = off + count * sizeof(T);
    if (used > cap) return nullptr;
    return (T*)(base + off);
  }
  void release() { if (base) cudaFree(base); base = nullptr; cap = 0; }
};

struct BrJob {
  cudaStream_t st = nullptr;
  BrArena arena, arena2;     // arena2: metablock scratch / bit buffers / output (sized after LZ77)
  u32* h_pinned = nullptr;    // small pinned readback area
  cudaEvent_t ev[10] = {};    // stage timers, created once per job
  BrJobStats stats;
};

extern "C" void br_job_destroy(BrJob* j);
extern "C" BrJob* br_job_create(void) {
  BrJob* j = new BrJob();
  if (cudaStreamCreateWithFlags(&j->st, cudaStreamNonBlocking) != cudaSuccess) { delete j; return nullptr; }
  if (cudaMallocHost(&j->h_pinned, 4096) != cudaSuccess) { cudaStreamDestroy(j->st); delete j; return nullptr; }
  for (auto& e : j->ev) if (cudaEventCreate(&e) != cudaSuccess) { br_job_destroy(j); return nullptr; }
  return j;
}
extern "C" void br_job_destroy(BrJob* j) {
  if (!j) return;
  j->arena.release(); j->arena2.release();
  for (auto& e : j->ev) if (e) cudaEventDestroy(e);
  if (j->h_pinned) cudaFreeHost(j->h_pinned);
  if (j->st) cudaStreamDestroy(j->st);
  delete j;
}
extern "C" const BrJobStats* br_job_stats(const BrJob* j) { return &j->stats; }
extern "C" void* br_job_stream(BrJob* j) { return (void*)j->st; }

static size_t scan_tmp_words(size_t n) { return n / SCAN_CHUNK + n / SCAN_CHUNK / SCAN_CHUNK + 64; }

// Compress one stream whose input already sits in device memory (d_in, n bytes).  The
// compressed bytes are left in device memory (*d_out, *out_size; valid until the next call on
// this job).  Returns 1 on success, 0 on failure (unsupported parameters, CUDA error).
// `cuts` (nullable): the stream is cut by FLUSH / EMIT_METADATA operations (br_pipeline.h BrCuts).
extern "C" int br_job_compress_device(BrJob* job, int quality, int lgwin, u32 size_hint,
                                      const u8* d_in, u32 n, const u8** d_out, size_t* out_size, const BrCuts* cuts) {
  BrDeviceTables* T = get_tables();
  if (!T || n == 0) return 0;
  BrStream s; memset(&s, 0, sizeof(s));
  if (!br_derive_params(quality, lgwin, size_hint, n, &s.P, cuts ? cuts->lgblock : 0)) return 0;
  BrParams& P = s.P;
  P.disable_ctx = cuts && cuts->disable_ctx ? 1u : 0u;
  P.stream_offset = cuts ? (cuts->stream_offset < P.max_backward ? cuts->stream_offset : P.max_backward) : 0u;   // encode.c:680
#ifdef BR_DEBUG_KNOBS   // experiment switches: never in the release build (an environment variable must not change the bytes)
  if (getenv("BR_HEAVY_MIN")) P.heavy_min = (u32)strtoul(getenv("BR_HEAVY_MIN"), 0, 10);
  if (getenv("BR_STEP_CAP")) P.step_cap = (u32)strtoul(getenv("BR_STEP_CAP"), 0, 10);
  if (getenv("BR_SWEEP_EPOCH")) P.sweep_epoch = (u32)strtoul(getenv("BR_SWEEP_EPOCH"), 0, 10);
  if (getenv("BR_FORCE_EPOCH")) P.force_epoch = (u32)strtoul(getenv("BR_FORCE_EPOCH"), 0, 10);
  const bool trace = getenv("BR_TRACE") != nullptr;
#else
  const bool trace = false;
#endif
  // chunk / block tables: the reference's input blocks (1 << lgblock bytes, shorter where a FLUSH cut the input)
  std::vector<BrBlk> hblk;
  const u32 ncuts = cuts ? cuts->n : 0;
  u32 nstreams = 1;
  for (u32 i = 0; i < ncuts; ++i) if (cuts->kind[i] == 3) ++nstreams;
  if (nstreams > 1) {
    // a batch of independent streams (each below BR_SMALL_STREAM, size_hint = the largest): no FLUSH cuts, all finished
    if (nstreams != ncuts + 1 || !cuts->is_final || !cuts->with_header || cuts->finish_empty || cuts->stream_offset ||
        cuts->pos[ncuts - 1] >= n || size_hint >= BR_SMALL_STREAM) return 0;
    P.multi = nstreams; P.pilot = 0;
    P.chunk_bits = br_batch_chunk_bits(n);
  }
  if (P.quick && P.stream_offset) return 0;   // (STREAM_OFFSET: quality 5..9)
  const u32 ch = 1u << P.chunk_bits;
  const bool is_final = cuts ? cuts->is_final != 0 : true;
  const int with_header = cuts ? cuts->with_header : 1;
  P.finish_empty = cuts && cuts->finish_empty ? 1u : 0u;
  if (!is_final && (ncuts == 0 || cuts->pos[ncuts - 1] != n)) return 0;   // an unfinished stream ends at a cut
  if (P.finish_empty && (!is_final || (ncuts && cuts->pos[ncuts - 1] == n))) return 0;
  u32 nb = 0;
  br_build_blocks(P, n, cuts ? cuts->pos : nullptr, ncuts, is_final && !P.finish_empty, nullptr, hblk, &nb, cuts ? cuts->kind : nullptr);   // chunks: k_build_chunks
  std::vector<u32> h_slot(((size_t)n >> P.lgblock) + 2, 0);
  { u32 b = 0; for (size_t i = 0; i < h_slot.size(); ++i) { const u64 p = (u64)i << P.lgblock; while (b + 1 < hblk.size() && hblk[b].end <= p) ++b; h_slot[i] = b; } }
  const u32 nblk = (u32)hblk.size();
  P.nblocks = nb;
  cudaStream_t st = job->st;
  memset(&job->stats, 0, sizeof(job->stats));
  cudaEvent_t* ev = job->ev;
  cudaEventRecord(ev[0], st);

  const u32 ntiles = (n + RADIX_TILE - 1) / RADIX_TILE;
  const size_t nwords = (size_t)(n + 31) / 32 + 2;
  const u32 cmd_stride = ch / 2 + 2;
  size_t need = 0;
  auto add = [&](size_t bytes) { need += (bytes + 255) & ~(size_t)255; };
  add((size_t)n + 64); add(2ull * n + 4); add(2ull * n + 4); add(4ull * n); add(2ull * n + 4); add(4ull * n); add(4ull * n); add(2ull * n + 4);
  add(256ull * ntiles * 4); add(scan_tmp_words(256ull * ntiles) * 4);
  add((P.nbuckets + 4) * 4ull); add(nwords * 4); add(2 * nwords * 4); add(nwords * 4); add(nwords * 4); add(nwords * 4); add(nwords * 4); add(nwords * 4); add((nwords + 64) * 4); add(((size_t)n / 1024 + 8) * 4);
  add(scan_tmp_words((size_t)n / 1024 + 8) * 4);
  add(nb * sizeof(BrBlockIn) * 2); add(nb * sizeof(BrBlockOut)); add((size_t)nb * cmd_stride * sizeof(BrCmd));
  for (int i = 0; i < 14; ++i) add(nb * 4ull + 64);
  add(nblk * sizeof(BrBlk)); add(nblk * sizeof(BrBlkIn)); add((P.nbuckets + 8) * 4ull);
  add(h_slot.size() * 4); add((ncuts + 2) * 4ull); add((ncuts + 2) * 8ull); add((nstreams + 2) * 8ull);
  add((nstreams + 2) * 4ull * 3); add(nblk * sizeof(BrMetaBlock)); add(1024ull * nstreams);
  // Launch bound: in forced mode every launch finalises at least one input block (at most two launches per block with
  // the conservative block-level re-marks), and a late uncompressed fallback restarts the count at most once per metablock.
  P.max_epochs = 4 * nb + 4096;
  add(((size_t)P.max_epochs + 2) * 4 + 64); add((nb + 1) * sizeof(BrMetaBlock)); add(4096);
  if (P.quick) { for (int i = 0; i < 4; ++i) add(4ull * n + 16); add((((size_t)n << P.qk_sweep_bits) + 16) * 4); add((((size_t)n << P.qk_sweep_bits) + 16) * 4); }   // 32-bit sort keys, slot entry points, slot reads
  need += 1 << 20;
  if (!job->arena.reserve(need)) return 0;
  BrArena& A = job->arena;
  u8* data = A.take<u8>((size_t)n + 64);
  u16* keys = A.take<u16>((size_t)n + 2); u16* K1 = A.take<u16>((size_t)n + 2); u32* V1 = A.take<u32>(n);
  u16* K2 = A.take<u16>((size_t)n + 2); u32* S = A.take<u32>(n); u32* rank = A.take<u32>(n); u16* tagS = A.take<u16>((size_t)n + 2);
  u32* hist = A.take<u32>(256ull * ntiles); u32* scan_tmp = A.take<u32>(scan_tmp_words(256ull * ntiles));
  u32* seg = A.take<u32>(P.nbuckets + 4);
  u32* bits_latest = A.take<u32>(nwords); u32* bits_cur = A.take<u32>(2 * (size_t)nwords);
  u32* srch_latest = A.take<u32>(nwords); u32* srch_cur = A.take<u32>(nwords); u32* bits_prev = A.take<u32>(nwords);
  u32* cover_cur = A.take<u32>(nwords);
  u32* storedS = A.take<u32>(nwords + 64); u32* prefS = A.take<u32>((size_t)n / 1024 + 8);
  u32* scan_tmp2 = A.take<u32>(scan_tmp_words((size_t)n / 1024 + 8));
  BrBlockIn* bin = A.take<BrBlockIn>(nb); BrBlockIn* bin_used = A.take<BrBlockIn>(nb);
  BrBlockOut* bout = A.take<BrBlockOut>(nb);
  BrCmd* cmd_blocks = A.take<BrCmd>((size_t)nb * cmd_stride);
  u32* dirty = A.take<u32>(nb + 16); u32* changed_bits = A.take<u32>(nb + 16);
  int* changed_epoch = A.take<int>(nb + 16); u32* ext_total = A.take<u32>(nb + 16);
  u32* cmd_off = A.take<u32>(nb + 16); u32* force_unc = A.take<u32>(nb + 16);
  u32* dirty_list = A.take<u32>(nb + 16); u32* block_mb = A.take<u32>(nb + 16);
  u32* ran_list = A.take<u32>(nb + 16); u32* lil_in = A.take<u32>(nb + 16);
  int* bitdep_epoch = A.take<int>(nb + 16);
  BrBlk* d_blk = A.take<BrBlk>(nblk); BrBlkIn* d_blkin = A.take<BrBlkIn>(nblk);
  u32* key_flips = A.take<u32>(P.nbuckets + 8);
  u32* slot_blk = A.take<u32>(h_slot.size()); u32* d_cut_kind = A.take<u32>(ncuts + 2); u64* d_cut_end = A.take<u64>(ncuts + 2);
  u64* d_stream_end = A.take<u64>(nstreams + 2);
  u32* epoch_cum = A.take<u32>((size_t)P.max_epochs + 2);
  BrMetaBlock* mbs = A.take<BrMetaBlock>(nb + 1);
  u32* d_stream_blk = A.take<u32>(nstreams + 2); u32* stream_nmb = A.take<u32>(nstreams + 2); u32* stream_ncmd = A.take<u32>(nstreams + 2);
  BrMetaBlock* mbs_stage = A.take<BrMetaBlock>(nblk);
  u32* counters = A.take<u32>(64); u32* hist_scratch = A.take<u32>(256ull * nstreams);
  if (!hist_scratch) return 0;
  u32 *qk_keys = nullptr, *qk_K1 = nullptr, *qk_V1 = nullptr, *qk_V2 = nullptr, *saw = nullptr, *qpred = nullptr;
  if (P.quick) {
    qk_keys = A.take<u32>((size_t)n + 4); qk_K1 = A.take<u32>((size_t)n + 4); qk_V1 = A.take<u32>((size_t)n + 4); qk_V2 = A.take<u32>((size_t)n + 4);
    saw = A.take<u32>(((size_t)n << P.qk_sweep_bits) + 16);
    qpred = A.take<u32>(((size_t)n << P.qk_sweep_bits) + 16);
    if (!qpred) return 0;
  }
  s.saw = saw; s.qpred = qpred;

  CK(cudaMemcpyAsync(data, d_in, n, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(data + n, 0, 64, st));
  s.data = data; s.S = S; s.rank = rank; s.seg = seg; s.bits_latest = bits_latest; s.bits_cur = bits_cur; s.bits_words = (u32)nwords;
  s.storedS = storedS; s.prefS = prefS; s.bin = bin; s.bin_used = bin_used; s.bout = bout;
  s.cmd_blocks = cmd_blocks; s.cmd_stride = cmd_stride; s.dirty = dirty; s.changed_bits = changed_bits;
  s.changed_epoch = changed_epoch; s.epoch_cum = epoch_cum;
  s.ext_total = ext_total; s.cmd_off = cmd_off; s.mbs = mbs; s.force_unc = force_unc;
  s.counters = counters; s.hist_scratch = hist_scratch;
  s.srch_latest = srch_latest; s.srch_cur = srch_cur; s.bits_prev = bits_prev; s.cover_cur = cover_cur; s.bitdep_epoch = bitdep_epoch; s.skeys = K2;
  s.dirty_list = dirty_list; s.block_mb = block_mb; s.ran_list = ran_list; s.lil_in = lil_in; s.blk = d_blk; s.nblk = nblk; s.blkin = d_blkin; s.key_flips = key_flips;
  { const u8* p = T->blob + 8;
    s.dict_size_bits = p; p += 32; s.dict_offsets = (const u32*)p; p += 128; s.dict = p; p += 122784;
    s.dict_hash_words = (const u16*)p; p += 65536; s.dict_hash_lengths = p; p += 32768; s.ctx_lut = p; }
  s.log2tab = T->log2tab; s.log2tab_n = T->log2tab_n;

  CK(cudaMemcpyAsync(d_blk, hblk.data(), nblk * sizeof(BrBlk), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(slot_blk, h_slot.data(), h_slot.size() * 4, cudaMemcpyHostToDevice, st));
  std::vector<u32> h_stream_blk(nstreams + 1, nblk);
  for (u32 bi = nblk; bi-- > 0;) h_stream_blk[hblk[bi].stream] = bi;
  CK(cudaMemcpyAsync(d_stream_blk, h_stream_blk.data(), (nstreams + 1) * 4ull, cudaMemcpyHostToDevice, st));
  s.stream_blk = d_stream_blk; s.stream_nmb = stream_nmb; s.stream_ncmd = stream_ncmd; s.mbs_stage = mbs_stage;
  if (ncuts) CK(cudaMemcpyAsync(d_cut_kind, cuts->kind, ncuts * 4, cudaMemcpyHostToDevice, st));
  s.slot_blk = slot_blk;
  CK(cudaMemsetAsync(bin, 0, nb * sizeof(BrBlockIn), st));
  k_build_chunks<<<nblk, 64, 0, st>>>(d_blk, bin, 1u << P.chunk_bits);
  CK(cudaStreamSynchronize(st));   // (hblk / h_slot are host vectors: their copies must be done before they go out of scope ... and before the arrays below are reused)
  CK(cudaMemsetAsync(bout, 0, nb * sizeof(BrBlockOut), st));
  CK(cudaMemsetAsync(bin_used, 0, nb * sizeof(BrBlockIn), st));
  CK(cudaMemsetAsync(changed_bits, 0, (nb + 16) * 4, st));
  CK(cudaMemsetAsync(changed_epoch, 0xFF, (nb + 16) * 4, st));
  CK(cudaMemsetAsync(bitdep_epoch, 0xFF, (nb + 16) * 4, st));
  CK(cudaMemsetAsync(srch_latest, 0, nwords * 4, st));
  CK(cudaMemsetAsync(key_flips, 0, (P.nbuckets + 8) * 4, st));
  CK(cudaMemsetAsync(force_unc, 0, (nb + 16) * 4, st));
  CK(cudaMemsetAsync(epoch_cum, 0, 64, st));
  CK(cudaMemsetAsync(counters, 0, 256, st));
  CK(cudaMemsetAsync(bits_cur, 0, 2 * (size_t)nwords * 4, st));

  // ---- position index: S, rank, seg
  if (P.quick) {
    // sorted by (slot, position): three stable 8-bit passes over 32-bit keys (slots have up to 20 bits + the overflow key)
    k_slot_keys<<<nblk, 256, 0, st>>>(P, data, d_blk, qk_keys);
    k_radix_count<0, u32><<<ntiles, 256, 0, st>>>(qk_keys, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, scan_tmp, st);
    k_radix_scatter<0, false, u32><<<ntiles, 256, 0, st>>>(qk_keys, nullptr, n, hist, ntiles, qk_K1, qk_V1, nullptr);
    k_radix_count<8, u32><<<ntiles, 256, 0, st>>>(qk_K1, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, scan_tmp, st);
    k_radix_scatter<8, true, u32><<<ntiles, 256, 0, st>>>(qk_K1, qk_V1, n, hist, ntiles, qk_keys, qk_V2, nullptr);
    k_radix_count<16, u32><<<ntiles, 256, 0, st>>>(qk_keys, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, scan_tmp, st);
    k_radix_scatter<16, true, u32><<<ntiles, 256, 0, st>>>(qk_keys, qk_V2, n, hist, ntiles, qk_K1, S, rank);
    k_seg<u32><<<(n + 1 + 255) / 256, 256, 0, st>>>(qk_K1, n, P.nbuckets, seg);
    k_slot_pred<<<nblk, 256, 0, st>>>(s, qpred);   // (s.S / s.rank / s.seg / s.blk / s.data are set; the kernel reads nothing else)
  } else {
  k_hash_keys<<<(n + 255) / 256, 256, 0, st>>>(P, data, keys);
  k_radix_count<0><<<ntiles, 256, 0, st>>>(keys, n, hist, ntiles);
  scan_exclusive(hist, 256 * ntiles, scan_tmp, st);
  k_radix_scatter<0, false><<<ntiles, 256, 0, st>>>(keys, nullptr, n, hist, ntiles, K1, V1, nullptr);
  k_radix_count<8><<<ntiles, 256, 0, st>>>(K1, n, hist, ntiles);
  scan_exclusive(hist, 256 * ntiles, scan_tmp, st);
  k_radix_scatter<8, true><<<ntiles, 256, 0, st>>>(K1, V1, n, hist, ntiles, K2, S, rank);
  k_seg<u16><<<(n + 1 + 255) / 256, 256, 0, st>>>(K2, n, P.nbuckets, seg);
  }
  // (the 16/32-entry rings of quality 5-6 gain nothing from tags -- measured: C2 58.1 ms with, 56.2 ms without -- so only
  // the deep rings of quality 7-9 use them: config C4 5.46 s -> 3.36 s)
  if (P.block_bits >= 6) k_tags<<<(n + 255) / 256, 256, 0, st>>>(data, S, n, tagS);
  s.tagS = tagS;
  CK(cudaMemsetAsync(bits_latest, 0, nwords * 4, st));
  k_init_bits<<<nblk, 256, 0, st>>>(s);
  cudaEventRecord(ev[1], st);

  // ---- LZ77 fixpoint
  u32* hp = job->h_pinned;
  u32 n_mbs = 0, total_cmds = 0;
  int rounds = 0;
  bool walk_pending = false;
  const u8* final_out = nullptr; size_t final_size = 0;
  for (;;) {   // rounds: repeated only when a metablock needs the late uncompressed fallback
    ++rounds;
    const u32 round_epoch0 = s.epoch;
    for (;;) {
      k_chain_a<<<(nblk + 31) / 32, 32, 0, st>>>(s);
      if (P.multi) {
        k_chain_b1<<<(P.multi + 3) / 4, 128, 0, st>>>(s);
        k_chain_b2<<<1, 32, 0, st>>>(s);
        k_chain_b3<<<(nblk + 255) / 256, 256, 0, st>>>(s);
      } else k_chain_b<<<1, 32, 0, st>>>(s);
      k_chain_c<<<(nblk + 31) / 32, 32, 0, st>>>(s);
      k_chain_d<<<(nb + 255) / 256, 256, 0, st>>>(s);
      CK(cudaMemcpyAsync(hp, counters, 128, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      u32 n_dirty = hp[0], n_sched = hp[5]; n_mbs = hp[1]; total_cmds = hp[2];
      job->stats.block_runs += hp[4];
      if (trace) fprintf(stderr, "epoch %u: dirty %u sched %u ran %u longest sweep %u | never %u state %u dict %u bits %u wrap %u\n", s.epoch, hp[0], hp[5], hp[4], hp[16], hp[9], hp[10], hp[11], hp[12], hp[13]);
      if (walk_pending) {
        float wms; cudaEventElapsedTime(&wms, ev[6], ev[7]); job->stats.ms_walk += wms; walk_pending = false;
        if (trace) fprintf(stderr, "   walk launch %u: %.2f ms\n", s.epoch, wms);
      }
      if (n_dirty == 0) break;
      if (s.epoch + 2 >= P.max_epochs) { fprintf(stderr, "brotli_b200: internal error: launch bound exceeded\n"); return 0; }
      ++s.epoch; ++job->stats.lz77_iterations;
      s.forced = s.epoch - round_epoch0 >= P.force_epoch ? 1u : 0u;
      CK(cudaMemsetAsync(bits_cur, 0, 2 * (size_t)nwords * 4, st));
      CK(cudaMemsetAsync(srch_cur, 0, nwords * 4, st));
      CK(cudaMemsetAsync(counters + 4, 0, 4, st));
      CK(cudaMemsetAsync(counters + 16, 0, 4, st));
      if (!P.quick) {   // (the S-ordered copy of the stored bits serves the bucket rings of quality 5..9)
        k_build_storedS<<<(n + 1023) / 1024, 1024, 0, st>>>(s, storedS, prefS);
        scan_exclusive(prefS, (n + 1023) / 1024, scan_tmp2, st);
      }
      cudaEventRecord(ev[6], st);
      const u32 wg = (n_sched + 3) / 4;
      if (P.quick) {
        if (P.multi) k_walk<0, true><<<wg, 128, 0, st>>>(s);
        else k_walk<0, false><<<wg, 128, 0, st>>>(s);
      } else if (P.multi) {
        if (P.block_bits >= 6) k_walk<BR_WALK_G_DEEP, true><<<wg, 128, 0, st>>>(s);
        else k_walk<BR_WALK_G_SMALL, true><<<wg, 128, 0, st>>>(s);
      } else if (P.block_bits >= 6) k_walk<BR_WALK_G_DEEP, false><<<wg, 128, 0, st>>>(s);
      else k_walk<BR_WALK_G_SMALL, false><<<wg, 128, 0, st>>>(s);
      cudaEventRecord(ev[7], st);
      walk_pending = true; ++job->stats.walk_launches; job->stats.launches += 6;
      job->stats.walk_bytes += (u64)n_sched * ch;
      CK(cudaMemcpyAsync(bits_prev, bits_latest, nwords * 4, cudaMemcpyDeviceToDevice, st));
      if (!P.quick) {   // (quality 2..4 check every run instead of marking: br_verify_run)
        CK(cudaMemsetAsync(cover_cur, 0, nwords * 4, st));
        k_cover<<<(nb * 32 + 127) / 128, 128, 0, st>>>(s);
      }
      k_commit<<<(nb * 32 + 127) / 128, 128, 0, st>>>(s);
      if (P.quick) k_verify<<<(nb * 32 + 127) / 128, 128, 0, st>>>(s);
    }
    cudaEventRecord(ev[2], st);
#ifdef BR_DEBUG_KNOBS
    if (trace) {   // the slowest chunk walks of the final state, and a per-megabyte profile
      std::vector<BrBlockOut> hb2(nb);
      cudaMemcpyAsync(hb2.data(), bout, nb * sizeof(BrBlockOut), cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
      const u32 per = (1u << 20) >> P.chunk_bits;
      for (u32 a = 0; a < nb; a += per) {
        unsigned long long cyc = 0, se = 0, ro = 0; u32 mx = 0;
        for (u32 k = a; k < a + per && k < nb; ++k) { cyc += hb2[k].dbg_kcycles; se += hb2[k].dbg_searches; ro += hb2[k].dbg_rows; if (hb2[k].dbg_kcycles > mx) mx = hb2[k].dbg_kcycles; }
        fprintf(stderr, "MiB %4u: %7.2f Mcycles per chunk (max %7.2f)  searches/chunk %5llu rows/search %5.1f\n", a / per, cyc / 1024.0 / per, mx / 1024.0, se / per, se ? (double)ro / se : 0.0);
      }
    }
#endif
    // ---- entropy stage
    std::vector<BrMetaBlock> hm(n_mbs);
    CK(cudaMemcpyAsync(hm.data(), mbs, n_mbs * sizeof(BrMetaBlock), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<u64> h_soff(n_mbs), h_ooff(n_mbs);
    size_t scratch_total = 0, outw_total = 0;
    for (u32 i = 0; i < n_mbs; ++i) {
      h_soff[i] = scratch_total; h_ooff[i] = outw_total;
      if (hm[i].compress) {
        scratch_total += ((size_t)br_mb_scratch_bytes(P, hm[i].nlit, hm[i].ncmd) + 255) & ~(size_t)255;
        outw_total += (2 * (size_t)(hm[i].end - hm[i].start) + 503) / 4 + 16;
      }
    }
    size_t out_cap = (size_t)n + ((size_t)n >> 3) + 4096 + 8ull * n_mbs;
    const size_t C = total_cmds;
    size_t need2 = scratch_total + outw_total * 4 + out_cap + (C + 2) * (sizeof(BrCmd) + 26) + ((size_t)n + 8) * 12 +
                   scan_tmp_words((size_t)n + C + 8) * 4 * 2 + n_mbs * (32 + sizeof(BrCopyDesc) + sizeof(BrMbAux)) + (1 << 16);
    if (!job->arena2.reserve(need2)) return 0;
    BrArena& B = job->arena2;
    u8* scratch = B.take<u8>(scratch_total + 256); u32* outbits = B.take<u32>(outw_total + 64);
    u32* out = B.take<u32>(out_cap / 4 + 16); BrCmd* cmds_all = B.take<BrCmd>(C + 1);
    u64* d_soff = B.take<u64>(n_mbs); u64* d_ooff = B.take<u64>(n_mbs);
    BrCopyDesc* desc = B.take<BrCopyDesc>(n_mbs); u32* res = B.take<u32>(16);
    u32* lit_ord = B.take<u32>(C + 2); u32* cmd_pos = B.take<u32>(C + 2); u32* dist_ord = B.take<u32>(C + 2);
    u32* cmd_len = B.take<u32>(C + 2); u32* lit_bit_base = B.take<u32>(C + 2); u32* cmd_mb = B.take<u32>(C + 2);
    u16* dist_sym = B.take<u16>(C + 2);
    u32* lit_pos = B.take<u32>((size_t)n + 2); u32* lit_cmd = B.take<u32>((size_t)n + 2); u32* lit_len = B.take<u32>((size_t)n + 2);
    BrMbAux* aux = B.take<BrMbAux>(n_mbs);
    u32* stmp = B.take<u32>(scan_tmp_words((size_t)n + C + 8) * 2);
    if (!stmp) return 0;
    CK(cudaMemcpyAsync(d_soff, h_soff.data(), n_mbs * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ooff, h_ooff.data(), n_mbs * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(outbits, 0, (outw_total + 64) * 4, st));
    CK(cudaMemsetAsync(out, 0, out_cap + 64, st));
    k_compact<<<(nb * 32 + 127) / 128, 128, 0, st>>>(s, cmds_all, block_mb);
    cudaEventRecord(ev[8], st);
    // E0: ordinals
    k_cmd_scan_inputs<<<(u32)((C + 1 + 255) / 256), 256, 0, st>>>(cmds_all, (u32)C, lit_ord, cmd_pos, dist_ord);
    scan_exclusive(lit_ord, (u32)C + 1, stmp, st);
    scan_exclusive(cmd_pos, (u32)C + 1, stmp, st);
    scan_exclusive(dist_ord, (u32)C + 1, stmp, st);
    CK(cudaMemcpyAsync(hp + 24, lit_ord + C, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 25, dist_ord + C, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    BrEnt e; memset(&e, 0, sizeof(e));
    e.cmds = cmds_all; e.total_cmds = (u32)C; e.total_lits = hp[24]; e.total_dist = hp[25];
    e.lit_ord = lit_ord; e.cmd_pos = cmd_pos; e.dist_ord = dist_ord; e.lit_pos = lit_pos; e.lit_cmd = lit_cmd;
    e.dist_sym = dist_sym; e.lit_len = lit_len; e.cmd_len = cmd_len; e.lit_bit_base = lit_bit_base; e.cmd_mb = cmd_mb;
    e.aux = aux; e.scratch = scratch; e.scratch_off = d_soff; e.outbits = outbits; e.out_off = d_ooff;
    if (e.total_lits > n) return 0;
    const u32 grid_mbs = n_mbs < 65535u ? n_mbs : 65535u;
    k_cmd_mb<<<dim3(64, grid_mbs), 256, 0, st>>>(s, cmd_mb, n_mbs);
    k_expand<<<(u32)((C + 127) / 128 + 1), 128, 0, st>>>(e);
    k_mb_setup<<<n_mbs, 32, 0, st>>>(s, e);
    if (P.mb_kind) k_prep_flat<<<n_mbs, 64, 0, st>>>(s, e);   // qualities 2, 3: no block split, one code per category
    else {
      const size_t smem_split = br_split_smem_bytes(256, 13) > br_split_smem_bytes(704, 1) ? br_split_smem_bytes(256, 13) : br_split_smem_bytes(704, 1);
      CK(cudaFuncSetAttribute(k_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_split));
      k_split<<<dim3(n_mbs, 3), 256, smem_split, st>>>(s, e);
      k_prep<<<n_mbs, 64, 0, st>>>(s, e);
    }
    k_lit_bits<<<(e.total_lits + 1 + 255) / 256, 256, 0, st>>>(s, e);
    k_cmd_bits<<<(u32)((C + 1 + 255) / 256), 256, 0, st>>>(s, e);
    scan_exclusive(lit_len, e.total_lits + 1, stmp, st);
    scan_exclusive(cmd_len, (u32)C + 1, stmp, st);
    k_emit_cmd<<<(u32)((C + 255) / 256), 256, 0, st>>>(s, e);
    if (e.total_lits) k_emit_lit<<<(e.total_lits + 255) / 256, 256, 0, st>>>(s, e);
    cudaEventRecord(ev[9], st);
    cudaEventRecord(ev[3], st);
    ++job->stats.encode_launches; job->stats.launches += 24;
    k_assemble_scan<<<1, 32, 0, st>>>(s, d_ooff, out, desc, res, with_header, d_cut_kind, d_cut_end, d_stream_end);
    CK(cudaMemcpyAsync(hp + 16, res, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hp[18]) {   // encode.c:604: a coded metablock is larger than its input -> it is stored raw (br_assemble_scan marked
                    // it in force_unc) and the parse behind it is redone with the restored distance cache
      if (rounds > (int)nblk + 8) return 0;   // (every round stores at least one more metablock raw)
      continue;
    }
    k_assemble_copy<<<dim3(64, grid_mbs), 256, 0, st>>>(s, desc, outbits, out, n_mbs);
    final_out = (const u8*)out; final_size = ((u64)hp[17] << 32) | hp[16];
    break;
  }
  cudaEventRecord(ev[4], st);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); job->stats.ms_index = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); job->stats.ms_lz77 = ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); job->stats.ms_entropy = ms;
  cudaEventElapsedTime(&ms, ev[3], ev[4]); job->stats.ms_assemble = ms;
  cudaEventElapsedTime(&ms, ev[0], ev[4]); job->stats.ms_total = ms;
  cudaEventElapsedTime(&ms, ev[8], ev[9]); job->stats.ms_encode = ms;
  job->stats.total_cmds = total_cmds; job->stats.launches += 12;
  job->stats.nblocks = nb; job->stats.n_metablocks = n_mbs; job->stats.rounds = (u32)rounds;
  job->stats.out_bytes = final_size; job->stats.in_bytes = n;
  if (cuts && cuts->end_bit && ncuts && nstreams == 1) {
    CK(cudaMemcpyAsync(cuts->end_bit, d_cut_end, ncuts * 8ull, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (cuts && cuts->stream_end && nstreams > 1) {
    CK(cudaMemcpyAsync(cuts->stream_end, d_stream_end, nstreams * 8ull, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  *d_out = final_out; *out_size = final_size;
  return 1;
}

// Debug / test hook: copies the position index of the last job to the host.
extern "C" __attribute__((visibility("default"))) int br_debug_sort(int quality, int lgwin, const u8* h_in, u32 n, u32* h_S, u32* h_seg) {
  BrParams P;
  if (!br_derive_params(quality, lgwin, n, n, &P)) return 0;
  if (P.quick) {   // qualities 2..4: the slot-sorted index (32-bit keys, three passes), one stream = one block from 0
    u8* data; u32 *k0, *k1, *v1, *v2, *S, *rank, *hist, *tmp, *seg; BrBlk* blk;
    const u32 ntiles = (n + RADIX_TILE - 1) / RADIX_TILE;
    BrBlk B; memset(&B, 0, sizeof(B)); B.start = 0; B.end = n; B.base = 0; B.send = n;
    CK(cudaMalloc(&data, (size_t)n + 64)); CK(cudaMemset(data, 0, (size_t)n + 64));
    CK(cudaMemcpy(data, h_in, n, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&blk, sizeof(BrBlk))); CK(cudaMemcpy(blk, &B, sizeof(BrBlk), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&k0, 4ull * n + 16)); CK(cudaMalloc(&k1, 4ull * n + 16)); CK(cudaMalloc(&v1, 4ull * n + 16)); CK(cudaMalloc(&v2, 4ull * n + 16));
    CK(cudaMalloc(&S, 4ull * n + 16)); CK(cudaMalloc(&rank, 4ull * n + 16)); CK(cudaMalloc(&hist, 1024ull * ntiles));
    CK(cudaMalloc(&tmp, scan_tmp_words(256ull * ntiles) * 4)); CK(cudaMalloc(&seg, (P.nbuckets + 4) * 4ull));
    k_slot_keys<<<1, 256>>>(P, data, blk, k0);
    k_radix_count<0, u32><<<ntiles, 256>>>(k0, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, tmp, 0);
    k_radix_scatter<0, false, u32><<<ntiles, 256>>>(k0, nullptr, n, hist, ntiles, k1, v1, nullptr);
    k_radix_count<8, u32><<<ntiles, 256>>>(k1, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, tmp, 0);
    k_radix_scatter<8, true, u32><<<ntiles, 256>>>(k1, v1, n, hist, ntiles, k0, v2, nullptr);
    k_radix_count<16, u32><<<ntiles, 256>>>(k0, n, hist, ntiles);
    scan_exclusive(hist, 256 * ntiles, tmp, 0);
    k_radix_scatter<16, true, u32><<<ntiles, 256>>>(k0, v2, n, hist, ntiles, k1, S, rank);
    k_seg<u32><<<(n + 1 + 255) / 256, 256>>>(k1, n, P.nbuckets, seg);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_S, S, 4ull * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_seg, seg, (P.nbuckets + 2) * 4ull, cudaMemcpyDeviceToHost));
    cudaFree(data); cudaFree(blk); cudaFree(k0); cudaFree(k1); cudaFree(v1); cudaFree(v2); cudaFree(S); cudaFree(rank); cudaFree(hist); cudaFree(tmp); cudaFree(seg);
    return 1;
  }
  u8* data; u16 *keys, *K1, *K2; u32 *V1, *S, *hist, *tmp, *seg;
  const u32 ntiles = (n + RADIX_TILE - 1) / RADIX_TILE;
  CK(cudaMalloc(&data, (size_t)n + 64)); CK(cudaMemset(data, 0, (size_t)n + 64));
  CK(cudaMemcpy(data, h_in, n, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&keys, 2ull * n + 4)); CK(cudaMalloc(&K1, 2ull * n + 4)); CK(cudaMalloc(&K2, 2ull * n + 4));
  CK(cudaMalloc(&V1, 4ull * n)); CK(cudaMalloc(&S, 4ull * n)); CK(cudaMalloc(&hist, 1024ull * ntiles));
  CK(cudaMalloc(&tmp, scan_tmp_words(256ull * ntiles) * 4)); CK(cudaMalloc(&seg, (P.nbuckets + 4) * 4));
  k_hash_keys<<<(n + 255) / 256, 256>>>(P, data, keys);
  k_radix_count<0><<<ntiles, 256>>>(keys, n, hist, ntiles);
  scan_exclusive(hist, 256 * ntiles, tmp, 0);
  k_radix_scatter<0, false><<<ntiles, 256>>>(keys, nullptr, n, hist, ntiles, K1, V1, nullptr);
  k_radix_count<8><<<ntiles, 256>>>(K1, n, hist, ntiles);
  scan_exclusive(hist, 256 * ntiles, tmp, 0);
  k_radix_scatter<8, true><<<ntiles, 256>>>(K1, V1, n, hist, ntiles, K2, S, nullptr);
  k_seg<u16><<<(n + 1 + 255) / 256, 256>>>(K2, n, P.nbuckets, seg);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_S, S, 4ull * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_seg, seg, (P.nbuckets + 2) * 4ull, cudaMemcpyDeviceToHost));
  cudaFree(data); cudaFree(keys); cudaFree(K1); cudaFree(K2); cudaFree(V1); cudaFree(S); cudaFree(hist); cudaFree(tmp); cudaFree(seg);
  return 1;
}
