// br_entropy2.h -- data-parallel formulation of the per-metablock entropy pipeline.
//
// The serial restatement in br_entropy.h (one warp per metablock) stays as the reference the
// sim checks against; this file decomposes the same computation:
//
//   E0  prefix sums over commands      literal ordinal / input position / distance ordinal
//   E1  br_expand_cmd                   literal ordinal -> input position, compact distance symbols
//   E2  br_decide_context_modeling      (br_entropy.h; one warp per metablock)
//   E3  br_split_cta                    greedy block split: the decision CHAIN of
//                                       metablock_inc.h:86 is serial, but each link is executed by a
//                                       whole CTA (histogram accumulation with shared-memory atomics,
//                                       the 3 x contexts entropy sums on separate threads)
//   E4  br_prep_codes                   smoothing + Huffman construction of every histogram in
//                                       parallel, then header / trees stored by one thread
//   E5  br_lit_bits / br_cmd_bits       per-symbol bit counts; prefix sums give every symbol its
//                                       absolute bit offset (the north-star "prefix-sum packing")
//   E6  br_emit_lit / br_emit_cmd       scatter of the code words with atomic OR
#pragma once
#include "br_entropy.h"

#if BR_GPU
#define BR_CTA_TID ((u32)threadIdx.x)
#define BR_CTA_N ((u32)blockDim.x)
BR_DEV void br_cta_sync() { __syncthreads(); }
BR_DEV u32 br_smem_add(u32* p, u32 v) { return atomicAdd(p, v); }
#else
#define BR_CTA_TID 0u
#define BR_CTA_N 1u
BR_DEV void br_cta_sync() {}
BR_DEV u32 br_smem_add(u32* p, u32 v) { u32 o = *p; *p = o + v; return o; }
#endif

// per-metablock results of the parallel stages
struct BrMbAux {
  u32 which;                 // 1, 2, 3 or 13 literal contexts (static map id)
  u32 num_types[3], num_blocks[3];   // literal, command, distance
  u32 nsym[3];
  u32 lit_base, dist_base;   // global ordinals of the first literal / distance symbol
  u32 hdr_bits;
  u32 var_off[3];            // byte offsets of the per-category block arrays inside the scratch
  u32 cmap_size;
};
// per block of a split (stored after E4): where it starts and what precedes its first symbol
struct BrBlockInfo { u32 start; u32 sw_nbits; u64 sw_bits; u32 type; u32 pad; };

BR_HD u32 br_mb2_var_bytes(u32 nblk) { return br_align8(nblk * ((u32)sizeof(BrBlockInfo) + 4) + nblk) + 8; }
#define BR_PREP_THREADS 64
BR_HD u32 br_mb2_scratch_bytes(u32 nlit, u32 ncmd) {
  u32 lb = nlit / 512 + 2, cb = ncmd / 1024 + 2, db = ncmd / 512 + 2;
  return br_align8((u32)sizeof(BrMbMem)) + br_mb2_var_bytes(lb) + br_mb2_var_bytes(cb) + br_mb2_var_bytes(db) +
         BR_PREP_THREADS * (u32)sizeof(BrHTree) * (2 * 704 + 2) + (3 * 256 + 8) * 4;   // + sizes / offsets of the stored codes
}
BR_DEV BrBlockInfo* br_mb2_blocks(u8* scratch, const BrMbAux& a, int cat) { return (BrBlockInfo*)(scratch + a.var_off[cat]); }
BR_DEV u32* br_mb2_lengths(u8* scratch, const BrMbAux& a, int cat, u32 nblk) {
  return (u32*)(scratch + a.var_off[cat] + nblk * (u32)sizeof(BrBlockInfo));
}
BR_DEV u8* br_mb2_types(u8* scratch, const BrMbAux& a, int cat, u32 nblk) {
  return scratch + a.var_off[cat] + nblk * ((u32)sizeof(BrBlockInfo) + 4);
}
BR_HD u32 br_mb2_nblk(int cat, u32 nlit, u32 ncmd) { return cat == 0 ? nlit / 512 + 2 : cat == 1 ? ncmd / 1024 + 2 : ncmd / 512 + 2; }

// Qualities 2 and 3 (BrParams::mb_kind != 0): one prefix code per category, no block splits, no contexts
// (br_entropy_flat.h).  This is the whole per-metablock scratch of such a job.
struct BrMbFlat {
  u32 lit_H[256], cmd_H[704], dist_H[64];
  u8 lit_depth[256], cmd_depth[704], dist_depth[64];
  u16 lit_bits[256], cmd_bits[704], dist_bits[64];
  BrHTree tree[2 * 704 + 2];
  BrTreeSc tsc;
};
BR_HD u32 br_mb_scratch_bytes(const BrParams& P, u32 nlit, u32 ncmd) {
  return P.mb_kind ? br_align8((u32)sizeof(BrMbFlat)) : br_mb2_scratch_bytes(nlit, ncmd);
}

// stream-wide arrays of the parallel entropy stage
struct BrEnt {
  const BrCmd* cmds;       // compacted commands of the stream
  u32 total_cmds, total_lits, total_dist;
  const u32* lit_ord;      // [total_cmds + 1] exclusive scan of insert_len
  const u32* cmd_pos;      // [total_cmds + 1] exclusive scan of insert_len + copy_len (absolute input position)
  const u32* dist_ord;     // [total_cmds + 1] exclusive scan of "has explicit distance"
  u32* lit_pos;            // [total_lits] input position of every literal
  u32* lit_cmd;            // [total_lits] command owning the literal
  u16* dist_sym;           // [total_dist]
  u32* lit_len;            // [total_lits + 1] bit counts, then (in place) their exclusive scan
  u32* cmd_len;            // [total_cmds + 1] likewise for command + distance parts
  u32* lit_bit_base;       // [total_cmds] add the scanned lit_len of a literal to get its bit offset
  u32* cmd_mb;             // [total_cmds] metablock of every command
  BrMbAux* aux;            // [n_mbs]
  u8* scratch; const u64* scratch_off;   // per-metablock scratch
  u32* outbits; const u64* out_off;      // per-metablock bit buffers (u32 words)
};

// largest i in [0, n) with a[i] <= x  (a is non-decreasing, a[0] <= x)
BR_DEV u32 br_upper_index(const u32* a, u32 n, u32 x) {
  u32 lo = 0, hi = n;
  while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (br_ldg(a + mid) <= x) lo = mid; else hi = mid; }
  return lo;
}

// ---------------------------------------------------------------------------------- E0 inputs
BR_DEV void br_cmd_scan_inputs(const BrCmd& c, u32* ins, u32* span, u32* hasd) {
  u32 cl = br_cmd_copy_len(c);
  *ins = c.insert_len; *span = c.insert_len + cl; *hasd = (cl && c.cmd_prefix >= 128) ? 1u : 0u;
}
// ---------------------------------------------------------------------------------- E1
// one warp handles 32 consecutive commands: literal ordinal -> position / owner, distance symbols
BR_DEV void br_expand_cmds(const BrEnt& e, u32 first_cmd) {
  const int lane = br_lane();
  u32 i = first_cmd + (u32)lane;
  u32 ord = 0, pos = 0, len = 0;
  if (i < e.total_cmds) {
    const BrCmd c = e.cmds[i];
    ord = e.lit_ord[i]; pos = e.cmd_pos[i]; len = c.insert_len;
    if (br_cmd_copy_len(c) && c.cmd_prefix >= 128) e.dist_sym[e.dist_ord[i]] = (u16)(c.dist_prefix & 0x3FF);
  }
  u32 cnt = e.total_cmds - first_cmd < BR_WARP ? e.total_cmds - first_cmd : BR_WARP;
  for (u32 t = 0; t < cnt; ++t) {
    u32 o = br_shfl(ord, (int)t), p = br_shfl(pos, (int)t), l = br_shfl(len, (int)t);
    for (u32 k = (u32)lane; k < l; k += BR_WARP) { e.lit_pos[o + k] = p + k; e.lit_cmd[o + k] = first_cmd + t; }
  }
}

// ---------------------------------------------------------------------------------- E3
struct BrSplitState {
  u32 target, block_size, num_blocks, num_types, last_ix[2], merge_last_count, action;
  double last_entropy[26];
};
// symbol `o` (metablock-relative ordinal) of category cat: value and histogram row (context)
BR_DEV void br_symbol_at(const BrStream& st, const BrEnt& e, const BrMetaBlock& mb, const BrMbAux& a, int cat, u32 o,
                         u32* sym, u32* ctx) {
  *ctx = 0;
  if (cat == 0) {
    u32 p = e.lit_pos[a.lit_base + o];
    *sym = st.data[p];
    if (a.which != 1) *ctx = br_static_ctx_map((int)a.which, BR_CTX_UTF8(st, br_data_or_zero(st, p, 1, mb.base), br_data_or_zero(st, p, 2, mb.base)));
  } else if (cat == 1) {
    *sym = e.cmds[mb.cmd_off + o].cmd_prefix;
  } else {
    *sym = e.dist_sym[a.dist_base + o];
  }
}
// shared memory plan (u32 words): cur[nc*A] | comb[2*nc*A] | (8-aligned) terms[3*nc*(A|1)] doubles | ent[3*nc] | state
BR_HD u32 br_split_smem_bytes(u32 A, u32 nc) {
  return (3 * nc * A) * 4 + 8 + (3 * nc * (A | 1) + 3 * nc) * 8 + (u32)sizeof(BrSplitState) + 16;
}
// One CTA = one (metablock, category).  H: [256/nc * nc][A] finished histograms (global).
BR_DEV void br_split_cta(const BrStream& st, const BrEnt& e, const BrMetaBlock& mb, BrMbAux& a, int cat,
                         u32 A, u32 nc, u32 min_block, double thr, u32 nsym,
                         u8* types, u32* lengths, u32* H, u32* smem) {
  const u32 tid = BR_CTA_TID, nt = BR_CTA_N;
  const u32 NA = nc * A, AS = A | 1u, max_types = 256 / nc;
  u32* cur = smem; u32* comb = smem + NA;
  double* terms = (double*)(((uintptr_t)(smem + 3 * NA) + 7) & ~(uintptr_t)7);
  double* ent = terms + 3 * nc * AS;
  BrSplitState* S = (BrSplitState*)(ent + 3 * nc);
  for (u32 x = tid; x < NA; x += nt) cur[x] = 0;
  if (tid == 0) {
    S->target = min_block; S->block_size = 0; S->num_blocks = 0; S->num_types = 0;
    S->last_ix[0] = S->last_ix[1] = 0; S->merge_last_count = 0; S->action = 0;
  }
  br_cta_sync();
  u32 o = 0;
  for (;;) {
    const bool final_call = (o >= nsym);
    u32 take = 0;
    if (!final_call) {
      take = br_min(nsym - o, S->target - S->block_size);
      for (u32 k = tid; k < take; k += nt) {
        u32 sym, ctx;
        br_symbol_at(st, e, mb, a, cat, o + k, &sym, &ctx);
        br_smem_add(cur + ctx * A + sym, 1);
      }
      o += take;
    }
    br_cta_sync();
    u32 bsz = S->block_size + take;
    if (!final_call && bsz != S->target) {   // input exhausted inside a block: the final call follows
      br_cta_sync();
      if (tid == 0) S->block_size = bsz;
      br_cta_sync();
      continue;
    }
    // ---------------- BlockSplitterFinishBlock (metablock_inc.h:86 / metablock.c:524)
    if (bsz < min_block) bsz = min_block;
    const u32 nb_now = S->num_blocks;
    if (nb_now == 0) {
      for (u32 x = tid; x < nc * A; x += nt) { u32 t = x / A, i = x - t * A; u32 p = cur[x]; terms[t * AS + i] = p ? br_dmul((double)p, br_fast_log2(st, p)) : 0.0; }
      br_cta_sync();
      for (u32 t = tid; t < nc; t += nt) {
        double r = 0; u32 sum = 0;
        for (u32 i = 0; i < A; ++i) { double v = terms[t * AS + i]; sum += cur[t * A + i]; if (v != 0.0) r = br_dsub(r, v); }
        if (sum) r = br_dadd(r, br_dmul((double)sum, br_fast_log2(st, sum)));
        if (r < (double)sum) r = (double)sum;
        ent[t] = r;
      }
      br_cta_sync();
      if (tid == 0) {
        lengths[0] = bsz; types[0] = 0;
        for (u32 i = 0; i < nc; ++i) { S->last_entropy[i] = ent[i]; S->last_entropy[nc + i] = ent[i]; }
        S->num_blocks = 1; S->num_types = 1; S->block_size = 0;
      }
      for (u32 x = tid; x < NA; x += nt) { H[x] = cur[x]; cur[x] = 0; }
      br_cta_sync();
    } else if (bsz > 0) {
      const u32 l0 = S->last_ix[0], l1 = S->last_ix[1];
      for (u32 x = tid; x < 2 * NA; x += nt) {
        u32 j = x / NA, r = x - j * NA;
        comb[x] = cur[r] + H[(j ? l1 : l0) * A + r];
      }
      br_cta_sync();
      for (u32 x = tid; x < 3 * NA; x += nt) {
        u32 t = x / A, i = x - t * A;
        u32 p = t < nc ? cur[x] : comb[x - NA];
        terms[t * AS + i] = p ? br_dmul((double)p, br_fast_log2(st, p)) : 0.0;
      }
      br_cta_sync();
      for (u32 t = tid; t < 3 * nc; t += nt) {
        const u32* src = t < nc ? cur + t * A : comb + (t - nc) * A;
        double r = 0; u32 sum = 0;
        for (u32 i = 0; i < A; ++i) { double v = terms[t * AS + i]; sum += src[i]; if (v != 0.0) r = br_dsub(r, v); }
        if (sum) r = br_dadd(r, br_dmul((double)sum, br_fast_log2(st, sum)));
        if (r < (double)sum) r = (double)sum;
        ent[t] = r;
      }
      br_cta_sync();
      if (tid == 0) {
        double diff[2] = {0.0, 0.0};
        for (u32 i = 0; i < nc; ++i)
          for (u32 j = 0; j < 2; ++j) {
            u32 jx = j * nc + i;
            diff[j] = br_dadd(diff[j], br_dsub(br_dsub(ent[nc + jx], ent[i]), S->last_entropy[jx]));
          }
        u32 nbk = S->num_blocks;
        if (S->num_types < max_types && diff[0] > thr && diff[1] > thr) {
          lengths[nbk] = bsz; types[nbk] = (u8)S->num_types;
          S->last_ix[1] = S->last_ix[0]; S->last_ix[0] = S->num_types * nc;
          for (u32 i = 0; i < nc; ++i) { S->last_entropy[nc + i] = S->last_entropy[i]; S->last_entropy[i] = ent[i]; }
          ++S->num_blocks; ++S->num_types;
          S->merge_last_count = 0; S->target = min_block; S->action = 1;
        } else if (diff[1] < br_dsub(diff[0], 20.0)) {
          lengths[nbk] = bsz; types[nbk] = types[nbk - 2];
          u32 t = S->last_ix[0]; S->last_ix[0] = S->last_ix[1]; S->last_ix[1] = t;
          for (u32 i = 0; i < nc; ++i) { S->last_entropy[nc + i] = S->last_entropy[i]; S->last_entropy[i] = ent[2 * nc + i]; }
          ++S->num_blocks;
          S->merge_last_count = 0; S->target = min_block; S->action = 2;
        } else {
          lengths[nbk - 1] += bsz;
          for (u32 i = 0; i < nc; ++i) {
            S->last_entropy[i] = ent[nc + i];
            if (S->num_types == 1) S->last_entropy[nc + i] = S->last_entropy[i];
          }
          if (++S->merge_last_count > 1) S->target += min_block;
          S->action = 3;
        }
        S->block_size = 0;
      }
      br_cta_sync();
      {
        const u32 act = S->action, dst = S->last_ix[0];
        for (u32 x = tid; x < NA; x += nt) {
          H[dst * A + x] = act == 1 ? cur[x] : act == 2 ? comb[NA + x] : comb[x];
          cur[x] = 0;
        }
      }
      br_cta_sync();
    }
    if (final_call) break;
  }
  if (tid == 0) { a.num_types[cat] = S->num_types; a.num_blocks[cat] = S->num_blocks; }
}

// ---------------------------------------------------------------------------------- E4
// CTA per metablock: codes of all histograms in parallel, then warp 0 writes the metablock
// header (brotli_bit_stream.c:947-1060) and the block tables used by E5/E6.
BR_DEV void br_prep_codes(const BrStream& st, const BrMetaBlock& mb, BrMbAux& a, u8* scratch, u32* out) {
  const u32 tid = BR_CTA_TID, nt = BR_CTA_N;
  BrMbMem* M = (BrMbMem*)scratch;
  const u32 nctx = a.which;
  const u32 nl = a.num_types[0] * nctx, ncm = a.num_types[1], nd = a.num_types[2];
  BrHTree* trees = (BrHTree*)(scratch + br_mb2_scratch_bytes(mb.nlit, mb.ncmd) - BR_PREP_THREADS * (u32)sizeof(BrHTree) * (2 * 704 + 2) - (3 * 256 + 8) * 4);
  BrHTree* mytree = trees + (size_t)(tid % BR_PREP_THREADS) * (2 * 704 + 2);
  // Every prefix code of the metablock is built by its own thread, which also measures how many bits its serialised form
  // takes (brotli_bit_stream.c:349: the store, run against a counting writer).  Once the header in front of the codes is
  // written the offsets are known and the same threads store their codes in parallel: all writes are atomic ORs.
  u32* tree_bits = (u32*)(trees + (size_t)BR_PREP_THREADS * (2 * 704 + 2));   // [nl + ncm + nd] sizes, then bit offsets
  const u32 ntrees = nl + ncm + nd;
  const u32 stride = nt < BR_PREP_THREADS ? nt : BR_PREP_THREADS;
  if (tid < BR_PREP_THREADS) {
    for (u32 t = tid; t < ntrees; t += stride) {
      u8 good[704];
      const u32* H; u8* dep; u32 len;
      if (t < nl) {
        H = M->lit_H + t * 256; dep = M->lit_depth + t * 256; len = 256;
        br_optimize_counts_for_rle(256, M->lit_H + t * 256, good);
        br_build_tree(H, 256, mytree, dep, M->lit_bits + t * 256);
      } else if (t < nl + ncm) {
        u32 x = t - nl;
        H = M->cmd_H + x * 704; dep = M->cmd_depth + x * 704; len = 704;
        br_optimize_counts_for_rle(704, M->cmd_H + x * 704, good);
        br_build_tree(H, 704, mytree, dep, M->cmd_bits + x * 704);
      } else {
        u32 x = t - nl - ncm;
        H = M->dist_H + x * 64; dep = M->dist_depth + x * 64; len = 64;
        br_optimize_counts_for_rle(64, M->dist_H + x * 64, good);
        br_build_tree(H, 64, mytree, dep, M->dist_bits + x * 64);
      }
      BrBitW cw; cw.out = nullptr; cw.ix = 0; cw.per_thread = 1;
      br_store_tree(H, len, len, (BrTreeSc*)mytree, dep, cw);   // (the big tree is done: its memory is the store's scratch)
      tree_bits[t] = cw.ix;
    }
  }
  // metablock.c:677 MapStaticContexts
  if (nctx > 1)
    for (u32 x = tid; x < (a.num_types[0] << 6); x += nt) M->cmap[x] = (x >> 6) * nctx + br_static_ctx_map((int)nctx, x & 63);
#if BR_GPU
  __threadfence_block();
#endif
  br_cta_sync();
  const int lane = br_lane();
  BrMbScratch* sc = &M->sc;
  const u32 cmap_size = nctx > 1 ? (a.num_types[0] << 6) : 0;
  BrBitW w; w.out = out; w.ix = 0; w.per_thread = 0;
  if (tid < BR_WARP) {
  // ---- warp 0: header in front of the codes, warp-uniform
  br_put_bits(w, 1, (u64)mb.is_last);
  if (mb.is_last) br_put_bits(w, 1, 0);
  br_store_mlen(mb.end - mb.start, w);
  if (!mb.is_last) br_put_bits(w, 1, 0);
  for (int cat = 0; cat < 3; ++cat) {
    const u32 hist_len = cat == 0 ? 256u : cat == 1 ? 704u : 64u;
    u32 nblk = br_mb2_nblk(cat, mb.nlit, mb.ncmd);
    BrSplit sp; sp.num_types = a.num_types[cat]; sp.num_blocks = a.num_blocks[cat];
    sp.types = br_mb2_types(scratch, a, cat, nblk); sp.lengths = br_mb2_lengths(scratch, a, cat, nblk);
    BrBlockEnc be;
    br_block_enc_init(be, hist_len, sp, &M->code[cat]);
    br_build_and_store_block_split_code(be, sc, w);
  }
  br_put_bits(w, 2, 0);
  br_put_bits(w, 4, 0);
  for (u32 i = 0; i < a.num_types[0]; ++i) br_put_bits(w, 2, 2);
  if (cmap_size == 0) br_store_trivial_context_map(nl, 6, sc, w);
  else br_encode_context_map(M->cmap, cmap_size, nl, sc, w);
  br_store_trivial_context_map(nd, 2, sc, w);
  if (lane == 0) {   // sizes -> bit offsets
    u32 acc = w.ix;
    for (u32 t = 0; t < ntrees; ++t) { const u32 v = tree_bits[t]; tree_bits[t] = acc; acc += v; }
    tree_bits[ntrees] = acc;
  }
  }
#if BR_GPU
  __threadfence_block();
#endif
  br_cta_sync();
  if (tid < BR_PREP_THREADS) {
    for (u32 t = tid; t < ntrees; t += stride) {
      BrBitW tw; tw.out = out; tw.ix = tree_bits[t]; tw.per_thread = 1;
      if (t < nl) br_store_tree(M->lit_H + t * 256, 256, 256, (BrTreeSc*)mytree, M->lit_depth + t * 256, tw);
      else if (t < nl + ncm) { const u32 x = t - nl; br_store_tree(M->cmd_H + x * 704, 704, 704, (BrTreeSc*)mytree, M->cmd_depth + x * 704, tw); }
      else { const u32 x = t - nl - ncm; br_store_tree(M->dist_H + x * 64, 64, 64, (BrTreeSc*)mytree, M->dist_depth + x * 64, tw); }
    }
  }
  if (tid >= BR_WARP) return;
  w.ix = tree_bits[ntrees];
  // ---- block tables: start ordinal of every block and the block-switch bits that precede it
  for (int cat = 0; cat < 3; ++cat) {
    u32 nblk = br_mb2_nblk(cat, mb.nlit, mb.ncmd);
    const u8* types = br_mb2_types(scratch, a, cat, nblk);
    const u32* lengths = br_mb2_lengths(scratch, a, cat, nblk);
    BrBlockInfo* bi = br_mb2_blocks(scratch, a, cat);
    const BrBlockCode* code = &M->code[cat];
    const u32 nb = a.num_blocks[cat];
    if (lane == 0) { u32 acc = 0; for (u32 b = 0; b < nb; ++b) { bi[b].start = acc; acc += lengths[b]; } }
    for (u32 b = (u32)lane; b < nb; b += BR_WARP) {
      u32 type = types[b];
      u64 bits = 0; u32 n = 0;
      if (b > 0 && a.num_types[cat] > 1) {
        // brotli_bit_stream.c:58 NextBlockTypeCode: last = types[b-1], second last = types[b-2]
        // (the calculator starts at last = 1, second last = 0 and block 0 is fed first)
        u32 prev1 = types[b - 1], prev2 = b >= 2 ? types[b - 2] : 1u;
        u32 tc = (type == prev1 + 1) ? 1u : (type == prev2) ? 0u : type + 2u;
        u32 lc = br_block_len_code(lengths[b]);
        bits = code->type_bits[tc]; n = code->type_depths[tc];
        bits |= (u64)code->len_bits[lc] << n; n += code->len_depths[lc];
        bits |= (u64)(lengths[b] - br_block_len_offset(lc)) << n; n += br_block_len_nbits(lc);
      }
      bi[b].type = type; bi[b].sw_nbits = n; bi[b].sw_bits = bits; bi[b].pad = 0;
    }
  }
  br_syncwarp();
  if (lane == 0) { a.hdr_bits = w.ix; a.cmap_size = cmap_size; }
}

// ---------------------------------------------------------------------------------- E5 / E6
BR_DEV u32 br_block_of(const BrBlockInfo* bi, u32 nb, u32 o) {
  u32 lo = 0, hi = nb;
  while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (bi[mid].start <= o) lo = mid; else hi = mid; }
  return lo;
}
// code word of literal `o` (global ordinal): returns its length; *sw / *swn: block switch before it
BR_DEV u32 br_lit_code(const BrStream& st, const BrEnt& e, u32 o, u32* code, u64* sw, u32* swn, u32* mb_out) {
  const u32 cmd = e.lit_cmd[o], m = e.cmd_mb[cmd];
  *mb_out = m; *swn = 0; *sw = 0; *code = 0;
  if (!st.mbs[m].compress) return 0;
  const BrMbAux& a = e.aux[m];
  u8* scratch = e.scratch + e.scratch_off[m];
  if (st.P.mb_kind) {   // qualities 2, 3: one literal code
    const BrMbFlat* F = (const BrMbFlat*)scratch;
    const u32 lit = st.data[e.lit_pos[o]];
    *code = F->lit_bits[lit];
    return F->lit_depth[lit];
  }
  const BrMbMem* M = (const BrMbMem*)scratch;
  const BrBlockInfo* bi = br_mb2_blocks(scratch, a, 0);
  const u32 orel = o - a.lit_base;
  const u32 b = br_block_of(bi, a.num_blocks[0], orel);
  const u32 p = e.lit_pos[o], lit = st.data[p], type = bi[b].type;
  u32 hix;
  const u32 base = st.mbs[m].base;
  if (a.cmap_size) hix = M->cmap[(type << 6) + BR_CTX_UTF8(st, br_data_or_zero(st, p, 1, base), br_data_or_zero(st, p, 2, base))] * 256 + lit;
  else hix = type * 256 + lit;
  if (bi[b].start == orel) { *swn = bi[b].sw_nbits; *sw = bi[b].sw_bits; }
  *code = M->lit_bits[hix];
  return M->lit_depth[hix];
}
BR_DEV void br_lit_bits(const BrStream& st, const BrEnt& e, u32 o) {
  u32 code, swn, m; u64 sw;
  u32 n = br_lit_code(st, e, o, &code, &sw, &swn, &m);
  e.lit_len[o] = n + swn;
}
// The seven pieces a command contributes, in stream order (brotli_bit_stream.c:1062-1106):
// [cmd block switch] cmd symbol, insert/copy extra bits, <literals>, [dist block switch] dist symbol, dist extra
struct BrCmdCode { u64 csw, extra, dsw; u32 cswn, csym, cn, en, dswn, dsym, dn, xbits, xn; };
BR_DEV void br_cmd_code(const BrStream& st, const BrEnt& e, u32 i, BrCmdCode* k) {
  const u32 m = e.cmd_mb[i];
  const BrMetaBlock& mb = st.mbs[m];
  k->csw = k->extra = k->dsw = 0;
  k->cswn = k->csym = k->cn = k->en = k->dswn = k->dsym = k->dn = k->xbits = k->xn = 0;
  if (!mb.compress) return;
  const BrMbAux& a = e.aux[m];
  u8* scratch = e.scratch + e.scratch_off[m];
  const BrMbMem* M = (const BrMbMem*)scratch;
  const BrCmd c = e.cmds[i];
  if (st.P.mb_kind) {   // qualities 2, 3: one command code, one distance code (brotli_bit_stream.c:1159 StoreDataWithHuffmanCodes)
    const BrMbFlat* F = (const BrMbFlat*)scratch;
    k->cn = F->cmd_depth[c.cmd_prefix]; k->csym = F->cmd_bits[c.cmd_prefix];
    const u32 clc = br_cmd_copy_len_code(c);
    const u32 ic = br_ins_code(c.insert_len), cc = br_copy_code(clc);
    const u32 insn = br_ins_extra(ic);
    const u64 insv = c.insert_len - br_ins_base(ic), copyv = clc - br_copy_base(cc);
    k->extra = (copyv << insn) | insv; k->en = insn + br_copy_extra(cc);
    if (br_cmd_copy_len(c) && c.cmd_prefix >= 128) {
      const u32 ds = c.dist_prefix & 0x3FF;
      k->dn = F->dist_depth[ds]; k->dsym = F->dist_bits[ds];
      k->xn = c.dist_prefix >> 10; k->xbits = c.dist_extra;
    }
    return;
  }
  {
    const BrBlockInfo* bi = br_mb2_blocks(scratch, a, 1);
    const u32 orel = i - mb.cmd_off, b = br_block_of(bi, a.num_blocks[1], orel);
    if (bi[b].start == orel) { k->cswn = bi[b].sw_nbits; k->csw = bi[b].sw_bits; }
    const u32 ix = bi[b].type * 704 + c.cmd_prefix;
    k->cn = M->cmd_depth[ix]; k->csym = M->cmd_bits[ix];
    // brotli_bit_stream.c:82 StoreCommandExtra
    u32 clc = br_cmd_copy_len_code(c);
    u32 ic = br_ins_code(c.insert_len), cc = br_copy_code(clc);
    u32 insn = br_ins_extra(ic);
    u64 insv = c.insert_len - br_ins_base(ic), copyv = clc - br_copy_base(cc);
    k->extra = (copyv << insn) | insv; k->en = insn + br_copy_extra(cc);
  }
  if (br_cmd_copy_len(c) && c.cmd_prefix >= 128) {
    const BrBlockInfo* bi = br_mb2_blocks(scratch, a, 2);
    const u32 orel = e.dist_ord[i] - a.dist_base, b = br_block_of(bi, a.num_blocks[2], orel);
    if (bi[b].start == orel) { k->dswn = bi[b].sw_nbits; k->dsw = bi[b].sw_bits; }
    const u32 ix = bi[b].type * 64 + (c.dist_prefix & 0x3FF);
    k->dn = M->dist_depth[ix]; k->dsym = M->dist_bits[ix];
    k->xn = c.dist_prefix >> 10; k->xbits = c.dist_extra;
  }
}
BR_DEV void br_cmd_bits(const BrStream& st, const BrEnt& e, u32 i) {
  BrCmdCode k;
  br_cmd_code(st, e, i, &k);
  e.cmd_len[i] = k.cswn + k.cn + k.en + k.dswn + k.dn + k.xn;
}
// after the exclusive scans of lit_len / cmd_len: place the command's own pieces
BR_DEV void br_emit_cmd(const BrStream& st, const BrEnt& e, u32 i) {
  const u32 m = e.cmd_mb[i];
  const BrMetaBlock& mb = st.mbs[m];
  if (!mb.compress) return;
  const BrMbAux& a = e.aux[m];
  BrCmdCode k;
  br_cmd_code(st, e, i, &k);
  u32* out = e.outbits + e.out_off[m];
  const u32 lo = e.lit_ord[i], ins = e.cmds[i].insert_len;
  u32 off = a.hdr_bits + (e.cmd_len[i] - e.cmd_len[mb.cmd_off]) + (e.lit_len[lo] - e.lit_len[a.lit_base]);
  br_put_bits_at(out, off, k.cswn, k.csw); off += k.cswn;
  br_put_bits_at(out, off, k.cn, k.csym); off += k.cn;
  br_put_bits_at(out, off, k.en, k.extra); off += k.en;
  e.lit_bit_base[i] = off - e.lit_len[lo];
  off += e.lit_len[lo + ins] - e.lit_len[lo];
  br_put_bits_at(out, off, k.dswn, k.dsw); off += k.dswn;
  br_put_bits_at(out, off, k.dn, k.dsym); off += k.dn;
  br_put_bits_at(out, off, k.xn, k.xbits); off += k.xn;
  if (i + 1 == mb.cmd_off + mb.ncmd) st.mbs[m].out_bits = off;
}
BR_DEV void br_emit_lit(const BrStream& st, const BrEnt& e, u32 o) {
  u32 code, swn, m; u64 sw;
  u32 n = br_lit_code(st, e, o, &code, &sw, &swn, &m);
  if (!st.mbs[m].compress) return;
  u32* out = e.outbits + e.out_off[m];
  u32 off = e.lit_bit_base[e.lit_cmd[o]] + e.lit_len[o];
  br_put_bits_at(out, off, swn, sw);
  br_put_bits_at(out, off + swn, n, code);
}
