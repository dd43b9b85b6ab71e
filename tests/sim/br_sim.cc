// tests/sim/br_sim.cc -- TEST INFRASTRUCTURE.  Compiles the product's warp-task device code
// (brotli_b200/csrc/br_lz77.h, br_chain.h, br_entropy.h) for the CPU with a one-lane "warp"
// (BR_SIM) so that the bit-exact logic can be debugged against the oracle in a container
// without a GPU.  The data-parallel kernels (hash, sort, scans, packing) have trivial
// sequential stand-ins here; their CUDA versions are checked on the GPU (tests -m gpu).
#define BR_SIM 1
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../../brotli_b200/csrc/br_params.h"
#include "../../brotli_b200/csrc/br_lz77.h"
#include "../../brotli_b200/csrc/br_chain.h"
#ifdef BR_SIM_ENTROPY
#include "../../brotli_b200/csrc/br_entropy.h"
#include "../../brotli_b200/csrc/br_entropy2.h"
#include "../../brotli_b200/csrc/br_assemble.h"
#include "../../brotli_b200/csrc/br_entropy_flat.h"
#endif

struct SimTables {
  std::vector<u8> blob;
  std::vector<double> log2tab;
};
static SimTables g_t;

extern "C" int sim_init(const u8* blob, size_t len, u32 log2n) {
  g_t.blob.assign(blob, blob + len);
  g_t.log2tab.resize(log2n);
  g_t.log2tab[0] = 0;
  for (u32 i = 1; i < log2n; ++i)
    g_t.log2tab[i] = i < 256 ? (double)(float)log2((double)i) : log2((double)i);
  return 1;
}

struct SimStream {
  BrStream s;
  std::vector<u8> data;
  std::vector<u32> S, rank, seg, bits_latest, bits_cur, bits_prev, srch_latest, srch_cur, cover_cur, storedS, prefS, dirty, changed_bits,
      epoch_cum, ext_total, lil_in, cmd_off, force_unc, counters, hist, block_mb;
  std::vector<int> changed_epoch, bitdep_epoch;
  std::vector<u16> skeys, tagS;
  std::vector<BrBlockIn> bin, bin_used;
  std::vector<BrBlockOut> bout;
  std::vector<BrCmd> cmd_blocks, cmds_all;
  std::vector<BrMetaBlock> mbs;
  std::vector<BrBlk> blks;
  std::vector<BrBlkIn> blkin;
  std::vector<u32> key_flips, saw, qpred;
  std::vector<u32> dirty_list, ran_list, slot_blk, stream_blk, stream_nmb, stream_ncmd;
  std::vector<BrMetaBlock> mbs_stage;
  int iterations = 0;
  u64 block_runs = 0;
  double model_cost = 0;
};

static void sim_build_sorted(SimStream& m) {
  BrStream& s = m.s; const BrParams& P = s.P; u32 n = P.n;
  std::vector<u32> key(n);
  u32 hashable = n >= P.htl ? n - P.htl + 1 : 0;  // positions with a full hash load
  for (u32 p = 0; p < n; ++p) key[p] = p >= hashable ? P.nbuckets : P.quick ? br_quick_slot(P, br_ld64u(s.data, p), p - br_stream_base_of(s, p)) : br_hash_key(P, s.data, p);
  m.seg.assign(P.nbuckets + 2, 0);
  for (u32 p = 0; p < n; ++p) m.seg[key[p] + 1]++;
  for (u32 k = 0; k <= P.nbuckets; ++k) m.seg[k + 1] += m.seg[k];
  std::vector<u32> cur(m.seg.begin(), m.seg.end() - 1);
  m.S.resize(n); m.rank.resize(n);
  m.skeys.resize(n + 1);
  for (u32 p = 0; p < n; ++p) { u32 j = cur[key[p]]++; m.S[j] = p; m.rank[p] = j; m.skeys[j] = (u16)key[p]; }
  s.skeys = m.skeys.data();
  m.tagS.resize(n + 1);
  for (u32 j = 0; j < n; ++j) m.tagS[j] = (u16)br_tag4(br_ld32u(s.data, m.S[j]));
  s.tagS = m.tagS.data();
  s.S = m.S.data(); s.rank = m.rank.data(); s.seg = m.seg.data();
}
static void sim_build_storedS(SimStream& m) {
  u32 n = m.s.P.n;
  std::fill(m.storedS.begin(), m.storedS.end(), 0);
  for (u32 j = 0; j < n; ++j) {
    u32 q = m.S[j];
    if ((m.bits_latest[q >> 5] >> (q & 31)) & 1) m.storedS[j >> 5] |= 1u << (j & 31);
  }
  u32 acc = 0;
  for (u32 b = 0; b * 1024 < n + 1024; ++b) {
    m.prefS[b] = acc;
    for (u32 w = b * 32; w < b * 32 + 32 && w < m.storedS.size(); ++w) acc += __builtin_popcount(m.storedS[w]);
  }
}

struct SimCuts { const u32* pos; const u32* kind; u32 n; int is_final; int with_header; int finish_empty; u64* end_bit; u32 size_hint; int lgblock; int disable_ctx; u32 stream_offset; u64* stream_end; };
static SimStream* sim_setup(int q, int lgwin, const u8* in, u32 n, const SimCuts* cuts = nullptr) {
  SimStream* m = new SimStream();
  BrStream& s = m->s;
  memset(&s, 0, sizeof(s));
  if (!br_derive_params(q, lgwin, cuts ? cuts->size_hint : n, n, &s.P, cuts ? cuts->lgblock : 0)) { delete m; return nullptr; }
  s.P.disable_ctx = cuts && cuts->disable_ctx ? 1u : 0u;
  s.P.stream_offset = cuts ? (cuts->stream_offset < s.P.max_backward ? cuts->stream_offset : s.P.max_backward) : 0u;
  BrParams& P = s.P;
  if (getenv("BR_SIM_HEAVY_MIN")) P.heavy_min = (u32)atoi(getenv("BR_SIM_HEAVY_MIN"));
  if (getenv("BR_SIM_STEP_CAP")) P.step_cap = (u32)atoi(getenv("BR_SIM_STEP_CAP"));
  if (getenv("BR_SIM_SWEEP_EPOCH")) P.sweep_epoch = (u32)atoi(getenv("BR_SIM_SWEEP_EPOCH"));
  if (getenv("BR_SIM_SWEEP_BLOCKS")) P.sweep_blocks = (u32)atoi(getenv("BR_SIM_SWEEP_BLOCKS"));
  if (getenv("BR_SIM_PILOT")) P.pilot = (u32)atoi(getenv("BR_SIM_PILOT"));
  if (getenv("BR_SIM_FORCE_EPOCH")) P.force_epoch = (u32)atoi(getenv("BR_SIM_FORCE_EPOCH"));
  { u32 ns = 1; for (u32 i = 0; cuts && i < cuts->n; ++i) if (cuts->kind[i] == 3) ++ns;
    if (ns > 1) { P.pilot = 0; P.multi = ns; P.chunk_bits = getenv("BR_SIM_BATCH_CHUNK_BITS") ? (u32)atoi(getenv("BR_SIM_BATCH_CHUNK_BITS")) : br_batch_chunk_bits(n); } }   // (as br_job_compress_device)
  const u32 ch = 1u << P.chunk_bits;
  std::vector<BrBlockIn> chunks;
  P.finish_empty = cuts && cuts->finish_empty ? 1u : 0u;
  br_build_blocks(P, n, cuts ? cuts->pos : nullptr, cuts ? cuts->n : 0, cuts ? (cuts->is_final != 0 && !cuts->finish_empty) : true, &chunks, m->blks, nullptr, cuts ? cuts->kind : nullptr);
  m->slot_blk.assign(((size_t)n >> P.lgblock) + 2, 0);
  { u32 bb = 0; for (size_t i = 0; i < m->slot_blk.size(); ++i) { const u64 pp = (u64)i << P.lgblock; while (bb + 1 < m->blks.size() && m->blks[bb].end <= pp) ++bb; m->slot_blk[i] = bb; } }
  s.slot_blk = m->slot_blk.data();
  P.nblocks = (u32)chunks.size();
  m->data.assign(in, in + n); m->data.resize(n + 64, 0);
  s.data = m->data.data();
  u32 nb = P.nblocks, words = (n + 31) / 32 + 2;
  m->bin = chunks; m->bin_used.resize(nb); m->bout.resize(nb);
  memset(m->bin_used.data(), 0, nb * sizeof(BrBlockIn));
  memset(m->bout.data(), 0, nb * sizeof(BrBlockOut));
  m->bits_latest.assign(words, 0); m->bits_cur.assign(2 * words, 0); m->srch_latest.assign(words, 0); m->srch_cur.assign(words, 0); m->cover_cur.assign(words, 0);
  // initial guess: everything stored except the unsearchable tail of each block
  for (u32 k = 0; k < nb; ++k)
    for (u32 p = m->bin[k].pos; p < m->bin[k].end && p + P.htl <= m->bin[k].blk_end; ++p) m->bits_latest[p >> 5] |= 1u << (p & 31);
  for (size_t bi = 1; bi < m->blks.size(); ++bi) {   // StitchToPreviousBlock positions
    const BrBlk& B = m->blks[bi];
    if (B.end - B.start >= P.htl - 1 && B.start - B.base >= 3)
      for (u32 q = B.start - 3; q < B.start; ++q) m->bits_latest[q >> 5] |= 1u << (q & 31);
  }
  m->storedS.assign(words + 32, 0); m->prefS.assign(n / 1024 + 4, 0);
  m->dirty.assign(nb, 0); m->changed_bits.assign(nb, 0); m->changed_epoch.assign(nb, -1); m->bitdep_epoch.assign(nb + 1, -1);
  P.max_epochs = 4 * nb + 4096;
  m->epoch_cum.assign(P.max_epochs + 2, 0);
  m->ext_total.assign(nb, 0); m->lil_in.assign(nb, 0); m->cmd_off.assign(nb, 0); m->force_unc.assign(nb + 1, 0);
  m->counters.assign(64, 0); m->hist.assign(256 * (size_t)(P.multi ? P.multi : 1), 0); m->mbs.resize(nb + 1);
  if (P.multi) {
    m->stream_blk.assign(P.multi + 1, (u32)m->blks.size());
    for (size_t bi = m->blks.size(); bi-- > 0;) m->stream_blk[m->blks[bi].stream] = (u32)bi;
    m->stream_nmb.assign(P.multi + 1, 0); m->stream_ncmd.assign(P.multi + 1, 0); m->mbs_stage.resize(m->blks.size());
    s.stream_blk = m->stream_blk.data(); s.stream_nmb = m->stream_nmb.data(); s.stream_ncmd = m->stream_ncmd.data(); s.mbs_stage = m->mbs_stage.data();
  }
  s.cmd_stride = ch / 2 + 2;
  m->cmd_blocks.resize((size_t)nb * s.cmd_stride);
  s.bits_latest = m->bits_latest.data(); s.bits_cur = m->bits_cur.data(); s.bits_words = (u32)m->bits_latest.size();
  s.srch_latest = m->srch_latest.data(); s.srch_cur = m->srch_cur.data(); s.cover_cur = m->cover_cur.data();
  s.storedS = m->storedS.data(); s.prefS = m->prefS.data();
  s.bin = m->bin.data(); s.bin_used = m->bin_used.data(); s.bout = m->bout.data();
  s.cmd_blocks = m->cmd_blocks.data(); s.dirty = m->dirty.data();
  s.changed_bits = m->changed_bits.data(); s.changed_epoch = m->changed_epoch.data(); s.bitdep_epoch = m->bitdep_epoch.data();
  s.epoch_cum = m->epoch_cum.data();
  s.ext_total = m->ext_total.data(); s.lil_in = m->lil_in.data(); s.cmd_off = m->cmd_off.data();
  s.mbs = m->mbs.data(); s.force_unc = m->force_unc.data(); s.counters = m->counters.data();
  s.hist_scratch = m->hist.data();
  s.blk = m->blks.data(); s.nblk = (u32)m->blks.size();
  m->blkin.resize(m->blks.size()); s.blkin = m->blkin.data();
  m->key_flips.assign(P.nbuckets + 2, 0); s.key_flips = m->key_flips.data();
  if (P.quick) { m->saw.assign(((size_t)n << P.qk_sweep_bits) + 4, 0xdeadbeefu); s.saw = m->saw.data(); }
  m->dirty_list.assign(nb + 1, 0); m->block_mb.assign(nb + 1, 0);
  s.dirty_list = m->dirty_list.data(); s.block_mb = m->block_mb.data();
  m->ran_list.assign(nb + 1, 0); s.ran_list = m->ran_list.data();
  const u8* p = g_t.blob.data() + 8;
  s.dict_size_bits = p; p += 32;
  s.dict_offsets = (const u32*)p; p += 128;
  s.dict = p; p += 122784;
  s.dict_hash_words = (const u16*)p; p += 65536;
  s.dict_hash_lengths = p; p += 32768;
  s.ctx_lut = p;
  s.log2tab = g_t.log2tab.data(); s.log2tab_n = (u32)g_t.log2tab.size();
  sim_build_sorted(*m);
  if (P.quick) {   // k_slot_pred
    m->qpred.assign(((size_t)n << P.qk_sweep_bits) + 4, 0);
    for (u32 p = 0; p < n; ++p) br_quick_pred_fill(s, p, br_stream_base_of(s, p), m->qpred.data() + ((size_t)p << P.qk_sweep_bits));
    s.qpred = m->qpred.data();
  }
  return m;
}

static void sim_lz77_fixpoint(SimStream& m) {
  BrStream& s = m.s; u32 nb = s.P.nblocks;
#ifdef BR_SIM_DEBUG
  if (getenv("BR_SIM_WATCH")) br_sim_watch = (u32)atoi(getenv("BR_SIM_WATCH"));
#endif
  // (the launch counter runs on through the rounds of a late fallback, as in br_job_compress_device: runs are stamped with
  // it -- BrBlockOut::epoch -- and the marks of br_commit_bits / br_verify_run are compared with those stamps)
  const u32 round_epoch0 = s.epoch;
  for (;;) {
    br_chain(s);
    if (getenv("BR_SIM_TRACE")) { u32 h[6] = {0}; u32 first = nb; for (u32 k = 0; k < nb; ++k) { h[s.dirty[k] & 7]++; if (s.dirty[k] && first == nb) first = k; }
      fprintf(stderr, "epoch %u dirty %u first %u: never %u state %u dict %u window %u wrap %u | flipped bits %u\n", s.epoch, s.counters[0], first, h[1], h[2], h[3], h[4], h[5], s.epoch_cum[s.epoch] - (s.epoch ? s.epoch_cum[s.epoch - 1] : 0));
#ifdef BR_SIM_DEBUG
      fprintf(stderr, "   last commit: flips %llu successor steps %llu marks %llu overlap marks %llu cap hits %llu\n", (unsigned long long)br_sim_cnt[3],
              (unsigned long long)br_sim_cnt[4], (unsigned long long)br_sim_cnt[0], (unsigned long long)br_sim_cnt[1], (unsigned long long)br_sim_cnt[2]);
      memset(br_sim_cnt, 0, sizeof(br_sim_cnt));
#endif
    }
    if (s.counters[0] == 0) break;
    ++s.epoch; ++m.iterations;
    sim_build_storedS(m);
    std::fill(m.bits_cur.begin(), m.bits_cur.end(), 0); std::fill(m.srch_cur.begin(), m.srch_cur.end(), 0);
    s.counters[4] = 0; s.counters[16] = 0;
    s.forced = s.epoch - round_epoch0 >= s.P.force_epoch;
    { u32 nd = s.counters[5]; std::vector<u32> dl(nd); for (u32 t = 0; t < nd; ++t) dl[t] = br_sched_entry(s, t); for (u32 k : dl) { const bool f = s.forced && k == s.counters[6];
      if (s.P.quick) { if (s.P.multi) br_walk_block<0, true>(s, k, f); else br_walk_block<0>(s, k, f); }
      else if (s.P.multi) { if (s.P.block_bits >= 6) br_walk_block<4, true>(s, k, f); else br_walk_block<1, true>(s, k, f); }
      else if (s.P.block_bits >= 6) br_walk_block<4>(s, k, f); else br_walk_block<1>(s, k, f); } }
    m.block_runs += s.counters[4];
#ifdef BR_SIM_DEBUG
    if (getenv("BR_SIM_TRACE")) { fprintf(stderr, "   work: searches %llu rows %llu groups %llu taken %llu | heavy %llu own-scan rows %llu countS %llu | own_set %llu set_range %llu dict %llu\n",
      (unsigned long long)br_sim_w[0], (unsigned long long)br_sim_w[1], (unsigned long long)br_sim_w[8], (unsigned long long)br_sim_w[9], (unsigned long long)br_sim_w[2], (unsigned long long)br_sim_w[3],
      (unsigned long long)br_sim_w[4], (unsigned long long)br_sim_w[5], (unsigned long long)br_sim_w[6], (unsigned long long)br_sim_w[7]); memset(br_sim_w, 0, sizeof(br_sim_w)); }
#endif
    { // cost model of a launch on the GPU, in chunk-walk times: the longest sweep (serial) or the whole work over ~2000 resident warps
      double c = (double)s.counters[16] > s.counters[4] / 2000.0 ? (double)s.counters[16] : s.counters[4] / 2000.0;
      m.model_cost += c;
      if (getenv("BR_SIM_TRACE")) fprintf(stderr, "   launch %u: sched %u ran %u longest sweep %u  (model cost %.1f, total %.1f)\n", s.epoch, s.counters[5], s.counters[4], s.counters[16], c, m.model_cost);
    }
    m.bits_prev = m.bits_latest; s.bits_prev = m.bits_prev.data();
    std::fill(m.cover_cur.begin(), m.cover_cur.end(), 0);
    for (u32 i = 0; i < s.counters[4]; ++i) br_cover_run(s, s.ran_list[i]);
    for (u32 i = 0; i < s.counters[4]; ++i) br_commit_bits(s, s.ran_list[i]);
    if (s.P.quick) for (u32 k = 0; k < nb; ++k) br_verify_run(s, k);
    if (s.epoch + 2 >= s.P.max_epochs) { fprintf(stderr, "sim: no fixpoint\n"); break; }
  }
  if (getenv("BR_SIM_TRACE")) {   // the static-dictionary gate (hash.h:186) at the fixpoint: chunks that start with it closed
    u32 closed = 0, hits = 0; for (u32 k = 0; k < nb; ++k) { const u64 l = ((u64)s.bin[k].dict_l_hi << 32) | s.bin[k].dict_l_lo, mm = ((u64)s.bin[k].dict_m_hi << 32) | s.bin[k].dict_m_lo; if (mm < (l >> 7)) ++closed; if (s.bout[k].dm) ++hits; }
    fprintf(stderr, "dictionary gate: %u of %u chunks start with the gate closed; %u chunks found dictionary matches\n", closed, nb, hits);
  }
  if (getenv("BR_SIM_VERIFY")) {
    // Is the fixpoint self-consistent?  Re-walk every chunk from its final in-state against the final
    // stored-bits and report the chunks whose bits or out-state come out differently.
    ++s.epoch;
    std::fill(m.bits_cur.begin(), m.bits_cur.end(), 0); std::fill(m.srch_cur.begin(), m.srch_cur.end(), 0);
    std::vector<BrBlockOut> old(s.bout, s.bout + nb);
    std::vector<BrCmd> old_cmds(m.cmd_blocks);
    s.counters[4] = 0;
    for (u32 k = 0; k < nb; ++k) { BrBlockOut o; u32 sp0 = 0xffffffffu; if (s.P.quick && s.P.multi) br_walk_one<0, true>(s, k, s.bin[k], o, k, sp0); else if (s.P.quick) br_walk_one<0>(s, k, s.bin[k], o, k, sp0); else br_walk_one<1>(s, k, s.bin[k], o, k, sp0); }
    m.bits_prev = m.bits_latest; s.bits_prev = m.bits_prev.data();
    std::fill(m.cover_cur.begin(), m.cover_cur.end(), 0);
    for (u32 k = 0; k < nb; ++k) br_cover_run(s, k);
    for (u32 k = 0; k < nb; ++k) {
      br_commit_bits(s, k);
      const BrBlockOut& a = old[k]; const BrBlockOut& b = s.bout[k];
      bool same = a.ncmd == b.ncmd && a.nlit == b.nlit && a.out_pos == b.out_pos && a.last_insert_len == b.last_insert_len &&
                  !memcmp(a.dc, b.dc, sizeof(a.dc)) && a.apply_rh == b.apply_rh;
      if (s.changed_bits[k] && getenv("BR_SIM_VERIFY")[0] == '2') {
        u32 shown = 0;
        for (u32 q = s.bin_used[k].start_pos; q < s.bout[k].out_pos && shown < 12; ++q) {
          u32 o = (m.bits_prev[q >> 5] >> (q & 31)) & 1, n2 = (m.bits_latest[q >> 5] >> (q & 31)) & 1;
          if (o != n2) { fprintf(stderr, "  pos %u: fixpoint %u rewalk %u\n", q, o, n2); ++shown; }
        }
      }
      if (!same || memcmp(&old_cmds[(size_t)k * s.cmd_stride], &m.cmd_blocks[(size_t)k * s.cmd_stride], (size_t)std::min(a.ncmd, b.ncmd) * sizeof(BrCmd))) {
        // first command that differs: where it starts in the input, what it was and what the re-walk says
        u32 pos = s.bin[k].start_pos;
        for (u32 i = 0; i < std::min(a.ncmd, b.ncmd); ++i) {
          const BrCmd& x = old_cmds[(size_t)k * s.cmd_stride + i]; const BrCmd& y = m.cmd_blocks[(size_t)k * s.cmd_stride + i];
          if (memcmp(&x, &y, sizeof(BrCmd))) {
            fprintf(stderr, "verify: chunk %u (run of launch %u) command %u at input %u: fixpoint insert %u copy %u dist_prefix %u extra %u | rewalk insert %u copy %u dist_prefix %u extra %u\n", k, a.epoch, i, pos,
                    x.insert_len, x.copy_len & 0x1FFFFFF, x.dist_prefix, x.dist_extra, y.insert_len, y.copy_len & 0x1FFFFFF, y.dist_prefix, y.dist_extra);
            break;
          }
          pos += x.insert_len + (x.copy_len & 0x1FFFFFF);
        }
      }
      if (s.changed_bits[k] || !same)
        fprintf(stderr, "verify: chunk %u [%u..%u) start %u: %u bits differ, out-state %s (ncmd %u/%u out_pos %u/%u)\n", k, s.bin[k].pos, s.bin[k].end,
                s.bin[k].start_pos, s.changed_bits[k], same ? "same" : "DIFFERENT", a.ncmd, b.ncmd, a.out_pos, b.out_pos);
    }
  }
  m.cmds_all.resize(s.counters[2] + 1);
  for (u32 k = 0; k < nb; ++k) br_compact_block(s, k, m.cmds_all.data(), m.block_mb.data());
}

// Runs the LZ77 stage; returns the commands of every metablock back to back.
extern "C" long sim_lz77(int q, int lgwin, const u8* in, u32 n, BrCmd* cmds_out, u32 cmds_cap,
                         u32* mb_info /* [5*i]: start,end,ncmd,compress,cmd_off */, u32 mb_cap, u32* stats) {
  SimStream* m = sim_setup(q, lgwin, in, n);
  if (!m) return -1;
  sim_lz77_fixpoint(*m);
  u32 nm = m->s.counters[1], total = m->s.counters[2];
  if (total > cmds_cap || nm > mb_cap) { delete m; return -2; }
  memcpy(cmds_out, m->cmds_all.data(), (size_t)total * sizeof(BrCmd));
  for (u32 i = 0; i < nm; ++i) {
    const BrMetaBlock& b = m->s.mbs[i];
    mb_info[5 * i] = b.start; mb_info[5 * i + 1] = b.end; mb_info[5 * i + 2] = b.ncmd;
    mb_info[5 * i + 3] = b.compress; mb_info[5 * i + 4] = b.cmd_off;
  }
  stats[0] = (u32)m->iterations; stats[1] = (u32)m->block_runs; stats[2] = m->s.P.nblocks;
  delete m;
  return nm;
}

// The position index of a one-shot job as the sim builds it (the counterpart of the product's br_debug_sort hook).
extern "C" int sim_debug_sort(int q, int lgwin, const u8* in, u32 n, u32* S_out, u32* seg_out) {
  SimStream* m = sim_setup(q, lgwin, in, n);
  if (!m) return 0;
  memcpy(S_out, m->S.data(), (size_t)n * 4);
  memcpy(seg_out, m->seg.data(), (size_t)(m->s.P.nbuckets + 2) * 4);
  delete m;
  return 1;
}

#ifdef BR_SIM_ENTROPY
static void put_bits_host(std::vector<u8>& o, u64& bit, u32 n, u64 v) {
  for (u32 i = 0; i < n; ++i, ++bit) {
    if ((bit >> 3) >= o.size()) o.resize((bit >> 3) + 1, 0);
    if ((v >> i) & 1) o[bit >> 3] |= (u8)(1u << (bit & 7));
  }
}

// ---- the data-parallel entropy stage (br_entropy2.h), run element by element
struct SimEnt {
  BrEnt e;
  std::vector<u32> lit_ord, cmd_pos, dist_ord, lit_pos, lit_cmd, lit_len, cmd_len, lit_bit_base, cmd_mb, outbits;
  std::vector<u16> dist_sym;
  std::vector<BrMbAux> aux;
  std::vector<u8> scratch;
  std::vector<u64> scratch_off, out_off;
};
static void sim_entropy2(SimStream& m, SimEnt& E) {
  BrStream& s = m.s;
  u32 nm = s.counters[1], C = s.counters[2];
  BrEnt& e = E.e;
  e.cmds = m.cmds_all.data(); e.total_cmds = C;
  E.lit_ord.assign(C + 1, 0); E.cmd_pos.assign(C + 1, 0); E.dist_ord.assign(C + 1, 0);
  for (u32 i = 0; i < C; ++i) {
    u32 a, b, d; br_cmd_scan_inputs(e.cmds[i], &a, &b, &d);
    E.lit_ord[i + 1] = E.lit_ord[i] + a; E.cmd_pos[i + 1] = E.cmd_pos[i] + b; E.dist_ord[i + 1] = E.dist_ord[i] + d;
  }
  e.total_lits = E.lit_ord[C]; e.total_dist = E.dist_ord[C];
  E.lit_pos.assign(e.total_lits + 1, 0); E.lit_cmd.assign(e.total_lits + 1, 0); E.dist_sym.assign(e.total_dist + 1, 0);
  E.lit_len.assign(e.total_lits + 2, 0); E.cmd_len.assign(C + 2, 0); E.lit_bit_base.assign(C + 1, 0); E.cmd_mb.assign(C + 1, 0);
  E.aux.assign(nm, BrMbAux());
  E.scratch_off.assign(nm, 0); E.out_off.assign(nm, 0);
  size_t st = 0, ot = 0;
  for (u32 i = 0; i < nm; ++i) {
    E.scratch_off[i] = st; E.out_off[i] = ot;
    const BrMetaBlock& mb = s.mbs[i];
    for (u32 c = mb.cmd_off; c < mb.cmd_off + mb.ncmd; ++c) E.cmd_mb[c] = i;
    if (mb.compress) { st += (br_mb_scratch_bytes(s.P, mb.nlit, mb.ncmd) + 255) & ~255u; ot += (2 * (size_t)(mb.end - mb.start) + 503) / 4 + 16; }
  }
  E.scratch.assign(st + 256, 0); E.outbits.assign(ot + 64, 0);
  e.lit_ord = E.lit_ord.data(); e.cmd_pos = E.cmd_pos.data(); e.dist_ord = E.dist_ord.data();
  e.lit_pos = E.lit_pos.data(); e.lit_cmd = E.lit_cmd.data(); e.dist_sym = E.dist_sym.data();
  e.lit_len = E.lit_len.data(); e.cmd_len = E.cmd_len.data(); e.lit_bit_base = E.lit_bit_base.data();
  e.cmd_mb = E.cmd_mb.data(); e.aux = E.aux.data(); e.scratch = E.scratch.data(); e.scratch_off = E.scratch_off.data();
  e.outbits = E.outbits.data(); e.out_off = E.out_off.data();
  for (u32 c = 0; c < C; c += BR_WARP) br_expand_cmds(e, c);
  std::vector<u32> smem(64 * 1024);
  for (u32 i = 0; i < nm; ++i) {
    const BrMetaBlock& mb = s.mbs[i];
    if (!mb.compress) continue;
    BrMbAux& a = E.aux[i];
    u8* sc = E.scratch.data() + E.scratch_off[i];
    BrMbMem* M = (BrMbMem*)sc;
    a.which = s.P.disable_ctx ? 1u : br_decide_context_modeling(s, mb.start, mb.end - mb.start, M->sc.rle_syms);
    a.lit_base = E.lit_ord[mb.cmd_off]; a.dist_base = E.dist_ord[mb.cmd_off];
    a.nsym[0] = mb.nlit; a.nsym[1] = mb.ncmd; a.nsym[2] = E.dist_ord[mb.cmd_off + mb.ncmd] - a.dist_base;
    u32 off = br_align8((u32)sizeof(BrMbMem));
    for (int cat = 0; cat < 3; ++cat) { a.var_off[cat] = off; off += br_mb2_var_bytes(br_mb2_nblk(cat, mb.nlit, mb.ncmd)); }
    if (s.P.mb_kind) { br_prep_flat(s, e, mb, a, sc, E.outbits.data() + E.out_off[i]); continue; }
    const u32 A[3] = {256, 704, 64}, NC[3] = {a.which, 1, 1}, MB[3] = {512, 1024, 512};
    const double TH[3] = {400.0, 500.0, 100.0};
    u32* H[3] = {M->lit_H, M->cmd_H, M->dist_H};
    for (int cat = 0; cat < 3; ++cat) {
      u32 nblk = br_mb2_nblk(cat, mb.nlit, mb.ncmd);
      br_split_cta(s, e, mb, a, cat, A[cat], NC[cat], MB[cat], TH[cat], a.nsym[cat], br_mb2_types(sc, a, cat, nblk),
                   br_mb2_lengths(sc, a, cat, nblk), H[cat], smem.data());
    }
    br_prep_codes(s, mb, a, sc, E.outbits.data() + E.out_off[i]);
  }
  for (u32 o = 0; o < e.total_lits; ++o) br_lit_bits(s, e, o);
  for (u32 i = 0; i < C; ++i) br_cmd_bits(s, e, i);
  { u32 acc = 0; for (u32 o = 0; o <= e.total_lits; ++o) { u32 v = E.lit_len[o]; E.lit_len[o] = acc; acc += v; } }
  { u32 acc = 0; for (u32 i = 0; i <= C; ++i) { u32 v = E.cmd_len[i]; E.cmd_len[i] = acc; acc += v; } }
  for (u32 i = 0; i < C; ++i) br_emit_cmd(s, e, i);
  for (u32 o = 0; o < e.total_lits; ++o) br_emit_lit(s, e, o);
}

// Full pipeline; returns compressed size or negative error.  cuts: see br_pipeline.h BrCuts.
static long sim_compress_impl(int q, int lgwin, const u8* in, u32 n, u8* out, size_t out_cap, u32* stats, const SimCuts* cuts) {
  if (n == 0) { if (out_cap < 1) return -3; out[0] = 6; return 1; }
  SimStream* m = sim_setup(q, lgwin, in, n, cuts);
  if (!m) return -1;
  BrStream& s = m->s;
  std::vector<u8> res;
  int rounds = 0;
  for (;;) {
    ++rounds;
    sim_lz77_fixpoint(*m);
    u32 nm = s.counters[1];
    SimEnt E;
    sim_entropy2(*m, E);
    if (getenv("BR_SIM_CHUNKAT")) { const u32 at = (u32)atoi(getenv("BR_SIM_CHUNKAT"));
      for (u32 k = 0; k < s.P.nblocks; ++k) if (s.bin[k].pos <= at + 3000 && s.bin[k].end + 3000 > at)
        fprintf(stderr, "CHUNK %u [%u,%u) blk [%u,%u) start %u dc %d %d %d %d rh %u se %u | out_pos %u ncmd %u dc %d %d %d %d valid %u epoch %u\n", k, s.bin[k].pos, s.bin[k].end, s.bin[k].blk_start, s.bin[k].blk_end,
                s.bin[k].start_pos, s.bin[k].dc[0], s.bin[k].dc[1], s.bin[k].dc[2], s.bin[k].dc[3], s.bin[k].apply_rh, s.bin[k].store_end, s.bout[k].out_pos, s.bout[k].ncmd, s.bout[k].dc[0], s.bout[k].dc[1], s.bout[k].dc[2], s.bout[k].dc[3], s.bout[k].valid, s.bout[k].epoch); }
    if (getenv("BR_SIM_MBS")) {   // the last metablock's commands (debugging against an instrumented reference)
      u32 which = nm - 1;
      if (getenv("BR_DBG_MB")) for (u32 i = 0; i < nm; ++i) if (s.mbs[i].start == (u32)atoi(getenv("BR_DBG_MB"))) which = i;
      const BrMetaBlock& lm = s.mbs[which]; u32 pos = lm.start;
      for (u32 i = 0; i < lm.ncmd; ++i) { const BrCmd& c = m->cmds_all[lm.cmd_off + i];
        fprintf(stderr, "SIMCMD %u ins %u copy %u dprefix %u dextra %u\n", pos, c.insert_len, c.copy_len & 0x1FFFFFF, c.dist_prefix & 0x3FF, c.dist_extra); pos += c.insert_len + (c.copy_len & 0x1FFFFFF); }
    }
    if (getenv("BR_SIM_MBS"))   // metablock table (debugging against an instrumented reference)
      for (u32 i = 0; i < nm; ++i) fprintf(stderr, "SIMMB start %u bytes %u ncmd %u nlit %u last %u compress %u bits %u\n", s.mbs[i].start, s.mbs[i].end - s.mbs[i].start,
                                           s.mbs[i].ncmd, s.mbs[i].nlit, s.mbs[i].is_last, s.mbs[i].compress, s.mbs[i].out_bits);
    // stream assembly: the shared scan (br_assemble.h), then the copies of k_assemble_copy
    std::vector<u32> outw(((size_t)n + ((size_t)n >> 3) + 4096 + 8ull * nm) / 4 + 16, 0);
    std::vector<BrCopyDesc> desc(nm + 1);
    std::vector<u64> cut_end((cuts ? cuts->n : 0) + 2, 0), stream_end(s.P.multi + 2, 0);
    u32 r4[4] = {0, 0, 0, 0};
    br_assemble_scan(s, E.out_off.data(), outw.data(), desc.data(), r4, (!cuts || cuts->with_header) ? 1 : 0, cuts ? cuts->kind : nullptr,
                     cut_end.data(), stream_end.data());
    const bool redo = r4[2] != 0;
    if (!redo) {
      for (u32 i = 0; i < nm; ++i) {
        const BrCopyDesc& d = desc[i];
        if (d.kind == 0) {
          const u32* src = E.outbits.data() + d.src_off;
          for (u32 b = 0; b < d.nbits; ++b) if ((src[b >> 5] >> (b & 31)) & 1) { const u64 db = d.dst_bit + b; outw[db >> 5] |= 1u << (db & 31); }
        } else {
          u8* ob = (u8*)outw.data() + (d.dst_bit >> 3);
          memcpy(ob, in + d.src_off, d.nbits);
        }
      }
      const u64 total = ((u64)r4[1] << 32) | r4[0];
      res.assign((u8*)outw.data(), (u8*)outw.data() + total);
      if (cuts && cuts->end_bit) for (u32 i = 0; i < r4[3]; ++i) cuts->end_bit[i] = cut_end[i];
      if (cuts && cuts->stream_end) for (u32 i = 0; i < s.P.multi; ++i) cuts->stream_end[i] = stream_end[i];
      break;
    }
    if (rounds > 64 + (int)s.P.multi) { delete m; return -4; }
  }
  stats[0] = (u32)m->iterations; stats[1] = (u32)m->block_runs; stats[2] = s.P.nblocks; stats[3] = (u32)rounds; stats[4] = (u32)m->model_cost;
  delete m;
  if (res.size() > out_cap) return -3;
  memcpy(out, res.data(), res.size());
  return (long)res.size();
}
extern "C" long sim_compress(int q, int lgwin, const u8* in, u32 n, u8* out, size_t out_cap, u32* stats) {
  return sim_compress_impl(q, lgwin, in, n, out, out_cap, stats, nullptr);
}
extern "C" long sim_compress_cuts(int q, int lgwin, u32 size_hint, const u8* in, u32 n, const u32* cut_pos, const u32* cut_kind, u32 ncuts,
                                  int is_final, int with_header, int finish_empty, u64* end_bit, u8* out, size_t out_cap, u32* stats,
                                  int lgblock, int disable_ctx, u32 stream_offset) {
  SimCuts c; c.lgblock = lgblock; c.disable_ctx = disable_ctx; c.stream_offset = stream_offset; c.pos = cut_pos; c.kind = cut_kind; c.n = ncuts; c.is_final = is_final; c.with_header = with_header; c.finish_empty = finish_empty; c.end_bit = end_bit; c.size_hint = size_hint; c.stream_end = nullptr;
  return sim_compress_impl(q, lgwin, in, n, out, out_cap, stats, &c);
}
#endif

// ------------------------------------------------------------------ quality 1 (br_q1.h)
#ifdef BR_SIM_ENTROPY
#include "../../brotli_b200/csrc/br_q1_plan.h"
// One stream through the q1 device code with a one-lane warp and a one-thread CTA.
extern "C" long sim_q1_compress_seg(int lgwin, const u8* in, u32 n, const size_t* calls, size_t ncalls, u8* out, size_t out_cap,
                                    int with_header, int end_op, u32 start_bits, u32* end_bit) {
  std::vector<BrQ1Stream> streams; std::vector<BrQ1Frag> frags; std::vector<BrQ1Block> blocks;
  br_q1_plan_stream(lgwin, 0, 0, 0, n, calls, ncalls, streams, frags, blocks, with_header, end_op, start_bits);
  const size_t bound = br_q1_stream_bound(frags, streams[0]);
  std::vector<u8> din((size_t)n + 64, 0); memcpy(din.data(), in, n);
  std::vector<u32> dout(bound / 4 + 16, 0), cmds((size_t)n + 16), hdr(blocks.size() * BR_Q1_HDR_WORDS + 1, 0), counters(16, 0);
  std::vector<u8> lits((size_t)n + 16);
  std::vector<BrQ1Codes> codes(blocks.size() + 1);
  std::vector<int> table((size_t)1 << 17);
  BrQ1 q; memset(&q, 0, sizeof(q));
  q.in = din.data(); q.out = dout.data(); q.cmds = cmds.data(); q.lits = lits.data();
  q.streams = streams.data(); q.frags = frags.data(); q.blocks = blocks.data(); q.codes = codes.data(); q.hdr = hdr.data();
  q.tables = table.data(); q.table_slot = 1u << 17; q.nstreams = 1; q.nfrags = (u32)frags.size(); q.nblocks = (u32)blocks.size();
  q.first_width = 8; q.counters = counters.data(); q.log2tab = g_t.log2tab.data(); q.log2tab_n = (u32)std::min<size_t>(g_t.log2tab.size(), 4096);
  for (u32 f = 0; f < q.nfrags; ++f) {
    if (getenv("BR_SIM_Q1_SHM") && frags[f].size <= 65536) {
      // the on-chip variant (u16 table, input copy aligned like the TMA destination of k_q1_parse_shm)
      std::vector<u16> t16((size_t)1 << 16, 0);
      const u32 a0 = frags[f].start & ~15u;
      std::vector<u8> buf((size_t)65536 + 128, 0);
      size_t nb = (size_t)(frags[f].start - a0) + frags[f].size + 16;
      if (a0 + nb > din.size()) nb = din.size() - a0;
      memcpy(buf.data(), din.data() + a0, nb);
      br_q1_parse_fragment_shm(q, f, t16.data(), buf.data());
    } else br_q1_parse_fragment(q, f, table.data());
  }
  BrQ1Smem* sm = new BrQ1Smem();
  for (u32 b = 0; b < q.nblocks; ++b) br_q1_prep_block(q, b, sm);
  delete sm;
  br_q1_chain_stream(q, 0);
  u32 scratch[8];
  for (u32 b = 0; b < q.nblocks; ++b) br_q1_emit_block(q, b, scratch);
  const size_t sz = streams[0].out_bytes;
  if (end_bit) *end_bit = streams[0].end_bit;
  if (sz > out_cap) return -1;
  memcpy(out, dout.data(), sz);
  return (long)sz;
}
extern "C" long sim_q1_compress(int lgwin, const u8* in, u32 n, const size_t* calls, size_t ncalls, u8* out, size_t out_cap) {
  return sim_q1_compress_seg(lgwin, in, n, calls, ncalls, out, out_cap, 1, 2, 0, nullptr);
}
#endif

#ifdef BR_SIM_ENTROPY
// A batch of independent streams laid end to end (cuts of kind 3, br_pipeline.h): bounds[k] = end of stream k (bounds[ns-1] = n).
// stream_end[k] receives the byte offset where stream k ends in `out`.
extern "C" long sim_compress_multi(int q, int lgwin, const u8* in, u32 n, const u32* bounds, u32 nstreams, u64* stream_end,
                                   u8* out, size_t out_cap, u32* stats) {
  std::vector<u32> kind(nstreams, 3);
  u32 hint = 0;
  for (u32 k = 0; k < nstreams; ++k) { const u32 sz = bounds[k] - (k ? bounds[k - 1] : 0); if (sz > hint) hint = sz; }
  SimCuts c; memset(&c, 0, sizeof(c));
  c.pos = bounds; c.kind = kind.data(); c.n = nstreams - 1; c.is_final = 1; c.with_header = 1; c.size_hint = hint; c.stream_end = stream_end;
  return sim_compress_impl(q, lgwin, in, n, out, out_cap, stats, &c);
}
#endif
