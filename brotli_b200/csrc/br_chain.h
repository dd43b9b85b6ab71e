// br_chain.h -- the serial glue of one stream: replays what EncodeData (c/enc/encode.c:985)
// does between blocks -- carrying distance cache / pending literals / dictionary counters
// from block to block, ExtendLastCommand eligibility, the merge-or-flush policy
// (encode.c:1141-1166), ShouldCompress (encode.c:457) -- over the walkers' per-block
// summaries.  One warp, O(number of blocks) work.  It also decides which blocks must be
// re-run: a block is dirty when the state it consumed differs from the state the chain now
// derives for it, or when stored-bits it may have consulted were changed by a later commit.
#pragma once
#include "br_cmd.h"

// Compare and commit the stored-bits a walker just produced for block k (warp task).
BR_DEV void br_commit_bits(const BrStream& s, u32 k) {
  const BrBlockIn in = s.bin_used[k];
  u32 w0 = in.pos >> 5, w1 = (in.end - 1) >> 5, diff = 0;
  for (u32 x = w0 + (u32)br_lane(); x <= w1; x += BR_WARP) {
    u32 m = 0xffffffffu;
    if (x == w0) m &= 0xffffffffu << (in.pos & 31);
    if (x == w1) m &= 0xffffffffu >> (31 - ((in.end - 1) & 31));
    u32 nv = s.bits_cur[x] & m, ov = s.bits_latest[x] & m;
    diff += (u32)br_popc(nv ^ ov);
    if (nv != ov) {
      if (m == 0xffffffffu) s.bits_latest[x] = nv;
      else { br_atomic_and(s.bits_latest + x, ~m); br_atomic_or(s.bits_latest + x, nv); }
    }
  }
  diff = br_warp_sum(diff);
  if (br_lane() == 0) {
    s.changed_bits[k] = diff;
    if (diff) s.changed_epoch[k] = (int)s.epoch;
  }
}

// encode.c:457 ShouldCompress
BR_DEV int br_should_compress(const BrStream& s, u32 start, u32 bytes, u32 num_literals, u32 num_commands) {
  if (bytes <= 2) return 0;
  if (num_commands < (bytes >> 8) + 2) {
    if ((double)num_literals > br_dmul(0.99, (double)bytes)) {
      u32* h = s.hist_scratch;
      for (u32 i = (u32)br_lane(); i < 256; i += BR_WARP) h[i] = 0;
      br_syncwarp();
      u32 t = (bytes + 12) / 13;
      for (u32 i = (u32)br_lane(); i < t; i += BR_WARP) br_atomic_add(h + s.data[start + 13 * i], 1);
#if BR_GPU
      __threadfence_block();
#endif
      br_syncwarp();
      const double thr = br_dmul(br_dmul((double)bytes, 7.92), 1.0 / 13.0);
      if (br_bits_entropy(s, h, 256) > thr) return 0;
    }
  }
  return 1;
}

BR_DEV void br_chain(const BrStream& s) {
  const BrParams& P = s.P;
  const int lane = br_lane();
  const u32 nb = P.nblocks, t_now = s.epoch;
  // totals of this launch's commits
  {
    u32 tot = 0;
    for (u32 k = (u32)lane; k < nb; k += BR_WARP)
      if (s.bout[k].valid && s.bout[k].epoch == t_now) tot += s.changed_bits[k];
    tot = br_warp_sum(tot);
    if (lane == 0 && t_now < BR_MAX_EPOCHS) s.epoch_changed[t_now] = tot;
    for (u32 k = (u32)lane; k < nb; k += BR_WARP) s.ext_total[k] = 0;
    br_syncwarp();
  }
  if (lane == 0) {
    u32 acc = 0;
    s.epoch_suffix[BR_MAX_EPOCHS] = 0;
    for (int e = BR_MAX_EPOCHS - 1; e >= 0; --e) {
      if ((u32)e <= t_now) acc += s.epoch_changed[e];
      s.epoch_suffix[e] = acc;
    }
  }
#if BR_GPU
  __threadfence_block();
#endif
  br_syncwarp();

  u32 num_cmds = 0, num_lits = 0, last_insert_len = 0;
  int dc[4] = {4, 11, 15, 16}, saved_dc[4] = {4, 11, 15, 16};
  u32 last_flush_pos = 0, first_block = 0, cmd_total = 0, n_mbs = 0, n_dirty = 0;
  u64 dict_l = 0, dict_m = 0;
  bool have_last = false;
  u32 lc_copy_len = 0, lc_dist_prefix = 0, lc_dist_extra = 0, lc_block = 0;
  const u32 blocksize = 1u << P.lgblock;

  for (u32 k = 0; k < nb; ++k) {
    const u32 pos = s.bin[k].pos, end = s.bin[k].end, is_last = s.bin[k].is_last;
    // ---- state this block must start from
    u32 ext_dist = 0;
    if (num_cmds > 0 && last_insert_len == 0 && have_last) {
      u32 dcode = br_cmd_restore_dcode(lc_dist_prefix, lc_dist_extra);
      int cmd_dist = dc[0];
      if (dcode < 16 || (cmd_dist > 0 && dcode - 15 == (u32)cmd_dist)) {
        u32 lpp = pos - (lc_copy_len & 0x1FFFFFF);
        u32 maxd = br_min(lpp, P.max_backward);
        if (cmd_dist > 0 && (u32)cmd_dist <= maxd) ext_dist = (u32)cmd_dist;
      }
    }
    const BrBlockOut out = s.bout[k];
    u32 dirty = out.valid ? 0u : 1u;  // non-zero: reason code (1 never ran, 2 state, 3 dict gate, 4 window bits, 5 counter wrap)
    if (!dirty) {
      const BrBlockIn u = s.bin_used[k];
      if (u.last_insert_len != last_insert_len || u.ext_dist != ext_dist ||
          u.dc[0] != dc[0] || u.dc[1] != dc[1] || u.dc[2] != dc[2] || u.dc[3] != dc[3]) dirty = 2;
#ifdef BR_SIM_DEBUG
      if (dirty == 2 && getenv("BR_SIM_TRACE2")) fprintf(stderr, "  blk %u: lil %u->%u ext %u->%u dc %d,%d,%d,%d -> %d,%d,%d,%d | out ncmd %u lil %u\n", k, u.last_insert_len, last_insert_len, u.ext_dist, ext_dist, u.dc[0], u.dc[1], u.dc[2], u.dc[3], dc[0], dc[1], dc[2], dc[3], out.ncmd, out.last_insert_len);
#endif
      u64 ul = ((u64)u.dict_l_hi << 32) | u.dict_l_lo, um = ((u64)u.dict_m_hi << 32) | u.dict_m_lo;
      if (!dirty && (ul != dict_l || um != dict_m) && out.gate_checks) {
        bool all_open = out.gate_fail == 0, all_closed = out.gate_fail == out.gate_checks;
        bool ok = (all_open && dict_m >= ((dict_l + out.dl) >> 7)) ||
                  (all_closed && dict_m < (dict_l >> 7));
        if (!ok) dirty = 3;
      }
      if (!dirty) {
        // stored-bits committed at or after this block's last run, inside its window
        u32 lowpos = pos > P.max_backward ? pos - P.max_backward : 0;
        int seen = (int)out.epoch;
        bool hit = false;
        for (u32 base = 0; base < k && !hit; base += BR_WARP) {
          u32 jj = k - 1 - base - (u32)lane;
          bool has = (base + (u32)lane) < k;
          bool inw = has && s.bin[jj].end > lowpos;
          bool ch = inw && s.changed_epoch[jj] >= seen;
          if (br_ballot(ch)) hit = true;
          if (br_ballot(has && !inw) || br_ballot(!has)) break;
        }
        if (hit) dirty = 4;
        u32 unseen = s.epoch_suffix[seen < BR_MAX_EPOCHS ? seen : BR_MAX_EPOCHS];
        if (!dirty && unseen > out.min_wrap_dist) dirty = 5;
      }
    }
    if (lane == 0) {
      BrBlockIn ni = s.bin[k];
      ni.last_insert_len = last_insert_len;
      for (int i = 0; i < 4; ++i) ni.dc[i] = dc[i];
      ni.ext_dist = ext_dist;
      ni.dict_l_lo = (u32)dict_l; ni.dict_l_hi = (u32)(dict_l >> 32);
      ni.dict_m_lo = (u32)dict_m; ni.dict_m_hi = (u32)(dict_m >> 32);
      s.bin[k] = ni;
      s.dirty[k] = dirty;
      s.cmd_off[k] = cmd_total;
    }
    if (dirty) ++n_dirty;
    // ---- carry on with the block's latest (possibly stale) summary
    if (out.valid) {
      if (out.ext_len && have_last) {
        lc_copy_len += out.ext_len;
        if (lane == 0) s.ext_total[lc_block] += out.ext_len;
      }
      if (out.ncmd > 0) {
        const BrCmd c = s.cmd_blocks[(size_t)k * s.cmd_stride + out.ncmd - 1];
        lc_copy_len = c.copy_len; lc_dist_prefix = c.dist_prefix; lc_dist_extra = c.dist_extra;
        lc_block = k; have_last = true;
      }
      num_cmds += out.ncmd; num_lits += out.nlit; last_insert_len = out.last_insert_len;
      for (int i = 0; i < 4; ++i) dc[i] = out.dc[i];
      dict_l += out.dl; dict_m += out.dm;
      cmd_total += out.ncmd;
    } else {
      last_insert_len += end - pos;
    }
    // ---- merge-or-flush (encode.c:1141)
    {
      const u32 processed = end - last_flush_pos;
      const bool next_fits = processed + blocksize <= P.max_mb;
      if (!is_last && !s.bin[k].force_flush && next_fits && num_lits < P.max_mb / 8 &&
          num_cmds < P.max_mb / 8) continue;
    }
    u32 tail = 0;
    if (last_insert_len > 0) {
      tail = last_insert_len; ++num_cmds; num_lits += tail; last_insert_len = 0; ++cmd_total;
    }
    const u32 bytes = end - last_flush_pos;
    int compress = br_should_compress(s, last_flush_pos, bytes, num_lits, num_cmds);
    if (compress && s.force_unc[n_mbs]) compress = 0;
    if (!compress) for (int i = 0; i < 4; ++i) dc[i] = saved_dc[i];
    if (lane == 0) {
      BrMetaBlock m;
      m.start = last_flush_pos; m.end = end; m.first_block = first_block; m.last_block = k;
      m.cmd_off = s.cmd_off[first_block]; m.ncmd = num_cmds; m.nlit = num_lits;
      m.is_last = is_last; m.compress = (u32)compress;
      m.prev_byte = last_flush_pos > 0 ? s.data[last_flush_pos - 1] : 0;
      m.prev_byte2 = last_flush_pos > 1 ? s.data[last_flush_pos - 2] : 0;
      m.pad0 = m.pad1 = 0; m.tail_insert = tail; m.out_bits = 0; m.scratch_off = 0;
      s.mbs[n_mbs] = m;
    }
    br_syncwarp();
    ++n_mbs;
    last_flush_pos = end; num_cmds = 0; num_lits = 0; have_last = false; first_block = k + 1;
    for (int i = 0; i < 4; ++i) saved_dc[i] = dc[i];
  }
  if (lane == 0) { s.counters[0] = n_dirty; s.counters[1] = n_mbs; s.counters[2] = cmd_total; }
}

// Gather one block's commands into the stream-wide compacted array, applying the
// ExtendLastCommand growth (encode.c:961) and the trailing insert-only command
// (encode.c:1169).  Warp task per block.
BR_DEV void br_compact_block(const BrStream& s, u32 k, BrCmd* cmds_all, const u32* block_mb) {
  const BrBlockOut out = s.bout[k];
  const BrCmd* src = s.cmd_blocks + (size_t)k * s.cmd_stride;
  BrCmd* dst = cmds_all + s.cmd_off[k];
  for (u32 i = (u32)br_lane(); i < out.ncmd; i += BR_WARP) {
    BrCmd c = src[i];
    if (i + 1 == out.ncmd && s.ext_total[k]) {
      c.copy_len += s.ext_total[k];
      c.cmd_prefix = br_length_code(c.insert_len,
          (u32)((int)(c.copy_len & 0x1FFFFFF) + (int)(c.copy_len >> 25)),
          (c.dist_prefix & 0x3FF) == 0);
    }
    dst[i] = c;
  }
  const BrMetaBlock mb = s.mbs[block_mb[k]];
  if (mb.last_block == k && mb.tail_insert && br_lane() == 0)
    dst[out.ncmd] = br_init_insert_cmd(mb.tail_insert);
}
