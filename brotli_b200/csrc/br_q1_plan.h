// br_q1_plan.h -- host-side planning of a quality-1 batch: streams -> fragments -> blocks.
// Shared by br_q1.cu and the CPU sim harness (tests/sim).
#pragma once
#include <string.h>
#include <vector>
#include "br_q1.h"

// encode.c:1425 seen from the caller: each CompressStream call that brings `a` bytes is cut into
// fragments of at most 1 << lgwin bytes; FINISH (the last call) closes the stream behind its last
// fragment, an empty one if that call brought nothing.  calls == nullptr: one call with everything.
// with_header = 0: a segment that continues a stream behind a FLUSH (byte aligned, no window bits);
// end_op = 1: the segment ends with a FLUSH instead of FINISH (no ISLAST, byte padding instead);
// end_op = 0: more input follows (no ISLAST, no padding: the segment may end mid-byte, BrQ1Stream::end_bit);
// start_bits: bits of the previous segment's last, partial byte (the next segment starts behind them).
static inline void br_q1_plan_stream(int lgwin, u32 si, u64 in_off, u64 out_off, size_t n, const size_t* calls, size_t ncalls,
                        std::vector<BrQ1Stream>& streams, std::vector<BrQ1Frag>& frags, std::vector<BrQ1Block>& blocks,
                        int with_header = 1, int end_op = 2, u32 start_bits = 0) {
  const size_t limit = (size_t)1 << lgwin;
  BrQ1Stream s; memset(&s, 0, sizeof(s));
  s.in_off = in_off; s.out_off = out_off; s.size = (u32)n; s.first_frag = (u32)frags.size();
  s.hdr_lgwin = with_header ? (u32)(lgwin < 18 ? 18 : lgwin) : 0u;
  s.flush_end = end_op == 1;
  s.start_bits = with_header ? 0u : (start_bits & 7u);
  size_t one = n;
  if (!calls) { calls = &one; ncalls = 1; }
  size_t pos = 0;
  for (size_t ci = 0; ci < ncalls; ++ci) {
    size_t a = calls[ci];
    const bool finish = ci + 1 == ncalls && end_op == 2;
    if (a == 0 && !finish) continue;            // PROCESS without input compresses nothing
    do {
      const size_t fsz = a < limit ? a : limit;
      BrQ1Frag f; memset(&f, 0, sizeof(f));
      f.stream = si; f.start = (u32)pos; f.size = (u32)fsz; f.first_block = (u32)blocks.size();
      f.is_last = finish && fsz == a;
      size_t ts = 256; u32 tb = 8;
      while (ts < ((size_t)1 << 17) && ts < fsz) { ts <<= 1; ++tb; }   // encode.c:148 HashTableSize
      f.table_bits = tb;
      for (size_t off = 0; off < fsz; off += BR_Q1_BLOCK) {
        BrQ1Block b; memset(&b, 0, sizeof(b));
        b.frag = (u32)frags.size(); b.start = (u32)(pos + off);
        b.size = (u32)(fsz - off < BR_Q1_BLOCK ? fsz - off : BR_Q1_BLOCK);
        blocks.push_back(b); ++f.nblocks;
      }
      frags.push_back(f);
      pos += fsz; a -= fsz;
    } while (a != 0);
  }
  s.nfrags = (u32)frags.size() - s.first_frag;
  streams.push_back(s);
}

// worst-case compressed bytes of a planned stream: every fragment is at most 31 bits above raw (:635)
static inline size_t br_q1_stream_bound(const std::vector<BrQ1Frag>& frags, const BrQ1Stream& s) {
  size_t b = 16;
  for (u32 f = s.first_frag; f < s.first_frag + s.nfrags; ++f) b += frags[f].size + 12;
  return (b + 15) & ~(size_t)15;
}

