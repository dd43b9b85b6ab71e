// br_cmd.h -- command / prefix-code arithmetic shared by the LZ77 walkers and the entropy
// coder.  Follows RFC 7932 sections 4-5 and the reference's c/enc/command.h, prefix.h
// (cited per function).  NPOSTFIX = NDIRECT = 0 (always, below quality 10 in GENERIC mode:
// c/enc/encode.c:616 ChooseDistanceParams).
#pragma once
#include "br_types.h"

// command.h:31 GetInsertLengthCode
BR_DEV u32 br_ins_code(u32 n) {
  if (n < 6) return n;
  if (n < 130) { u32 nb = br_log2floor(n - 2) - 1u; return (nb << 1) + ((n - 2) >> nb) + 2; }
  if (n < 2114) return br_log2floor(n - 66) + 10;
  if (n < 6210) return 21;
  if (n < 22594) return 22;
  return 23;
}
// command.h:49 GetCopyLengthCode
BR_DEV u32 br_copy_code(u32 n) {
  if (n < 10) return n - 2;
  if (n < 134) { u32 nb = br_log2floor(n - 6) - 1u; return (nb << 1) + ((n - 6) >> nb) + 4; }
  if (n < 2118) return br_log2floor(n - 70) + 12;
  return 23;
}
// RFC 7932 section 5: number of extra bits and base value per insert / copy length code.
// (closed forms of the tables at command.c:15-24)
BR_DEV u32 br_ins_extra(u32 c) {
  return c < 6 ? 0 : c < 16 ? (c - 4) >> 1 : c < 21 ? c - 10 : c == 21 ? 12 : c == 22 ? 14 : 24;
}
BR_DEV u32 br_ins_base(u32 c) {
  if (c < 6) return c;
  if (c < 16) { u32 nb = (c - 4) >> 1; return ((2 + (c & 1)) << nb) + 2; }
  if (c < 22) return (1u << (c - 10)) + 66;
  return c == 22 ? 6210 : 22594;
}
BR_DEV u32 br_copy_extra(u32 c) {
  return c < 8 ? 0 : c < 18 ? (c - 6) >> 1 : c < 23 ? c - 12 : 24;
}
BR_DEV u32 br_copy_base(u32 c) {
  if (c < 8) return c + 2;
  if (c < 18) { u32 nb = (c - 6) >> 1; return ((2 + (c & 1)) << nb) + 6; }
  if (c < 23) return (1u << (c - 12)) + 70;
  return 2118;
}
// command.h:62 CombineLengthCodes
BR_DEV u32 br_combine_codes(u32 ic, u32 cc, int use_last) {
  u32 bits64 = (cc & 7u) | ((ic & 7u) << 3);
  if (use_last && ic < 8u && cc < 16u) return (cc < 8u) ? bits64 : (bits64 | 64u);
  u32 off = 2u * ((cc >> 3) + 3u * (ic >> 3));
  off = (off << 5) + 0x40u + ((0x520D40u >> off) & 0xC0u);
  return off | bits64;
}
BR_DEV u16 br_length_code(u32 ins, u32 copy, int use_last) {
  return (u16)br_combine_codes(br_ins_code(ins), br_copy_code(copy), use_last);
}
// prefix.h:23 PrefixEncodeCopyDistance
BR_DEV void br_prefix_encode_distance(u32 dcode, u16* code, u32* extra) {
  if (dcode < 16) { *code = (u16)dcode; *extra = 0; return; }
  u32 dist = 4 + (dcode - 16);
  u32 bucket = br_log2floor(dist) - 1;
  u32 prefix = (dist >> bucket) & 1;
  u32 offset = (2 + prefix) << bucket;
  *code = (u16)((bucket << 10) | (16 + 2 * (bucket - 1) + prefix));
  *extra = dist - offset;
}
// command.h:120 InitCommand
BR_DEV BrCmd br_init_cmd(u32 ins, u32 copylen, int delta, u32 dcode) {
  BrCmd c;
  u32 d = (u32)(u8)(int8_t)delta;
  c.insert_len = ins;
  c.copy_len = copylen | (d << 25);
  br_prefix_encode_distance(dcode, &c.dist_prefix, &c.dist_extra);
  c.cmd_prefix = br_length_code(ins, (u32)((int)copylen + delta), (c.dist_prefix & 0x3FF) == 0);
  return c;
}
// command.h:138 InitInsertCommand
BR_DEV BrCmd br_init_insert_cmd(u32 ins) {
  BrCmd c;
  c.insert_len = ins;
  c.copy_len = 4u << 25;
  c.dist_extra = 0;
  c.dist_prefix = 16;
  c.cmd_prefix = br_length_code(ins, 4, 0);
  return c;
}
BR_DEV u32 br_cmd_copy_len(const BrCmd& c) { return c.copy_len & 0x1FFFFFF; }
// command.h:176 CommandCopyLenCode
BR_DEV u32 br_cmd_copy_len_code(const BrCmd& c) {
  u32 m = c.copy_len >> 25;
  int delta = (int8_t)(u8)(m | ((m & 0x40) << 1));
  return (u32)((int)(c.copy_len & 0x1FFFFFF) + delta);
}
// command.h:147 CommandRestoreDistanceCode
BR_DEV u32 br_cmd_restore_dcode(u32 dist_prefix, u32 dist_extra) {
  u32 dcode = dist_prefix & 0x3FFu;
  if (dcode < 16) return dcode;
  u32 nbits = dist_prefix >> 10;
  u32 hcode = dcode - 16;
  u32 offset = ((2u + (hcode & 1u)) << nbits) - 4u;
  return offset + dist_extra + 16;
}
// fast_log.h:51 FastLog2 through the host-generated table (see br_host.cc: entries < 256 are
// float-rounded like fast_log.c:13, the rest come from the host's libm log2()).
BR_DEV double br_fast_log2(const BrStream& s, u32 v) {
  return br_ldg(s.log2tab + (v < s.log2tab_n ? v : s.log2tab_n - 1));
}
// bit_cost.c:18 BrotliBitsEntropy -- strictly sequential accumulation.
BR_DEV double br_bits_entropy(const BrStream& s, const u32* pop, u32 size) {
  u32 sum = 0;
  double retval = 0;
  for (u32 i = 0; i < size; ++i) {
    u32 p = pop[i];
    sum += p;
    if (p) retval = br_dsub(retval, br_dmul((double)p, br_fast_log2(s, p)));
  }
  if (sum) retval = br_dadd(retval, br_dmul((double)sum, br_fast_log2(s, sum)));
  if (retval < (double)sum) retval = (double)sum;
  return retval;
}

// Static-dictionary gate (hash.h:186): a search that found nothing probes the dictionary (two
// lookups; ONE for the shallow probe of the quality 2..4 hashers, hash.h:179 -- `per`) only while
// matches >= lookups >> 7; once the gate closes it stays closed.  A chunk
// walked from some counters reports its lookups dl, matches dm and how its gate checks went.
// Given NEW counters (l, m) at its start: is that walk still what the reference would have done,
// and how many lookups / matches does it really add (*edl, *edm)?  Returns 0 if it must be re-walked.
BR_DEV int br_dict_gate_valid(u64 l, u64 m, u32 dl, u32 dm, u32 gate_checks, u32 gate_fail, u32* edl, u32* edm, u32 per = 2) {
  const bool closed = m < (l >> 7);
  *edl = 0; *edm = 0;
  if (gate_checks == 0) return 1;                       // never asked
  if (closed) return dm == 0;                            // no lookups happen; fine if none had succeeded
  if (gate_fail != 0) return 0;                          // some lookups were skipped that would now happen
  if (m >= ((l + dl) >> 7)) { *edl = dl; *edm = dm; return 1; }   // stays open throughout
  if (dm != 0) return 0;                                 // closes somewhere inside and matches are involved
  // no lookup succeeded: the parse is the same; lookups stop once (l >> 7) exceeds m
  const u64 lim = (m + 1) << 7;                          // first l with (l >> 7) > m
  u64 x = lim - l;                                       // > 0 here
  if (per == 2) x = (x + 1) & ~(u64)1;                   // lookups come in pairs (one gate check per search)
  *edl = x < dl ? (u32)x : dl;
  return 1;
}
