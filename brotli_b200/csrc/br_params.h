// br_params.h -- host-side parameter derivation: what SanitizeParams / ComputeLgBlock /
// ChooseHasher / ComputeRbBits / MaxMetablockSize (c/enc/quality.h:59-225) decide for a
// given (quality, lgwin, size_hint).  Host only.
#pragma once
#include "br_types.h"

// Returns 0 if the combination is outside what this library implements (the BrotliEncoder*
// entry points then fail instead of silently doing something else).
static inline int br_derive_params(int quality, int lgwin, u32 size_hint, u32 n, BrParams* P) {
  memset(P, 0, sizeof(*P));
  if (quality < 5 || quality > 9) return 0;   // q0-4: other hashers; q10-11: Zopfli path
  if (lgwin < 17 || lgwin > 24) return 0;      // <=16: forgetful-chain hashers; >24: large window
  P->quality = quality; P->lgwin = lgwin;
  P->lgblock = 16;                              // quality.h:86
  if (quality >= 9 && lgwin > 16) P->lgblock = lgwin < 18 ? lgwin : 18;
  P->hash64 = (size_hint >= (1u << 20) && lgwin >= 19) ? 1 : 0;   // quality.h:182
  P->block_bits = quality - 1;
  P->bucket_bits = P->hash64 ? 15 : (quality < 7 ? 14 : 15);
  P->ndist = quality < 7 ? 4 : quality < 9 ? 10 : 16;
  P->htl = P->hash64 ? 8 : 4;
  int rb = 1 + (lgwin > P->lgblock ? lgwin : P->lgblock);        // quality.h:99
  P->rmask = (1u << rb) - 1;
  P->max_backward = (1u << lgwin) - 16;
  P->spree = quality < 9 ? 64 : 512;                              // quality.h:116
  P->max_mb = 1u << (rb < 24 ? rb : 24);                          // quality.h:103
  P->size_hint = size_hint;
  P->n = n;
  P->nbuckets = 1u << P->bucket_bits;
  P->cpb_shift = (u32)P->lgblock - BR_CHUNK_BITS;
  P->heavy_min = 65536;
  P->step_cap = 4096;
  P->sweep_epoch = 3;    // text / web input settles in 3 launches; what is still dirty then is swept run by run
  P->force_epoch = 64;
  P->sweep_blocks = P->lgblock >= 18 ? 8 : 32;   // sweeps are at most 2 MiB of input long
  return 1;
}
