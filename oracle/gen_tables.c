/* Dumps the DATA tables of the Brotli format / encoder heuristics that both the
 * oracle and the product need, into one binary blob.  Run once in the
 * build container by oracle/gen_tables.sh; the blob is committed as
 * brotli_b200/data/brotli_tables.bin.  No reference source text is copied:
 * this program links the reference's table objects and writes their bytes.
 *
 * Blob layout (little endian), see brotli_b200/csrc/br_tables.h:
 *   u32 magic 'BRTB', u32 version=1
 *   u8  size_bits_by_length[32]      RFC 7932 Appendix A (NDBITS)
 *   u32 offsets_by_length[32]
 *   u8  dictionary[122784]           RFC 7932 Appendix A
 *   u16 hash_words[32768]            c/enc/dictionary_hash_inc.h:2
 *   u8  hash_lengths[32768]          c/enc/dictionary_hash_inc.h:987
 *   u8  context_lut[2048]            c/common/context.c (RFC 7932 section 7.1)
 */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include "common/dictionary.h"
#include "common/context.h"
#include "enc/dictionary_hash.h"

int main(int argc, char** argv) {
  const BrotliDictionary* d = BrotliGetDictionary();
  FILE* f = fopen(argv[1], "wb");
  uint32_t hdr[2] = {0x42545242u, 1u};
  if (!f) return 1;
  fwrite(hdr, 4, 2, f);
  fwrite(d->size_bits_by_length, 1, 32, f);
  fwrite(d->offsets_by_length, 4, 32, f);
  if (d->data_size != 122784) return 2;
  fwrite(d->data, 1, d->data_size, f);
  fwrite(kStaticDictionaryHashWords, 2, 32768, f);
  fwrite(kStaticDictionaryHashLengths, 1, 32768, f);
  fwrite(_kBrotliContextLookupTable, 1, 2048, f);
  fclose(f);
  return 0;
}
