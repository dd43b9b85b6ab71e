/* brotli/shared_dictionary.h for libbrotlienc_b200: only what the encoder header needs
 * (reference: c/include/brotli/shared_dictionary.h:37-44).  Dictionaries themselves are outside the implemented path. */
#ifndef BROTLI_COMMON_SHARED_DICTIONARY_H_
#define BROTLI_COMMON_SHARED_DICTIONARY_H_
#include <brotli/port.h>
#include <brotli/types.h>
typedef enum BrotliSharedDictionaryType {
  BROTLI_SHARED_DICTIONARY_RAW = 0,        /* LZ77 prefix dictionary */
  BROTLI_SHARED_DICTIONARY_SERIALIZED = 1  /* serialized shared dictionary */
} BrotliSharedDictionaryType;
#endif  /* BROTLI_COMMON_SHARED_DICTIONARY_H_ */
