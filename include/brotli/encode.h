/* brotli/encode.h for libbrotlienc_b200 -- the encoder API of google/brotli (c/include/brotli/encode.h), declared for
 * this library: identical names, enum values, argument meaning and error behaviour, so sources written against the
 * reference header compile against this one and binaries linked against libbrotlienc.so.1 run against this library.
 * Each declaration cites the reference line it mirrors.  What the GPU path implements is stated in brotli_b200.h;
 * everything outside it fails (BROTLI_FALSE / NULL) instead of producing other bytes. */
#ifndef BROTLI_ENC_ENCODE_H_
#define BROTLI_ENC_ENCODE_H_
#include <brotli/port.h>
#include <brotli/shared_dictionary.h>
#include <brotli/types.h>
#if defined(__cplusplus) || defined(c_plusplus)
extern "C" {
#endif

#define BROTLI_MIN_WINDOW_BITS 10          /* encode.h:24 */
#define BROTLI_MAX_WINDOW_BITS 24          /* encode.h:30 */
#define BROTLI_LARGE_MAX_WINDOW_BITS 30    /* encode.h:35 */
#define BROTLI_MIN_INPUT_BLOCK_BITS 16     /* encode.h:37 */
#define BROTLI_MAX_INPUT_BLOCK_BITS 24     /* encode.h:39 */
#define BROTLI_MIN_QUALITY 0               /* encode.h:41 */
#define BROTLI_MAX_QUALITY 11              /* encode.h:43 */

typedef enum BrotliEncoderMode {           /* encode.h:45 */
  BROTLI_MODE_GENERIC = 0, BROTLI_MODE_TEXT = 1, BROTLI_MODE_FONT = 2
} BrotliEncoderMode;
typedef enum BrotliEncoderBase64Mode {     /* encode.h:61 */
  BROTLI_BASE64_MODE_DISABLED = 0, BROTLI_BASE64_MODE_ENABLED = 1
} BrotliEncoderBase64Mode;
#define BROTLI_DEFAULT_BASE64_MODE BROTLI_BASE64_MODE_DISABLED
#define BROTLI_DEFAULT_MAX_BASE64_REGIONS 16
typedef enum BrotliEncoderSimdHasher {     /* encode.h:74 */
  BROTLI_SIMD_HASHER_DEFAULT = 0, BROTLI_SIMD_HASHER_ENABLED = 1, BROTLI_SIMD_HASHER_DISABLED = 2
} BrotliEncoderSimdHasher;
#define BROTLI_DEFAULT_SIMD_HASHER BROTLI_SIMD_HASHER_DEFAULT
#define BROTLI_DEFAULT_QUALITY 11
#define BROTLI_DEFAULT_WINDOW 22
#define BROTLI_DEFAULT_MODE BROTLI_MODE_GENERIC

typedef enum BrotliEncoderOperation {      /* encode.h:93 */
  BROTLI_OPERATION_PROCESS = 0, BROTLI_OPERATION_FLUSH = 1,
  BROTLI_OPERATION_FINISH = 2, BROTLI_OPERATION_EMIT_METADATA = 3
} BrotliEncoderOperation;
typedef enum BrotliEncoderParameter {      /* encode.h:160 */
  BROTLI_PARAM_MODE = 0, BROTLI_PARAM_QUALITY = 1, BROTLI_PARAM_LGWIN = 2,
  BROTLI_PARAM_LGBLOCK = 3, BROTLI_PARAM_DISABLE_LITERAL_CONTEXT_MODELING = 4,
  BROTLI_PARAM_SIZE_HINT = 5, BROTLI_PARAM_LARGE_WINDOW = 6, BROTLI_PARAM_NPOSTFIX = 7,
  BROTLI_PARAM_NDIRECT = 8, BROTLI_PARAM_STREAM_OFFSET = 9, BROTLI_PARAM_BASE64_MODE = 10,
  BROTLI_PARAM_MAX_BASE64_REGIONS = 11, BROTLI_PARAM_SIMD_HASHER = 12
} BrotliEncoderParameter;

typedef struct BrotliEncoderStateStruct BrotliEncoderState;                              /* encode.h:271 */
typedef struct BrotliEncoderPreparedDictionaryStruct BrotliEncoderPreparedDictionary;    /* encode.h:317 */

/* encode.h:289 -- fails once the first CompressStream call has been made (c/enc/encode.c:63) */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderSetParameter(BrotliEncoderState* state, BrotliEncoderParameter param, uint32_t value);
/* encode.h:306 -- alloc_func and free_func both NULL or both set; state and buffers come from them */
BROTLI_ENC_API BrotliEncoderState* BrotliEncoderCreateInstance(brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque);
/* encode.h:314 */
BROTLI_ENC_API void BrotliEncoderDestroyInstance(BrotliEncoderState* state);
/* encode.h:343 -- custom dictionaries are outside the implemented path: returns NULL */
BROTLI_ENC_API BrotliEncoderPreparedDictionary* BrotliEncoderPrepareDictionary(BrotliSharedDictionaryType type, size_t data_size,
    const uint8_t data[BROTLI_ARRAY_PARAM(data_size)], int quality, brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque);
/* encode.h:348 */
BROTLI_ENC_API void BrotliEncoderDestroyPreparedDictionary(BrotliEncoderPreparedDictionary* dictionary);
/* encode.h:361 -- returns BROTLI_FALSE */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderAttachPreparedDictionary(BrotliEncoderState* state, const BrotliEncoderPreparedDictionary* dictionary);
/* encode.h:375 */
BROTLI_ENC_API size_t BrotliEncoderMaxCompressedSize(size_t input_size);
/* encode.h:405 -- one-shot; bytes equal the reference's for the same arguments */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderCompress(int quality, int lgwin, BrotliEncoderMode mode, size_t input_size,
    const uint8_t input_buffer[BROTLI_ARRAY_PARAM(input_size)], size_t* encoded_size,
    uint8_t encoded_buffer[BROTLI_ARRAY_PARAM(*encoded_size)]);
/* encode.h:473 -- PROCESS / FLUSH / FINISH / EMIT_METADATA */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderCompressStream(BrotliEncoderState* state, BrotliEncoderOperation op, size_t* available_in,
    const uint8_t** next_in, size_t* available_out, uint8_t** next_out, size_t* total_out);
/* encode.h:486 */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderIsFinished(BrotliEncoderState* state);
/* encode.h:495 */
BROTLI_ENC_API BROTLI_BOOL BrotliEncoderHasMoreOutput(BrotliEncoderState* state);
/* encode.h:526 -- pointer into internal storage, valid until the next call on the state */
BROTLI_ENC_API const uint8_t* BrotliEncoderTakeOutput(BrotliEncoderState* state, size_t* size);
/* encode.h:542 */
BROTLI_ENC_API uint32_t BrotliEncoderVersion(void);

#if defined(__cplusplus) || defined(c_plusplus)
}  /* extern "C" */
#endif
#endif  /* BROTLI_ENC_ENCODE_H_ */
