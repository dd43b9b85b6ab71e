/* brotli_b200.h -- C ABI of libbrotlienc_b200.so.
 *
 * Drop-in boundary: the encoder half of google/brotli's public API,
 * /root/reference/c/include/brotli/encode.h (function by function below), with identical
 * names, argument meaning, enum values and error behaviour, so that the `brotli` CLI
 * (c/tools/brotli.c:1357), the Python extension (python/_brotli.c:403,529), the Go wrapper
 * (go/cbrotli/writer.go:30) and the JNI wrapper (java/org/brotli/wrapper/enc/encoder_jni.cc:78)
 * can link against this library instead of libbrotlienc.  See INTEGRATION.md.
 *
 * The hot path runs on an NVIDIA B200 (sm_100a).  There is NO CPU implementation inside
 * the library: without a CUDA device, or for parameters outside the implemented path
 * (quality 5..9, lgwin 17..24, no custom dictionary, modes GENERIC/TEXT), the entry points
 * return BROTLI_FALSE / NULL instead of producing different bytes.
 *
 * Plain C: pointers and sizes only, no torch / CUDA types.
 */
#ifndef BROTLI_B200_H_
#define BROTLI_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BROTLI_BOOL int              /* c/include/brotli/types.h:49 */
#define BROTLI_TRUE 1
#define BROTLI_FALSE 0
typedef void* (*brotli_alloc_func)(void* opaque, size_t size);   /* types.h:73 */
typedef void (*brotli_free_func)(void* opaque, void* address);   /* types.h:81 */

typedef enum BrotliEncoderMode {     /* encode.h:45 */
  BROTLI_MODE_GENERIC = 0, BROTLI_MODE_TEXT = 1, BROTLI_MODE_FONT = 2
} BrotliEncoderMode;
typedef enum BrotliEncoderOperation {  /* encode.h:93 */
  BROTLI_OPERATION_PROCESS = 0, BROTLI_OPERATION_FLUSH = 1,
  BROTLI_OPERATION_FINISH = 2, BROTLI_OPERATION_EMIT_METADATA = 3
} BrotliEncoderOperation;
typedef enum BrotliEncoderParameter {  /* encode.h:160 */
  BROTLI_PARAM_MODE = 0, BROTLI_PARAM_QUALITY = 1, BROTLI_PARAM_LGWIN = 2,
  BROTLI_PARAM_LGBLOCK = 3, BROTLI_PARAM_DISABLE_LITERAL_CONTEXT_MODELING = 4,
  BROTLI_PARAM_SIZE_HINT = 5, BROTLI_PARAM_LARGE_WINDOW = 6, BROTLI_PARAM_NPOSTFIX = 7,
  BROTLI_PARAM_NDIRECT = 8, BROTLI_PARAM_STREAM_OFFSET = 9, BROTLI_PARAM_BASE64_MODE = 10,
  BROTLI_PARAM_MAX_BASE64_REGIONS = 11, BROTLI_PARAM_SIMD_HASHER = 12
} BrotliEncoderParameter;

typedef struct BrotliEncoderStateStruct BrotliEncoderState;                 /* encode.h:271 */
typedef struct BrotliEncoderPreparedDictionaryStruct BrotliEncoderPreparedDictionary;

#define BROTLI_B200_API __attribute__((visibility("default")))

/* encode.h:289 */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderSetParameter(BrotliEncoderState* state, BrotliEncoderParameter param, uint32_t value);
/* encode.h:306 */
BROTLI_B200_API BrotliEncoderState* BrotliEncoderCreateInstance(brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque);
/* encode.h:314 */
BROTLI_B200_API void BrotliEncoderDestroyInstance(BrotliEncoderState* state);
/* encode.h:343 -- custom dictionaries are outside the implemented path: returns NULL */
BROTLI_B200_API BrotliEncoderPreparedDictionary* BrotliEncoderPrepareDictionary(int type, size_t data_size, const uint8_t* data, int quality, brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque);
/* encode.h:348 */
BROTLI_B200_API void BrotliEncoderDestroyPreparedDictionary(BrotliEncoderPreparedDictionary* dictionary);
/* encode.h:361 -- returns BROTLI_FALSE */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderAttachPreparedDictionary(BrotliEncoderState* state, const BrotliEncoderPreparedDictionary* dictionary);
/* encode.h:375 */
BROTLI_B200_API size_t BrotliEncoderMaxCompressedSize(size_t input_size);
/* encode.h:405 -- one-shot; the bytes equal the reference's for the same arguments */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderCompress(int quality, int lgwin, BrotliEncoderMode mode, size_t input_size, const uint8_t* input_buffer, size_t* encoded_size, uint8_t* encoded_buffer);
/* encode.h:473 */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderCompressStream(BrotliEncoderState* state, BrotliEncoderOperation op, size_t* available_in, const uint8_t** next_in, size_t* available_out, uint8_t** next_out, size_t* total_out);
/* encode.h:486 */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderIsFinished(BrotliEncoderState* state);
/* encode.h:495 */
BROTLI_B200_API BROTLI_BOOL BrotliEncoderHasMoreOutput(BrotliEncoderState* state);
/* encode.h:526 */
BROTLI_B200_API const uint8_t* BrotliEncoderTakeOutput(BrotliEncoderState* state, size_t* size);
/* encode.h:542 */
BROTLI_B200_API uint32_t BrotliEncoderVersion(void);

/* ---- B200 extensions (not in the reference API) ------------------------------------ */
/* Same as BrotliEncoderCompress but input and output live in DEVICE memory of the current
 * CUDA device (plain device pointers).  *encoded_size: in = capacity, out = bytes written. */
BROTLI_B200_API BROTLI_BOOL BrotliB200CompressDevice(int quality, int lgwin, size_t input_size, const void* d_input, size_t* encoded_size, void* d_encoded);
/* Many independent streams (one stream per shard / per object).  Inputs and
 * outputs are host pointers.  Quality 5..9: streams are spread over `threads` host workers, each with
 * its own CUDA stream.  Quality 1 (compress_fragment_two_pass.c:612, one independent fragment coder per
 * stream): the whole batch is ONE device batch -- four kernel launches, one copy each way -- and `threads`
 * only parallelises the host-side packing.  encoded_sizes[i]: in = capacity of outputs[i], out = bytes written.
 * Returns the number of streams compressed successfully. */
BROTLI_B200_API size_t BrotliB200CompressBatch(int quality, int lgwin, size_t count, const uint8_t* const* inputs, const size_t* input_sizes, uint8_t* const* outputs, size_t* encoded_sizes, int threads);
/* Quality 1, device resident: `count` non-empty streams sit in ONE device buffer, stream i at d_inputs + input_offsets[i]
 * (host array), input_sizes[i] bytes (host array).  The compressed streams are packed into d_encoded (device), stream i at
 * d_encoded + encoded_offsets[i] (host array of count + 1 entries, 16-byte aligned starts, last = bytes used),
 * encoded_sizes[i] bytes (host array; 0 = that stream must go through the host call: it compressed to more than
 * BrotliEncoderMaxCompressedSize).  Bytes equal BrotliEncoderCompress(1, lgwin, GENERIC, ...) per stream.
 * Returns the number of streams compressed. */
BROTLI_B200_API size_t BrotliB200CompressBatchDevice(int quality, int lgwin, size_t count, const void* d_inputs, const uint64_t* input_offsets, const size_t* input_sizes, void* d_encoded, size_t encoded_capacity, uint64_t* encoded_offsets, size_t* encoded_sizes);
/* Timings (milliseconds, CUDA events) and counters of the calling thread's last compress call.
 * out[0..15]: total, index (hash+sort), lz77, entropy, assemble, lz77 iterations, block runs,
 * blocks, metablocks, kernel launches, summed k_walk ms, k_encode_mb ms, k_walk launches,
 * k_encode_mb launches, input bytes walked over all k_walk launches, commands emitted. */
BROTLI_B200_API void BrotliB200LastStats(double out[16]);
/* quality-1 batch pipeline of the calling thread's last call: ms total / h2d / parse / code / pack / d2h, then
   streams, fragments, blocks, input bytes, compressed bytes, launches */
BROTLI_B200_API void BrotliB200LastStatsQ1(double out[12]);
/* 1 if a usable CUDA device is present. */
BROTLI_B200_API int BrotliB200Available(void);

#ifdef __cplusplus
}
#endif
#endif  /* BROTLI_B200_H_ */
