/* brotli/types.h as seen by users of libbrotlienc_b200 -- same names and values as the reference's
 * c/include/brotli/types.h:49-81 (BROTLI_BOOL is int, the allocator callback pair), written for this library.
 * Shares the reference's include guard: whichever copy is found first wins, both describe the same ABI. */
#ifndef BROTLI_COMMON_TYPES_H_
#define BROTLI_COMMON_TYPES_H_
#include <stddef.h>
#include <stdint.h>

#define BROTLI_BOOL int
#define BROTLI_TRUE 1
#define BROTLI_FALSE 0
#define TO_BROTLI_BOOL(X) (!!(X) ? BROTLI_TRUE : BROTLI_FALSE)
#define BROTLI_MAKE_UINT64_T(high, low) ((((uint64_t)(high)) << 32) | low)
#define BROTLI_UINT32_MAX (~((uint32_t)0))
#define BROTLI_SIZE_MAX (~((size_t)0))

/* allocate `size` bytes; NULL on failure (types.h:73) */
typedef void* (*brotli_alloc_func)(void* opaque, size_t size);
/* release a block obtained from the paired allocator; address may be NULL (types.h:81) */
typedef void (*brotli_free_func)(void* opaque, void* address);
#endif  /* BROTLI_COMMON_TYPES_H_ */
