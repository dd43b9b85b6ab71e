/* brotli/port.h for libbrotlienc_b200: the export / deprecation macros the public encoder header uses
 * (reference: c/include/brotli/port.h:237-305).  ELF + GCC/Clang only: this library exists for Linux + CUDA. */
#ifndef BROTLI_COMMON_PORT_H_
#define BROTLI_COMMON_PORT_H_
#if defined(__GNUC__)
#define BROTLI_PUBLIC __attribute__((visibility("default")))
#define BROTLI_INTERNAL __attribute__((visibility("hidden")))
#define BROTLI_DEPRECATED __attribute__((deprecated))
#else
#define BROTLI_PUBLIC
#define BROTLI_INTERNAL
#define BROTLI_DEPRECATED
#endif
#define BROTLI_COMMON_API BROTLI_PUBLIC
#define BROTLI_DEC_API BROTLI_PUBLIC
#define BROTLI_ENC_API BROTLI_PUBLIC
#define BROTLI_ENC_EXTRA_API BROTLI_INTERNAL
/* array sizes in prototypes: a C99 feature, not C++ (port.h:290-305) */
#if !defined(__cplusplus) && defined(__STDC_VERSION__) && (__STDC_VERSION__ >= 199901L) && !defined(__STDC_NO_VLA__)
#define BROTLI_ARRAY_PARAM(name) (name)
#else
#define BROTLI_ARRAY_PARAM(name)
#endif
#endif  /* BROTLI_COMMON_PORT_H_ */
