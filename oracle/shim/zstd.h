/* Prototype-only shim for libzstd 1.5.5. TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_ZSTD_SHIM_H
#define ORACLE_ZSTD_SHIM_H
#include <stddef.h>
typedef enum { ZSTD_fast = 1, ZSTD_dfast = 2, ZSTD_greedy = 3, ZSTD_lazy = 4, ZSTD_lazy2 = 5,
               ZSTD_btlazy2 = 6, ZSTD_btopt = 7, ZSTD_btultra = 8, ZSTD_btultra2 = 9 } ZSTD_strategy;
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int level);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
#endif
