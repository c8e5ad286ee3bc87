#ifndef ORACLE_LZO1X_SHIM_H
#define ORACLE_LZO1X_SHIM_H
#include "lzoconf.h"
#define LZO1X_1_MEM_COMPRESS (16384L * sizeof(void *))
#define LZO1X_999_MEM_COMPRESS (14 * 16384L * sizeof(short))
int lzo1x_1_compress(const unsigned char *src, lzo_uint src_len, unsigned char *dst, lzo_uint *dst_len, void *wrkmem);
int lzo1x_999_compress(const unsigned char *src, lzo_uint src_len, unsigned char *dst, lzo_uint *dst_len, void *wrkmem);
int lzo1x_decompress_safe(const unsigned char *src, lzo_uint src_len, unsigned char *dst, lzo_uint *dst_len, void *wrkmem);
#endif
