/* Prototype-only shim: liblzo2 is ABSENT in this image; symbols are stubbed. */
#ifndef ORACLE_LZOCONF_SHIM_H
#define ORACLE_LZOCONF_SHIM_H
#include <stddef.h>
typedef size_t lzo_uint;
typedef unsigned char *lzo_bytep;
typedef void *lzo_voidp;
#define LZO_E_OK 0
#define LZO_OK 0
int lzo_init(void);
typedef int (*lzo_compress_t)(const unsigned char *src, lzo_uint src_len, unsigned char *dst, lzo_uint *dst_len, void *wrkmem);
#endif
