/* Prototype-only shim for liblz4 1.9.4. TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_LZ4_SHIM_H
#define ORACLE_LZ4_SHIM_H
int LZ4_compress_default(const char *src, char *dst, int srcSize, int dstCapacity);
int LZ4_compressBound(int inputSize);
#endif
