/* Prototype-only shim for libbz2 1.0.8. TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_BZLIB_SHIM_H
#define ORACLE_BZLIB_SHIM_H
#define BZ_OK 0
#define BZ_OUTBUFF_FULL (-8)
int BZ2_bzBuffToBuffCompress(char *dest, unsigned int *destLen, char *source, unsigned int sourceLen,
                             int blockSize100k, int verbosity, int workFactor);
int BZ2_bzBuffToBuffDecompress(char *dest, unsigned int *destLen, char *source, unsigned int sourceLen,
                               int small, int verbosity);
#endif
