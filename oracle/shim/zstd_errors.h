#ifndef ORACLE_ZSTD_ERRORS_SHIM_H
#define ORACLE_ZSTD_ERRORS_SHIM_H
enum { ZSTD_error_dstSize_tooSmall = 70 };
#endif
