/* Hand-written config.h so the unmodified reference compiles without autotools.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef ORACLE_CONFIG_H
#define ORACLE_CONFIG_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE 1
#endif
#define PACKAGE "lrzip-next"
#define PACKAGE_VERSION "0.14.0"
#define LRZIP_MAJOR_VERSION 0
#define LRZIP_MINOR_VERSION 14
#define LRZIP_MINOR_SUBVERSION 0
#define HAVE_SYS_MMAN_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_RESOURCE_H 1
#define HAVE_UNISTD_H 1
#define HAVE_ERRNO_H 1
#define HAVE_ENDIAN_H 1
#define HAVE_ARPA_INET_H 1
#define HAVE_PTHREAD_H 1
#define HAVE_STRING_H 1
#define HAVE_MALLOC_H 1
#define HAVE_ALLOCA_H 1
#define HAVE_CTYPE_H 1
#define HAVE_STRERROR 1
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_SHORT 2
#define __UNUSED__ __attribute__((unused))
#define LIBBZ3_ABI1 1
#endif
