/* Prototype-only shim for libgcrypt 1.10 (runtime .so present, headers absent).
 * Only what the reference calls. TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_GCRYPT_SHIM_H
#define ORACLE_GCRYPT_SHIM_H
#include <stddef.h>
typedef unsigned int gpg_error_t;
typedef unsigned int gpg_err_code_t;
typedef gpg_error_t gcry_error_t;
typedef struct gcry_md_handle *gcry_md_hd_t;
typedef struct gcry_cipher_handle *gcry_cipher_hd_t;
enum { GCRY_MD_MD5 = 1, GCRY_MD_RMD160 = 3, GCRY_MD_SHA256 = 8, GCRY_MD_SHA384 = 9,
       GCRY_MD_SHA512 = 10, GCRY_MD_CRC32 = 302, GCRY_MD_SHA3_256 = 313,
       GCRY_MD_SHA3_512 = 315, GCRY_MD_SHAKE128 = 316, GCRY_MD_SHAKE256 = 317 };
enum { GCRY_MD_FLAG_SECURE = 1 };
enum { GCRY_CIPHER_AES128 = 7, GCRY_CIPHER_AES256 = 9 };
enum { GCRY_CIPHER_MODE_CBC = 3 };
enum { GCRY_CIPHER_SECURE = 1, GCRY_CIPHER_CBC_CTS = 4 };
enum { GCRY_KDF_SCRYPT = 48 };
gcry_error_t gcry_md_open(gcry_md_hd_t *h, int algo, unsigned int flags);
void gcry_md_close(gcry_md_hd_t h);
void gcry_md_reset(gcry_md_hd_t h);
void gcry_md_write(gcry_md_hd_t h, const void *buf, size_t len);
unsigned char *gcry_md_read(gcry_md_hd_t h, int algo);
gcry_error_t gcry_md_extract(gcry_md_hd_t h, int algo, void *buf, size_t len);
void gcry_create_nonce(void *buf, size_t len);
gcry_error_t gcry_kdf_derive(const void *pass, size_t passlen, int algo, int subalgo,
                             const void *salt, size_t saltlen, unsigned long iter,
                             size_t keysize, void *key);
gcry_error_t gcry_cipher_open(gcry_cipher_hd_t *h, int algo, int mode, unsigned int flags);
void gcry_cipher_close(gcry_cipher_hd_t h);
gcry_error_t gcry_cipher_setkey(gcry_cipher_hd_t h, const void *key, size_t len);
gcry_error_t gcry_cipher_setiv(gcry_cipher_hd_t h, const void *iv, size_t len);
gcry_error_t gcry_cipher_encrypt(gcry_cipher_hd_t h, void *out, size_t outsize, const void *in, size_t inlen);
gcry_error_t gcry_cipher_decrypt(gcry_cipher_hd_t h, void *out, size_t outsize, const void *in, size_t inlen);
int gpg_strerror_r(gpg_error_t err, char *buf, size_t buflen);
static inline gpg_err_code_t gpg_err_code(gpg_error_t err) { return (gpg_err_code_t)(err & 65535u); }
#endif
