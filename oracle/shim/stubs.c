/* Stubs for libraries that are absent from this image (liblzo2, libbzip3).
 * None of the BASELINE configs that can be pinned here reach them. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <stddef.h>
static void missing(const char *what) { fprintf(stderr, "oracle/_ref: %s is not available in this image\n", what); abort(); }
int lzo_init(void) { return 0; }
int lzo1x_1_compress(const unsigned char *a, size_t b, unsigned char *c, size_t *d, void *e) { (void)a;(void)b;(void)c;(void)d;(void)e; missing("liblzo2"); return -1; }
int lzo1x_999_compress(const unsigned char *a, size_t b, unsigned char *c, size_t *d, void *e) { (void)a;(void)b;(void)c;(void)d;(void)e; missing("liblzo2"); return -1; }
int lzo1x_decompress_safe(const unsigned char *a, size_t b, unsigned char *c, size_t *d, void *e) { (void)a;(void)b;(void)c;(void)d;(void)e; missing("liblzo2"); return -1; }
struct bz3_state;
struct bz3_state *bz3_new(int32_t bs) { (void)bs; missing("libbzip3"); return NULL; }
void bz3_free(struct bz3_state *s) { (void)s; }
int8_t bz3_last_error(struct bz3_state *s) { (void)s; return -1; }
const char *bz3_strerror(struct bz3_state *s) { (void)s; return "libbzip3 missing"; }
int32_t bz3_encode_block(struct bz3_state *s, uint8_t *b, int32_t n) { (void)s;(void)b;(void)n; missing("libbzip3"); return -1; }
int32_t bz3_decode_block(struct bz3_state *s, uint8_t *b, size_t bs, int32_t n, int32_t o) { (void)s;(void)b;(void)bs;(void)n;(void)o; missing("libbzip3"); return -1; }
