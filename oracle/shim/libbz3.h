/* Prototype-only shim: libbzip3 is ABSENT in this image; symbols are stubbed to abort. */
#ifndef ORACLE_LIBBZ3_SHIM_H
#define ORACLE_LIBBZ3_SHIM_H
#include <stdint.h>
#include <stddef.h>
#define BZ3_OK 0
struct bz3_state;
struct bz3_state *bz3_new(int32_t block_size);
void bz3_free(struct bz3_state *state);
int8_t bz3_last_error(struct bz3_state *state);
const char *bz3_strerror(struct bz3_state *state);
int32_t bz3_encode_block(struct bz3_state *state, uint8_t *buffer, int32_t size);
int32_t bz3_decode_block(struct bz3_state *state, uint8_t *buffer, size_t buffer_size, int32_t size, int32_t orig_size);
#endif
