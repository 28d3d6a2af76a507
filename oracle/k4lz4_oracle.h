/* k4lz4_oracle.h -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 * See k4lz4_oracle.c for the reference file:line each function follows. */
#ifndef K4LZ4_ORACLE_H
#define K4LZ4_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define K4O_PICKLE_CORRUPT (-1000)   /* stands for the reference's InvalidDataException */

int k4o_max_output_size(int n);
int k4o_compress_fast(const uint8_t *src, int n, uint8_t *dst, int cap, int enforce32);
int k4o_decompress_safe(const uint8_t *src, int n, uint8_t *dst, int cap);
int k4o_codec_encode(const uint8_t *src, int n, uint8_t *dst, int cap, int level, int enforce32);
int k4o_codec_decode(const uint8_t *src, int n, uint8_t *dst, int cap);
int k4o_decompress_safe_usingDict(const uint8_t *src, int n, uint8_t *dst, int cap, const uint8_t *dict, int dictSize);
int k4o_decompress_safe_partial(const uint8_t *src, int n, uint8_t *dst, int target);
int k4o_codec_decode_dict(const uint8_t *src, int n, uint8_t *dst, int cap, const uint8_t *dict, int dictSize);
int k4o_codec_partial_decode(const uint8_t *src, int n, uint8_t *dst, int target);
int k4o_pickle_bound(int n);
int k4o_pickle(const uint8_t *src, int n, uint8_t *dst, uint8_t *scratch, int level);
int k4o_pickle_writer_bound(int n);
int k4o_pickle_writer(const uint8_t *src, int n, uint8_t *dst, int level);
int k4o_unpickled_size(const uint8_t *src, int n);
int k4o_unpickle(const uint8_t *src, int n, uint8_t *dst, int dstLen);

uint32_t k4o_xxh32(const uint8_t *p, size_t len, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
