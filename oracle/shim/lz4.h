// Build shim (test infrastructure): prototypes of the three LZ4 entry points the
// reference's doc store calls; the image ships liblz4.so.1 without headers.
#ifndef WSR_SHIM_LZ4_H
#define WSR_SHIM_LZ4_H
#ifdef __cplusplus
extern "C" {
#endif
int LZ4_compress_default(const char *src, char *dst, int srcSize, int dstCapacity);
int LZ4_decompress_safe(const char *src, char *dst, int compressedSize, int dstCapacity);
int LZ4_compressBound(int inputSize);
#ifdef __cplusplus
}
#endif
#endif
