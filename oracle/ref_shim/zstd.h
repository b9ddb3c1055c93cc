// Declaration-level stand-in for <zstd.h> (the container has the libzstd.so.1 runtime but no headers): the four functions
// src/View/RawBinary.cpp:42-74 calls, with libzstd's published signatures.  TEST INFRASTRUCTURE ONLY (oracle/ref_shim).
#ifndef REF_SHIM_ZSTD_H_
#define REF_SHIM_ZSTD_H_
#include <cstddef>
extern "C" {
std::size_t ZSTD_compressBound(std::size_t srcSize);
std::size_t ZSTD_compress(void* dst, std::size_t dstCapacity, const void* src, std::size_t srcSize, int compressionLevel);
std::size_t ZSTD_decompress(void* dst, std::size_t dstCapacity, const void* src, std::size_t compressedSize);
unsigned ZSTD_isError(std::size_t code);
}
#endif
