// extern "C" doorway to the reference's own dsn::utils::crc64_calc / crc32_calc
// (src/utils/crc.h:37,56), so tests can pin the oracle's crc64 against the real thing.
#include <cstddef>
#include <cstdint>
#include "utils/crc.h"
extern "C" __attribute__((visibility("default"))) uint64_t ref_crc64(const void *p, size_t n, uint64_t init)
{
    return dsn::utils::crc64_calc(p, n, init);
}
extern "C" __attribute__((visibility("default"))) uint32_t ref_crc32(const void *p, size_t n, uint32_t init)
{
    return dsn::utils::crc32_calc(p, n, init);
}
