/* Force-included (-include) when compiling the reference's own C sources on Linux.
 * Supplies the few MSVC-only names they use (Network.c:26,114,152; comparator.c:26-27;
 * ViT_seq.c:371).  Test infrastructure only; contains no reference code. */
#ifndef VIT_REF_SHIM_H
#define VIT_REF_SHIM_H
#include <stdio.h>
#include <errno.h>
#include <string.h>
#include <time.h>
typedef int errno_t;
static inline errno_t fopen_s(FILE** f, const char* name, const char* mode) {
    *f = fopen(name, mode);
    return *f ? 0 : errno;
}
#define strncpy_s(dst, dsz, src, n) (strncpy((dst), (src), (n)), 0)
#ifndef CLK_TCK
#define CLK_TCK CLOCKS_PER_SEC
#endif
#endif
