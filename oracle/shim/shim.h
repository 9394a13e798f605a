/* Force-included portability shim for building the UNMODIFIED reference sources
 * (MSVC dialect) with g++ on Linux.  Test infrastructure only (oracle/_ref). */
#ifndef MIRO_ORACLE_SHIM_H
#define MIRO_ORACLE_SHIM_H
#ifdef __cplusplus
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <algorithm>
#endif
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <xmmintrin.h>
#ifdef INFINITY
#undef INFINITY            /* clashes with `const float INFINITY` in Miro.h:68 */
#endif
#define __forceinline inline __attribute__((always_inline))
#define __declspec(x) MIRO_SHIM_DECLSPEC_##x      /* __declspec(align(n)) is the only use in the tree */
#define MIRO_SHIM_DECLSPEC_align(n) __attribute__((aligned(n)))
#ifndef _MM_ALIGN16
#define _MM_ALIGN16 __attribute__((aligned(16)))
#endif
static inline void* _aligned_malloc(size_t size, size_t alignment) {
    void* p = 0;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    if (posix_memalign(&p, alignment, size ? size : alignment) != 0) return 0;
    return p;
}
static inline void _aligned_free(void* p) { free(p); }
/* ray counter bumped by the sed-instrumented Scene::trace (one closest/any-hit query = one ray) */
extern unsigned long long g_miro_trace_calls[32 * 16];
#endif
