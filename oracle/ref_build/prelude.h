/* Force-included (-include) ahead of every reference translation unit when building oracle/_ref.
 * It supplies the std headers kt forgets on Linux (/root/reference/kt/src/kt/inl/kt.inl:123-125,
 * inl/Memory.inl:31) and renames two identifiers so the reference sources compile IN PLACE, unmodified:
 *   - IAllocator has no Free(); kt::Delete calls one (/root/reference/kt/src/kt/inl/Memory.inl:45).
 *     Renaming every `Free(` token to `FreeUnsized(` in all TUs is self-consistent.
 *   - ::_stricmp (/root/reference/kt/src/kt/Strings.cpp:79) is strcasecmp on POSIX.
 * Test infrastructure only. */
#pragma once
#ifdef __cplusplus
#include <utility>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <new>
#endif
#include <string.h>
#include <strings.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#define Free(...) FreeUnsized(__VA_ARGS__)
#define _stricmp strcasecmp
