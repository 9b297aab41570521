/*
 * ref_kt_fixups.cpp — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 * Compiles two kt translation units from /root/reference in place, with the two POSIX problems of the submodule
 * worked around by the preprocessor instead of by editing the files:
 *   - kt/src/kt/Memory.cpp:29 passes Min(16, align) to posix_memalign: it fails for align < 8 and under-aligns
 *     the 32-byte aligned tiles.  Every allocation is made 64-byte aligned instead.
 *   - kt/src/kt/Concurrency.cpp:187-194 LogicalCoreCount() has no POSIX branch (falls off the end).  The broken
 *     definition is renamed away and a run-time controlled one is supplied, so the harness chooses the reference's
 *     worker count (Renderer.cpp:141): 1 == the single-threaded canonical build.
 */
#include <stdlib.h>
#include <stdint.h>

#define posix_memalign(pp, al, sz) posix_memalign(pp, 64, sz)
#include "kt/src/kt/Memory.cpp"
#undef posix_memalign

#define LogicalCoreCount LogicalCoreCount_reference_nonposix
#include "kt/src/kt/Concurrency.cpp"
#undef LogicalCoreCount

uint32_t g_srref_logical_cores = 1;

namespace kt
{
uint32_t LogicalCoreCount()
{
	return g_srref_logical_cores;
}
}
