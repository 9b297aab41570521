/* Shim for the three Win32 names the reference's TaskSystem.cpp uses
 * (/root/reference/SoftRast/TaskSystem.cpp:2,38,186): VirtualAlloc + flags and _mm_pause.
 * Test infrastructure only (oracle/_ref build). */
#pragma once
#include <stddef.h>
#include <sys/mman.h>
#include <immintrin.h>
#define MEM_COMMIT 0x1000
#define MEM_RESERVE 0x2000
#define PAGE_READWRITE 0x04
static inline void* VirtualAlloc(void* addr, size_t size, int, int)
{
	void* p = mmap(addr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	return p == MAP_FAILED ? nullptr : p;
}
