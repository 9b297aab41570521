/* Shim for the MSVC <intrin.h> the reference's Texture.cpp includes
 * (/root/reference/SoftRast/Texture.cpp:2). Test infrastructure only. */
#pragma once
#include <x86intrin.h>
