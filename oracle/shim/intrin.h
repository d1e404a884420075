/* stand-in for MSVC <intrin.h>: the hot path only needs the x86 SIMD intrinsics */
#pragma once
#include <immintrin.h>
/* the C++ standard headers pulled in above may #undef the Windows-style min macro */
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
