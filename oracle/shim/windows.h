/* Minimal stand-in for <windows.h>: just enough typedefs/macros for the two
 * hot-path translation units of the reference (deconvolve.cpp, rschecksf.cpp)
 * to compile with g++ on Linux.  Test infrastructure only (see oracle/README.md). */
#pragma once
#include <stdint.h>
#include <stddef.h>
typedef uint32_t DWORD;
typedef uint64_t DWORD64;
typedef void* PVOID;
typedef void* HANDLE;
typedef int BOOL;
typedef int32_t LONG;
typedef int64_t LONG64;
typedef unsigned char BOOLEAN;
#define MAX_PATH 260
#define WINAPI
#define __declspec(x) __attribute__((x))
#define align(a) aligned(a)
#define __forceinline inline __attribute__((always_inline))
#define UNREFERENCED_PARAMETER(x) (void)(x)
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
