// TEST INFRASTRUCTURE ONLY (tests/cudasim): __half / __half2 over the compiler's _Float16.
#pragma once
#include <stdint.h>
#include <string.h>
struct __half { _Float16 v; };
struct __half2 { __half x, y; };
inline __half __float2half_rn(float f) { __half h; h.v = (_Float16)f; return h; }
inline float __half2float(__half h) { return (float)h.v; }
#ifndef CUDASIM_FLOAT2_DEFINED
#define CUDASIM_FLOAT2_DEFINED
#endif
