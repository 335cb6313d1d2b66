// TEST INFRASTRUCTURE ONLY (tests/cudasim): __nv_bfloat16 with round-to-nearest-even conversion.
#pragma once
#include <stdint.h>
#include <string.h>
struct __nv_bfloat16 { uint16_t x; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  __nv_bfloat16 r;
  if ((u & 0x7fffffffu) > 0x7f800000u) { r.x = 0x7fff; return r; }   // NaN -> canonical NaN, as the device does
  u += 0x7fffu + ((u >> 16) & 1u);
  r.x = (uint16_t)(u >> 16);
  return r;
}
inline float __bfloat162float(__nv_bfloat16 h) { uint32_t u = (uint32_t)h.x << 16; float f; memcpy(&f, &u, 4); return f; }
