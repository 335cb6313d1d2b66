// TEST INFRASTRUCTURE ONLY (tests/cudasim): host runtime stand-in of tests/hostsim plus the SIMT emulator.
#pragma once
#include "../hostsim/cuda_runtime.h"
#include "cudasim.h"
#include "cuda_fp16.h"
inline float2 __half22float2(__half2 h) { float2 r; r.x = (float)h.x.v; r.y = (float)h.y.v; return r; }
