// cusim stand-in for <cuda_fp16.h> (TEST INFRASTRUCTURE, see cuda_runtime.h): only what csrc/mmaconv.cuh names.
#pragma once
#include "cuda_runtime.h"
struct __half2 { uint32_t bits; };
static inline float2 __half22float2(__half2 h) { return cusim::unpack_f16x2(h.bits); }
