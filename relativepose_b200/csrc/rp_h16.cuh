// rp_h16.cuh -- the 16-bit operand / activation-storage format of the tensor-core convolution kernels.
//
// RP_H16_FP16 = 1 (default): IEEE half (11-bit significand).  The networks' activations are BatchNorm-normalised and the
// raw convolution outputs of normalised inputs are O(1..100), far inside half's +-65504 range, so the three extra
// significand bits over bfloat16 come for free: the descriptor head's max-abs error against the float32 reference drops
// ~8x (tests/test_gpu_via_completion.py prints both).  RP_H16_FP16 = 0 restores bfloat16 (range of float32).
// tcgen05.mma kind::f16 takes either; only the a_format / b_format bits of the instruction descriptor differ.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifndef RP_H16_FP16
#define RP_H16_FP16 1
#endif

#if RP_H16_FP16
typedef __half rp_h16;
typedef __half2 rp_h162;
#define RP_H16_UMMA_FORMAT 0
__device__ __forceinline__ rp_h162 rp_f2_to_h2(float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ float2 rp_h2_to_f2(rp_h162 v) { return __half22float2(v); }
__device__ __forceinline__ rp_h16 rp_f_to_h(float a) { return __float2half_rn(a); }
__device__ __forceinline__ float rp_h_to_f(rp_h16 a) { return __half2float(a); }
#else
typedef __nv_bfloat16 rp_h16;
typedef __nv_bfloat162 rp_h162;
#define RP_H16_UMMA_FORMAT 1
__device__ __forceinline__ rp_h162 rp_f2_to_h2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 rp_h2_to_f2(rp_h162 v) { return __bfloat1622float2(v); }
__device__ __forceinline__ rp_h16 rp_f_to_h(float a) { return __float2bfloat16_rn(a); }
__device__ __forceinline__ float rp_h_to_f(rp_h16 a) { return __bfloat162float(a); }
#endif
