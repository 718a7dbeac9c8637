// sx_device_util.cuh -- small device helpers shared by the site-model kernels.
//
// Arithmetic discipline: the reference's site models run in un-fused IEEE float/double (x86-64 SSE2, no FMA, FLT_EVAL_METHOD 0).
// Every float/double operation that takes part in a value the reference would also compute is written with an explicit
// round-to-nearest intrinsic (or compiled under -fmad=false), so the compiler can never contract a*b+c.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sx_libm_mirror_d.h"

__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_div(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// error_prob_to_qphred<double>  blt_util/qscore.hh:40-66 :  floor(-10*max(-307, log10(p)) + 0.5); log10 = the reference's libm's, bit for bit
__device__ __forceinline__ int error_prob_to_qphred_d(double prob)
{
    const double l = sx_log10(prob);
    const double m = (-307.0 < l) ? l : -307.0; // std::max(minlog10, l): returns minlog10 unless minlog10 < l
    return static_cast<int>(floor(d_add(d_mul(-10.0, m), 0.5)));
}

// ln_error_prob_to_qphred<float>  blt_util/qscore.hh:50-72 with FloatType = float:
//   ln10 = std::log(10.f) (float); lnProb/ln10 in float; -10.*max(...) in double, narrowed to float on return;
//   then floor(float + 0.5) in double.
__device__ __forceinline__ int ln_error_prob_to_qphred_f(float lnProb, float ln10f)
{
    const float d = f_div(lnProb, ln10f);
    const float m = (-37.0f < d) ? d : -37.0f;
    const float r = static_cast<float>(d_mul(-10.0, static_cast<double>(m)));
    return static_cast<int>(floor(d_add(static_cast<double>(r), 0.5)));
}
