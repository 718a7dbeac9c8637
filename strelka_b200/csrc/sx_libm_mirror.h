// sx_libm_mirror.h -- bit-exact restatements of the two single-precision libm routines that sit INSIDE the
// reference's per-call float arithmetic and therefore cannot be tabulated on the host:
//
//   logf  <- val[0] = std::log(eprob) with float eprob        blt_common/position_snp_call_pprob_digt.cpp:352
//   powf  <- val = std::pow(eprob, vexp) with float operands  blt_common/adjust_joint_eprob.cpp:66
//
// The reference is linked against the host's glibc; on every x86_64 server CPU with FMA+AVX2 glibc (>= 2.28)
// dispatches logf/powf to the *_fma ifunc variants of the ARM "optimized routines" algorithms
// (sysdeps/ieee754/flt-32/e_logf.c, e_powf.c).  The integer VCF fields (PL, GQ) are rounded from float sums of
// these values, so a 1-ulp difference flips a PL about once per 1e4-1e5 sites: the device has to reproduce the
// exact operation sequence, including which multiply-adds are fused.  The fused/unfused structure below was read
// off the disassembly of glibc 2.39's __logf_fma / __powf_fma; the tables are the algorithm's published constants.
// tests/test_libm_mirror.py checks both routines against the live libm (exhaustively for logf).
//
// Only the "main path" is mirrored: finite, positive, normal x (and for powf |y*log2(x)| < 126).  Callers
// guarantee that domain (error probabilities of phred 3..70 and exponents in [min_vexp, 1]); outside it the
// functions return NaN so a parity test fails loudly rather than silently diverging.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
// device build: __device__-only functions, tables in constant memory, explicit rounding intrinsics
#define SX_HD __device__ __forceinline__
#define SX_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SX_DMUL(a, b) __dmul_rn((a), (b))
#define SX_DADD(a, b) __dadd_rn((a), (b))
#define SX_F2U(x) __float_as_uint(x)
#define SX_U2F(x) __uint_as_float(x)
#define SX_D2U(x) ((uint64_t)__double_as_longlong(x))
#define SX_U2D(x) __longlong_as_double((long long)(x))
#define SX_MIRROR_CONST static __constant__
#else
// host build (oracle test exports only): compile with -ffp-contract=off so that only SX_FMA fuses
#include <string.h>
#define SX_HD static inline
#define SX_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define SX_DMUL(a, b) ((a) * (b))
#define SX_DADD(a, b) ((a) + (b))
static inline uint32_t sx_f2u_(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float sx_u2f_(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
static inline uint64_t sx_d2u_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double sx_u2d_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define SX_F2U(x) sx_f2u_(x)
#define SX_U2F(x) sx_u2f_(x)
#define SX_D2U(x) sx_d2u_(x)
#define SX_U2D(x) sx_u2d_(x)
#define SX_MIRROR_CONST static const
#endif

// {invc, logc} pairs: c near the centre of the i-th of 16 subintervals of [OFF, 2*OFF), OFF = 0x3f330000
SX_MIRROR_CONST double sx_logf_tab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aap-1,  0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

// same invc, logc = log2(c)
SX_MIRROR_CONST double sx_powf_log2_tab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
    0x1.49539f0f010bp+0,  -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8eap+0,  -0x1.97c1d1b3b7afp-3,
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4,  0x1.ca4b31f026aap-1,  0x1.476a9543891bap-3,
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3,  0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2,  0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2};

// 2^(i/32) with the exponent bits pre-adjusted (asuint64(2^(i/32)) - (i << 47))
SX_MIRROR_CONST uint64_t sx_exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// glibc __logf (FMA variant).  x must be positive, finite and normal.
SX_HD float sx_logf(float x)
{
    uint32_t ix = SX_F2U(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return SX_U2F(0x7fc00000u);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = sx_logf_tab[2 * i];
    const double logc = sx_logf_tab[2 * i + 1];
    const double z = (double)SX_U2F(iz);
    const double r = SX_FMA(z, invc, -1.0);
    const double y0 = SX_FMA((double)k, 0x1.62e42fefa39efp-1, logc);
    const double r2 = SX_DMUL(r, r);
    double y = SX_FMA(r, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2);
    y = SX_FMA(r2, -0x1.00ea348b88334p-2, y);
    y = SX_FMA(r2, y, SX_DADD(r, y0));
    return (float)y;
}

// glibc __powf (FMA variant), main path only: x positive finite normal, y finite, |y*log2(x)| < 126.
SX_HD float sx_powf(float x, float y)
{
    const uint32_t ix = SX_F2U(x);
    const uint32_t iy = SX_F2U(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return SX_U2F(0x7fc00000u);
    if (2u * iy - 1u >= 2u * 0x7f800000u - 1u) return SX_U2F(0x7fc00000u);
    // log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int32_t)top >> 23;
    const double invc = sx_powf_log2_tab[2 * i];
    const double logc = sx_powf_log2_tab[2 * i + 1];
    const double z = (double)SX_U2F(iz);
    const double r = SX_FMA(z, invc, -1.0);
    const double y0 = SX_DADD(logc, (double)k);
    double yy = SX_FMA(r, 0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2);
    const double p = SX_FMA(r, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1);
    const double r2 = SX_DMUL(r, r);
    double q = SX_FMA(r, 0x1.71547652ab82bp+0, y0);
    const double r4 = SX_DMUL(r2, r2);
    q = SX_FMA(r2, p, q);
    yy = SX_FMA(yy, r4, q);
    const double ylogx = SX_DMUL((double)y, yy);
    if (((SX_D2U(ylogx) >> 47) & 0xffffu) > 0x80beu) return SX_U2F(0x7fc00000u);
    // exp2_inline
    const double shift = 0x1.8p+47;
    double kd = SX_DADD(ylogx, shift);
    const uint64_t ki = SX_D2U(kd);
    kd = SX_DADD(kd, -shift);
    const double rr = SX_DADD(ylogx, -kd);
    uint64_t t = sx_exp2f_tab[ki & 31u];
    t += ki << 47;
    const double s = SX_U2D(t);
    const double zz = SX_FMA(rr, 0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3);
    const double rr2 = SX_DMUL(rr, rr);
    double e = SX_FMA(rr, 0x1.62e42ff0c52d6p-1, 1.0);
    e = SX_FMA(zz, rr2, e);
    e = SX_DMUL(e, s);
    return (float)e;
}
