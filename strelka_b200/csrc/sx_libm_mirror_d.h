// sx_libm_mirror_d.h -- bit-exact restatements of the double-precision libm routines inside the reference's site-model epilogue:
//
//   exp    <- normalizeLogDistro: prob[i] = std::exp(x[i] - max)                      blt_util/prob_util.hh:202
//   log10  <- error_prob_to_qphred: -10 * std::log10(prob)                            blt_util/qscore.hh:62-72
//   log    <- (inside log10: __ieee754_log10 calls __ieee754_log)
//
// The reference is linked against the host's glibc; on x86_64 CPUs with FMA + AVX2 glibc (>= 2.28) dispatches exp and log to the *_fma ifunc
// variants of the ARM "optimized routines" algorithms (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c), and log10 is the older
// __ieee754_log10 (e_log10.c: exponent split, ivln10 * log(mantissa), unfused).  CUDA's exp / log10 are within 1 ulp of these, which left
// `ref_pprob` equal to 1e-12 only and the integer qualities equal "as observed"; with these the double outputs of K2a are the reference's bits.
// The fused / unfused structure below was read off the disassembly of glibc 2.39's __exp_fma (libm.so.6 + 0x79b60), __log_fma (+ 0x79d50) and
// __ieee754_log10 (+ 0x2b6e0); the tables are the algorithms' published constants (tools/gen_libm_d_tables.py).  tests/test_libm_mirror.py checks
// the three against the live libm on ~10^8 arguments each, dense where the epilogue uses them (exp: [-1100, 0] incl. the subnormal results;
// log10: (0, 1] down to subnormals, and around 1).
//
// Domain: every double for exp; every non-negative double for log and log10 (negative arguments: the default NaN; NaN arguments are not
// propagated bit for bit).  No exception flag / errno is raised (nobody reads them).
#pragma once

#include "sx_libm_mirror.h"

// the tables live in global memory on the device (read through L1): the lanes of a warp look up different entries, which a __constant__ bank
// would serialise
#if defined(__CUDACC__)
#define SX_MIRROR_TAB static __device__ const
#else
#define SX_MIRROR_TAB static const
#endif
#include "sx_libm_mirror_d_tables.inc"

#if defined(__CUDACC__)
#define SX_DSUB(a, b) __dsub_rn((a), (b))
#else
#define SX_DSUB(a, b) ((a) - (b))
#endif

SX_HD double sx_exp(const double x)
{
    const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p+52, NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    const uint64_t ix = SX_D2U(x);
    uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu)
    {
        if (abstop - 0x3c9u >= 0x80000000u) return SX_DADD(1.0, x); // |x| < 2^-54
        if (abstop >= 0x409u)
        {
            if (ix == 0xfff0000000000000ull) return 0.0;
            if (abstop >= 0x7ffu) return SX_DADD(1.0, x);
            return (ix >> 63) ? 0.0 : SX_U2D(0x7ff0000000000000ull); // __math_uflow / __math_oflow
        }
        abstop = 0; // |x| in [512, 1024): the result may over- or underflow, handled at the end
    }
    double kd = SX_FMA(x, InvLn2N, Shift);
    const uint64_t ki = SX_D2U(kd);
    kd = SX_DSUB(kd, Shift);
    double r = SX_FMA(kd, NegLn2hiN, x);
    r = SX_FMA(kd, NegLn2loN, r);
    const uint32_t idx = 2u * (uint32_t)(ki & 127u);
    const uint64_t top = ki << 45;
    const double tail = SX_U2D(sx_exp_tab[idx]);
    uint64_t sbits = sx_exp_tab[idx + 1u] + top;
    const double p23 = SX_FMA(r, C3, C2);
    const double t3 = SX_DADD(r, tail);
    const double r2 = SX_DMUL(r, r);
    const double p45 = SX_FMA(r, C5, C4);
    const double q = SX_FMA(p23, r2, t3);
    const double r4 = SX_DMUL(r2, r2);
    const double tmp = SX_FMA(r4, p45, q);
    if (abstop == 0)
    {
        if ((ki & 0x80000000ull) == 0)
        {
            sbits -= 1009ull << 52; // k > 0: the exponent of scale might have overflowed
            const double scale = SX_U2D(sbits);
            return SX_DMUL(0x1p1009, SX_FMA(scale, tmp, scale));
        }
        sbits += 1022ull << 52; // k < 0: care in the subnormal range
        const double scale = SX_U2D(sbits);
        const double st = SX_DMUL(scale, tmp);
        double y = SX_DADD(scale, st);
        if (y < 1.0)
        {
            double lo = SX_DADD(SX_DSUB(scale, y), st);
            const double hi = SX_DADD(1.0, y);
            lo = SX_DADD(SX_DADD(SX_DSUB(1.0, hi), y), lo);
            y = SX_DSUB(SX_DADD(hi, lo), 1.0);
            if (y == 0.0) y = 0.0;
        }
        return SX_DMUL(0x1p-1022, y);
    }
    const double scale = SX_U2D(sbits);
    return SX_FMA(scale, tmp, scale);
}

// x: positive, finite (normal or subnormal)
SX_HD double sx_log(double x)
{
    const double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3, A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
    const double B0 = -0x1.0000000000000p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3, B3 = 0x1.999999995dd0cp-3, B4 = -0x1.55555556745a7p-3,
                 B5 = 0x1.24924a344de30p-3, B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4, B9 = 0x1.78182f7afd085p-4,
                 B10 = -0x1.5521375d145cdp-4;
    uint64_t ix = SX_D2U(x);
    if (ix - 0x3fee000000000000ull < 0x3090000000000ull) // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    {
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = SX_DSUB(x, 1.0);
        const double r2 = SX_DMUL(r, r);
        const double r3 = SX_DMUL(r, r2);
        const double pa = SX_FMA(r2, B3, SX_FMA(r, B2, B1));
        const double pb = SX_FMA(r2, B6, SX_FMA(r, B5, B4));
        double pc = SX_FMA(r2, B9, SX_FMA(r, B8, B7));
        pc = SX_FMA(r3, B10, pc);
        pc = SX_FMA(pc, r3, pb);
        pc = SX_FMA(pc, r3, pa);
        // hi + lo = r + r^2 * B0, in extra precision
        const double t = SX_FMA(r, 0x1p27, r);
        const double rhi = SX_FMA(-0x1p27, r, t);
        const double rlo = SX_DSUB(r, rhi);
        const double rhi2 = SX_DMUL(rhi, rhi);
        const double hi = SX_FMA(rhi2, B0, r);
        double lo = SX_FMA(rhi2, B0, SX_DSUB(r, hi));
        lo = SX_FMA(SX_DMUL(B0, rlo), SX_DADD(r, rhi), lo);
        const double y = SX_FMA(pc, r3, lo);
        return SX_DADD(hi, y);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u)
    {
        // subnormal (positive, non-zero): normalize; everything else is outside the domain and gets the qNaN the reference's log returns for x < 0
        if ((ix << 1) == 0) return SX_U2D(0xfff0000000000000ull); // log(+-0) = -inf
        if ((ix >> 63) || top >= 0x7ff0u) return ix == 0x7ff0000000000000ull ? x : SX_U2D(0x7ff8000000000000ull);
        ix = SX_D2U(SX_DMUL(x, 0x1p52));
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
    const int32_t k = (int32_t)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = sx_log_tab[2u * i], logc = sx_log_tab[2u * i + 1u];
    const double z = SX_U2D(iz);
    const double kd = (double)k;
    const double r = SX_FMA(z, invc, -1.0);
    const double w = SX_FMA(kd, Ln2hi, logc);
    const double hi = SX_DADD(r, w);
    double lo = SX_DADD(SX_DSUB(w, hi), r);
    lo = SX_FMA(kd, Ln2lo, lo);
    const double r2 = SX_DMUL(r, r);
    const double p12 = SX_FMA(r, A2, A1);
    const double r3 = SX_DMUL(r, r2);
    const double p34 = SX_FMA(r, A4, A3);
    lo = SX_FMA(r2, A0, lo);
    const double p = SX_FMA(p34, r2, p12);
    return SX_DADD(SX_FMA(r3, p, lo), hi);
}

// __ieee754_log10 (e_log10.c; no fused operation in it).  x: positive, finite
SX_HD double sx_log10(double x)
{
    const double ivln10 = 0x1.bcb7b1526e50ep-2, log10_2hi = 0x1.34413509f6000p-2, log10_2lo = 0x1.9fef311f12b36p-42;
    uint64_t hx = SX_D2U(x);
    int32_t k = -1023;
    if ((int64_t)hx <= 0xfffffffffffffll)
    {
        if ((hx & 0x7fffffffffffffffull) == 0) return SX_U2D(0xfff0000000000000ull); // log10(+-0) = -inf
        if (hx >> 63) return SX_U2D(0x7ff8000000000000ull);                            // log10(x < 0) = NaN
        x = SX_DMUL(x, 0x1p54);
        hx = SX_D2U(x);
        k = -1077;
    }
    if (hx > 0x7fefffffffffffffull) return SX_DADD(x, x);
    k += (int32_t)(hx >> 52);
    const uint32_t i = (uint32_t)k >> 31;
    const double y = (double)(int32_t)(k + (int32_t)i);
    hx = (hx & 0xfffffffffffffull) | ((uint64_t)(0x3ffu - i) << 52);
    const double t = SX_DMUL(y, log10_2lo);
    const double l = sx_log(SX_U2D(hx));
    const double z = SX_DADD(SX_DMUL(ivln10, l), t);
    return SX_DADD(z, SX_DMUL(y, log10_2hi));
}

// __log1p_fma (sysdeps/ieee754/dbl-64/s_log1p.c, the fdlibm algorithm; fused operations as in glibc 2.39's libm.so.6 + 0x7aff0).  x > -1
// (x <= -1 and NaN return the default NaN / -inf like the reference's, without the exceptions)
SX_HD double sx_log1p(const double x)
{
    const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33;
    const double Lp1 = 0x1.5555555555593p-1, Lp2 = 0x1.999999997fa04p-2, Lp3 = 0x1.2492494229359p-2, Lp4 = 0x1.c71c51d8e78afp-3, Lp5 = 0x1.7466496cb03dep-3,
                 Lp6 = 0x1.39a09d078c69fp-3, Lp7 = 0x1.2f112df3e5244p-3;
    const uint64_t ux = SX_D2U(x);
    const int32_t hx = (int32_t)(ux >> 32);
    const int32_t ax = hx & 0x7fffffff;
    int32_t k = 1, hu = 0;
    double f = 0.0, c = 0.0;
    if (hx < 0x3FDA827A) // x < 0.41422
    {
        if (ax >= 0x3ff00000) return x == -1.0 ? SX_U2D(0xfff0000000000000ull) : SX_U2D(0x7ff8000000000000ull); // x <= -1
        if (ax < 0x3e200000)                                                                                       // |x| < 2^-29
        {
            if (ax < 0x3c900000) return x; // |x| < 2^-54
            return SX_FMA(-SX_DMUL(x, x), 0.5, x);
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) // -0.2929 < x < 0.41422
        {
            k = 0;
            f = x;
            hu = 1;
        }
    }
    else if (hx >= 0x7ff00000) return SX_DADD(x, x);
    if (k != 0)
    {
        double u;
        if (hx < 0x43400000)
        {
            u = SX_DADD(1.0, x);
            hu = (int32_t)(SX_D2U(u) >> 32);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? SX_DSUB(1.0, SX_DSUB(u, x)) : SX_DSUB(x, SX_DSUB(u, 1.0)); // correction term
            c = c / u;
        }
        else
        {
            u = x;
            hu = (int32_t)(SX_D2U(u) >> 32);
            k = (hu >> 20) - 1023;
            c = 0.0;
        }
        hu &= 0x000fffff;
        if (hu < 0x6a09e) u = SX_U2D((SX_D2U(u) & 0xffffffffull) | ((uint64_t)(uint32_t)(hu | 0x3ff00000) << 32)); // normalize u
        else
        {
            k += 1;
            u = SX_U2D((SX_D2U(u) & 0xffffffffull) | ((uint64_t)(uint32_t)(hu | 0x3fe00000) << 32)); // normalize u / 2
            hu = (0x00100000 - hu) >> 2;
        }
        f = SX_DSUB(u, 1.0);
    }
    const double hfsq = SX_DMUL(SX_DMUL(0.5, f), f);
    const double kd = (double)k;
    if (hu == 0) // |f| < 2^-20
    {
        if (f == 0.0)
        {
            if (k == 0) return 0.0;
            return SX_FMA(kd, ln2_hi, SX_FMA(kd, ln2_lo, c));
        }
        const double R = SX_DMUL(SX_FMA(-f, 0x1.5555555555555p-1, 1.0), hfsq);
        if (k == 0) return SX_DSUB(f, R);
        return SX_FMA(kd, ln2_hi, -SX_DSUB(SX_DSUB(R, SX_FMA(kd, ln2_lo, c)), f));
    }
    const double s = f / SX_DADD(f, 2.0);
    const double z = SX_DMUL(s, s);
    const double R2 = SX_FMA(z, Lp3, Lp2), R3 = SX_FMA(z, Lp5, Lp4), R4 = SX_FMA(z, Lp7, Lp6);
    const double z2 = SX_DMUL(z, z), z4 = SX_DMUL(z2, z2), z6 = SX_DMUL(z2, z4);
    double R = SX_FMA(z, Lp1, SX_DMUL(z2, R2));
    R = SX_FMA(z4, R3, R);
    R = SX_FMA(z6, R4, R);
    const double t = SX_DMUL(SX_DADD(R, hfsq), s);
    if (k == 0) return SX_DSUB(f, SX_DSUB(hfsq, t));
    return SX_FMA(kd, ln2_hi, -SX_DSUB(SX_DSUB(hfsq, SX_DADD(SX_FMA(kd, ln2_lo, c), t)), f));
}

// __expf_fma (sysdeps/ieee754/flt-32/e_expf.c: the ARM algorithm, double arithmetic inside; fused operations as in glibc 2.39's libm.so.6 +
// 0x7dc40).  The somatic model's float log-sum (getLogSum<float>: std::exp(float)) is the one caller.  Every float argument.
SX_MIRROR_TAB uint64_t sx_expf_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull,
    0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull,
    0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull,
    0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
SX_HD float sx_expf(const float x)
{
    const double Shift = 0x1.8p+52, InvLn2N = 0x1.71547652b82fep+5, C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
    const uint32_t ix = SX_F2U(x);
    const uint32_t abstop = (ix >> 20) & 0x7ffu;
    if (abstop > 0x42au) // |x| >= 88 or NaN
    {
        if (ix == 0xff800000u) return 0.0f;
        if (abstop > 0x7f7u) return x + x;
        if (x > 0x1.62e42ep6f) return SX_U2F(0x7f800000u); // overflow
        if (x < -0x1.9fe368p6f) return 0.0f;              // underflow
        if (x < -0x1.9d1d9ep6f) return SX_U2F(0x00000001u); // __math_may_uflowf: 0x1.4p-75f * 0x1.4p-75f rounds to the smallest subnormal
    }
    const double xd = (double)x;
    double kd = SX_FMA(InvLn2N, xd, Shift);
    const uint64_t ki = SX_D2U(kd);
    kd = SX_DSUB(kd, Shift);
    const double r = SX_FMA(InvLn2N, xd, -kd);
    const double s = SX_U2D(sx_expf_tab[ki & 31u] + (ki << 47));
    const double z = SX_FMA(r, C0, C1);
    const double r2 = SX_DMUL(r, r);
    double y = SX_FMA(r, C2, 1.0);
    y = SX_FMA(z, r2, y);
    y = SX_DMUL(y, s);
    return (float)y;
}
