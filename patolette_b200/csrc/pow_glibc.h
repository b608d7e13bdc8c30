// pow_glibc.h - bit-exact restatement of glibc's double pow() (x86-64 FMA variant).
//
// The reference evaluates libm pow() up to 9 times per pixel in its colour
// transforms (lib/src/color/sRGB.c:70-110, eotf.c:29-57, CIELuv.c:54-164) and the
// parity bar for everything downstream is bit-exact, so the device code has to
// produce the very bits glibc produces.  CUDA's own pow() differs in the last
// place on a few percent of inputs.  This header re-derives, operation by
// operation, what `__pow_fma` of glibc 2.39 executes (the algorithm is Szabolcs
// Nagy's pow from ARM optimized-routines: log via a 128-entry table + degree-7
// polynomial in double-double, exp via a 128-entry 2^(k/128) table), including
// where the compiler fused multiply-adds - read off the disassembly of the
// system libm, since fusion points change the last bit.  The constant tables
// come from the same libm (tools/extract_glibc_pow_tables.py).
//
// Compiles as CUDA device code (built with -fmad=false; FMAs below are explicit)
// and as plain C++ on the host, where tests/test_pow_host.py checks it against
// libm pow() bit for bit on tens of millions of inputs.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdbool.h>

#include "glibc_pow_data.h"

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PB_HD static inline
#endif

#define PB_LOG_TAB(i) pb_log_tab_ptr[(i)]
#define PB_EXP_TAB(i) pb_exp_tab_ptr[(i)]
#if defined(__CUDA_ARCH__)
#define PB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PB_MUL(a, b) __dmul_rn((a), (b))
#define PB_ADD(a, b) __dadd_rn((a), (b))
#define PB_SUB(a, b) __dsub_rn((a), (b))
#else
#define PB_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define PB_MUL(a, b) ((a) * (b))
#define PB_ADD(a, b) ((a) + (b))
#define PB_SUB(a, b) ((a) - (b))
#endif

PB_HD double pb_asdouble(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
PB_HD uint64_t pb_asuint64(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

// Returns 0 if y is not an integer, 1 if odd, 2 if even (e_pow.c checkint).
PB_HD int pb_checkint(uint64_t iy) {
    int e = (int)((iy >> 52) & 0x7ff);
    if (e < 0x3ff) return 0;
    if (e > 0x3ff + 52) return 2;
    if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
    if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
    return 2;
}
PB_HD bool pb_zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * 0x7ff0000000000000ULL - 1; }

// exp(x + xtail) * (-1 if sign_bias), the second half of pow (e_pow.c exp_inline).
PB_HD double pb_exp_inline(double x, double xtail, uint32_t sign_bias,
                             const uint64_t *pb_exp_tab_ptr) {
    uint32_t abstop = (uint32_t)(pb_asuint64(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) { // |x| < 2^-54 or |x| >= 512
        if (abstop - 0x3c9u >= 0x80000000u) {
            double one = PB_ADD(1.0, x); // WANT_ROUNDING
            return sign_bias ? -one : one;
        }
        if (abstop >= 0x409u) { // |x| >= 1024: overflow / underflow
            double big = pb_asdouble(0x7000000000000000ULL), tiny = pb_asdouble(0x1000000000000000ULL);
            if (pb_asuint64(x) >> 63) return PB_MUL(sign_bias ? -tiny : tiny, tiny);
            return PB_MUL(sign_bias ? -big : big, big);
        }
        abstop = 0; // large x is special-cased below
    }
    const double Shift = pb_asdouble(GLIBC_EXP_SHIFT);
    // z = InvLn2N * x; kd = z + Shift   (fused in the binary)
    double kd = PB_FMA(x, pb_asdouble(GLIBC_EXP_INVLN2N), Shift);
    uint64_t ki = pb_asuint64(kd);
    kd = PB_SUB(kd, Shift);
    // r = x + kd*NegLn2hiN + kd*NegLn2loN   (two fused steps)
    double r = PB_FMA(kd, pb_asdouble(GLIBC_EXP_NEGLN2HIN), x);
    r = PB_FMA(kd, pb_asdouble(GLIBC_EXP_NEGLN2LON), r);
    r = PB_ADD(xtail, r);
    uint64_t idx = 2 * (ki % 128);
    uint64_t top = (ki + sign_bias) << 45;
    double tail = pb_asdouble(PB_EXP_TAB(idx));
    uint64_t sbits = PB_EXP_TAB(idx + 1) + top;
    double r2 = PB_MUL(r, r);
    // tmp = tail + r + r2*(C2 + r*C3) + r2*r2*(C4 + r*C5)
    double p23 = PB_FMA(r, pb_asdouble(GLIBC_EXP_C3), pb_asdouble(GLIBC_EXP_C2));
    double tr = PB_ADD(tail, r);
    double p45 = PB_FMA(r, pb_asdouble(GLIBC_EXP_C5), pb_asdouble(GLIBC_EXP_C4));
    double t = PB_FMA(p23, r2, tr);
    double r4 = PB_MUL(r2, r2);
    double tmp = PB_FMA(p45, r4, t);
    if (abstop == 0) { // e_pow.c specialcase()
        if ((ki & 0x80000000ULL) == 0) { // k > 0
            sbits -= 1009ULL << 52;
            double scale = pb_asdouble(sbits);
            return PB_MUL(PB_FMA(scale, tmp, scale), pb_asdouble(0x7f00000000000000ULL)); // * 0x1p1009
        }
        sbits += 1022ULL << 52; // k < 0: care in the subnormal range
        double scale = pb_asdouble(sbits);
        double st = PB_MUL(scale, tmp);
        double y = PB_ADD(scale, st);
        double ay = y < 0 ? -y : y;
        if (ay < 1.0) {
            double one = y < 0.0 ? -1.0 : 1.0;
            double lo = PB_ADD(PB_SUB(scale, y), st);
            double hi = PB_ADD(one, y);
            lo = PB_ADD(PB_ADD(PB_SUB(one, hi), y), lo);
            y = PB_SUB(PB_ADD(hi, lo), one);
            if (y == 0.0) y = pb_asdouble(sbits & 0x8000000000000000ULL);
        }
        return PB_MUL(pb_asdouble(0x0010000000000000ULL), y); // 0x1p-1022 * y
    }
    double scale = pb_asdouble(sbits);
    return PB_FMA(scale, tmp, scale);
}

// log_tab: GLIBC_POW_LOG_TAB (or a shared-memory copy), exp_tab: GLIBC_EXP_TAB (ditto).
PB_HD double pb_pow(double x, double y, const uint64_t *pb_log_tab_ptr,
                    const uint64_t *pb_exp_tab_ptr) {
    uint32_t sign_bias = 0;
    uint64_t ix = pb_asuint64(x), iy = pb_asuint64(y);
    uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        // Special cases: (x < 0x1p-126 or inf or nan) or (|y| < 0x1p-65 or |y| >= 0x1p63 or nan).
        if (pb_zeroinfnan(iy)) {
            if (2 * iy == 0) return 1.0; // issignaling(x) ignored
            if (ix == 0x3ff0000000000000ULL) return 1.0;
            if (2 * ix > 2 * 0x7ff0000000000000ULL || 2 * iy > 2 * 0x7ff0000000000000ULL) return PB_ADD(x, y);
            if (2 * ix == 2 * 0x3ff0000000000000ULL) return 1.0;
            if ((2 * ix < 2 * 0x3ff0000000000000ULL) == !(iy >> 63)) return 0.0; // |x|<1 && y==inf or |x|>1 && y==-inf
            return PB_MUL(y, y);
        }
        if (pb_zeroinfnan(ix)) {
            double x2 = PB_MUL(x, x);
            if (ix >> 63 && pb_checkint(iy) == 1) { x2 = -x2; sign_bias = 1; }
            // 1 / x2 for y < 0 (division by zero raises in libm; the value is the same)
            return (iy >> 63) ? (1.0 / x2) : x2;
        }
        if (ix >> 63) { // finite x < 0
            int yint = pb_checkint(iy);
            if (yint == 0) return pb_asdouble(0xfff8000000000000ULL); // __math_invalid: x86 default NaN
            if (yint == 1) sign_bias = 0x800u << 7; // SIGN_BIAS = 0x800 << EXP_TABLE_BITS
            ix &= 0x7fffffffffffffffULL;
            topx &= 0x7ff;
        }
        if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
            // Note: sign_bias == 0 here because y is not odd.
            if (ix == 0x3ff0000000000000ULL) return 1.0;
            if ((topy & 0x7ff) < 0x3beu) return ix > 0x3ff0000000000000ULL ? PB_ADD(1.0, y) : PB_SUB(1.0, y);
            // |y| huge
            double big = pb_asdouble(0x7000000000000000ULL), tiny = pb_asdouble(0x1000000000000000ULL);
            return ((ix > 0x3ff0000000000000ULL) == (topy < 0x800u)) ? PB_MUL(big, big) : PB_MUL(tiny, tiny);
        }
        if (topx == 0) { // subnormal x: normalise
            ix = pb_asuint64(PB_MUL(x, pb_asdouble(0x4330000000000000ULL))); // x * 0x1p52
            ix &= 0x7fffffffffffffffULL;
            ix -= 52ULL << 52;
        }
    }

    // ---- log_inline(ix, &lo) ------------------------------------------------
    uint64_t tmp = ix - 0x3fe6955500000000ULL;
    int i = (int)((tmp >> 45) & 127);
    int k = (int)((int64_t)tmp >> 52);
    uint64_t iz = ix - (tmp & (0xfffULL << 52));
    double z = pb_asdouble(iz);
    double kd = (double)k;
    double invc = pb_asdouble(PB_LOG_TAB(3 * i));
    double logc = pb_asdouble(PB_LOG_TAB(3 * i + 1));
    double logctail = pb_asdouble(PB_LOG_TAB(3 * i + 2));
    double r = PB_FMA(z, invc, -1.0);
    // k*Ln2 + log(c) + r   (both k*Ln2 products are fused in the binary)
    double t1 = PB_FMA(kd, pb_asdouble(GLIBC_POW_LN2HI), logc);
    double t2 = PB_ADD(t1, r);
    double lo1 = PB_FMA(kd, pb_asdouble(GLIBC_POW_LN2LO), logctail);
    double lo2 = PB_ADD(PB_SUB(t1, t2), r);
    const double A0 = pb_asdouble(GLIBC_POW_LOG_POLY[0]), A1 = pb_asdouble(GLIBC_POW_LOG_POLY[1]),
                 A2 = pb_asdouble(GLIBC_POW_LOG_POLY[2]), A3 = pb_asdouble(GLIBC_POW_LOG_POLY[3]),
                 A4 = pb_asdouble(GLIBC_POW_LOG_POLY[4]), A5 = pb_asdouble(GLIBC_POW_LOG_POLY[5]),
                 A6 = pb_asdouble(GLIBC_POW_LOG_POLY[6]);
    double ar = PB_MUL(A0, r);
    double ar2 = PB_MUL(r, ar);
    double ar3 = PB_MUL(r, ar2);
    double hi = PB_ADD(t2, ar2);
    double lo3 = PB_FMA(ar, r, -ar2);
    double lo4 = PB_ADD(PB_SUB(t2, hi), ar2);
    // p = ar3 * (A1 + r*A2 + ar2*(A3 + r*A4 + ar2*(A5 + r*A6))), Horner steps fused
    double q12 = PB_FMA(r, A2, A1);
    double q34 = PB_FMA(r, A4, A3);
    double q56 = PB_FMA(r, A6, A5);
    double inner = PB_FMA(q56, ar2, q34);
    double poly = PB_FMA(ar2, inner, q12);
    // lo = lo1 + lo2 + lo3 + lo4 + p   (p's product fused into the last add)
    double lo = PB_ADD(PB_ADD(PB_ADD(lo1, lo2), lo3), lo4);
    lo = PB_FMA(ar3, poly, lo);
    double lhi = PB_ADD(hi, lo);
    double llo = PB_ADD(PB_SUB(hi, lhi), lo);

    // ---- ehi + elo = y * (lhi + llo) ------------------------------------------
    double ehi = PB_MUL(y, lhi);
    double elo = PB_FMA(y, llo, PB_FMA(y, lhi, -ehi));
    return pb_exp_inline(ehi, elo, sign_bias, pb_exp_tab_ptr);
}
