/* Host check: pb_pow (patolette_b200/csrc/pow_glibc.h) vs libm pow, bit for bit. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "pow_glibc.h"

static uint64_t s[2] = { 0x9E3779B97F4A7C15ULL, 0xD1B54A32D192ED03ULL };
static uint64_t rnd(void) { /* xorshift128+ */
    uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b;
}
static double u01(void) { return (double)(rnd() >> 11) * 0x1p-53; }

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    const double ys[] = { 2.4, 1.0 / 2.4, 0.1593017578125, 78.84375, 1 / 0.1593017578125, 1 / 78.84375, 1.0 / 3.0, 3.0, 2.0, -1.5, 0.5 };
    const int ny = sizeof ys / sizeof ys[0];
    long bad = 0, total = 0;
    for (int j = 0; j < ny; j++) {
        for (long i = 0; i < n; i++) {
            double x;
            switch (i & 7) {
            case 0: x = u01(); break;
            case 1: x = u01() * 1.2; break;
            case 2: x = u01() * 1e-4; break;
            case 3: x = u01() * 1e4; break;
            case 4: x = 0.8 + 0.4 * u01(); break;
            case 5: x = exp((u01() - 0.5) * 1400); break;      /* full exponent range */
            case 6: x = -u01(); break;                          /* negative -> NaN or signed */
            default: { uint64_t b = rnd(); memcpy(&x, &b, 8); } /* raw bits */
            }
            double a = pow(x, ys[j]), b = pb_pow(x, ys[j], GLIBC_POW_LOG_TAB, GLIBC_EXP_TAB);
            total++;
            if (isnan(a) && isnan(b)) continue;
            if (memcmp(&a, &b, 8) != 0) {
                if (bad < 10) printf("MISMATCH x=%a y=%a libm=%a mine=%a\n", x, ys[j], a, b);
                bad++;
            }
        }
    }
    /* hand-picked specials */
    const double sx[] = { 0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, NAN, 0x1p-1074, 0x1p-1040, 1e308, 2.0, 0.5 };
    const double sy[] = { 0.0, -0.0, 1.0, -1.0, 3.0, 2.0, 0.5, INFINITY, -INFINITY, NAN, 1e-30, 1e30, -1e30, 1074.0, -1074.0, 2.4 };
    for (unsigned a = 0; a < sizeof sx / 8; a++) for (unsigned b = 0; b < sizeof sy / 8; b++) {
        double p = pow(sx[a], sy[b]), q = pb_pow(sx[a], sy[b], GLIBC_POW_LOG_TAB, GLIBC_EXP_TAB);
        total++;
        if (isnan(p) && isnan(q)) continue;
        if (memcmp(&p, &q, 8) != 0) { if (bad < 20) printf("SPECIAL MISMATCH x=%a y=%a libm=%a mine=%a\n", sx[a], sy[b], p, q); bad++; }
    }
    printf("checked=%ld mismatches=%ld\n", total, bad);
    return bad != 0;
}
