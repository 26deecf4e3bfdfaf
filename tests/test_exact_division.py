"""The kernels replace IEEE divisions by the FMA-corrected quotient q' = q + (x - q*d)*rcp(d) wherever
the operands come from a small known domain (itw_device.cuh: div_by_rcp).  That quotient is NOT
correctly rounded in general, so every domain it is used on is checked exhaustively here (C, OpenMP).
The integer-threshold form of the two-bit index search (bc7.cuh) is checked against the float formula here too."""
import os
import subprocess
import tempfile

SRC = r"""
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline float dq(float x, float d, float r) { float q = x * r; return fmaf(fmaf(-q, d, x), r, q); }
int main(void) {
    long long bad = 0;
    /* (1) x / 255 for EVERY float: only -0 (result +0) and +-inf (NaN) may differ */
    const float r255 = 1.0f / 255.0f;
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (long long i = 0; i < (1ll << 32); i++) {
        uint32_t u = (uint32_t)i; float x; memcpy(&x, &u, 4);
        if (x != x || isinf(x) || u == 0x80000000u) continue;
        float a = x / 255.0f, b = dq(x, 255.0f, r255);
        uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
        bad += (ua != ub);
    }
    printf("div255 %lld\n", bad);
    /* (2) integer x in [0, 2^24] divided by a count 1..16 (moments / count) */
    bad = 0;
    for (int n = 1; n <= 16; n++) {
        float fn = (float)n, rn = 1.0f / fn;
        #pragma omp parallel for reduction(+:bad)
        for (int x = 0; x <= (1 << 24); x++) bad += ((float)x / fn != dq((float)x, fn, rn));
    }
    printf("count %lld\n", bad);
    /* (3) integer |x| <= 2^18 by integer d in [1, 2^18] (index-search projection; symmetric in sign) */
    bad = 0;
    #pragma omp parallel for reduction(+:bad) schedule(dynamic, 64)
    for (int d = 1; d <= (1 << 18); d++) {
        float fd = (float)d, rd = 1.0f / fd; long long b2 = 0;
        #pragma omp simd reduction(+:b2)
        for (int x = 0; x <= (1 << 18); x++) { float fx = (float)x; b2 += (fx / fd != dq(fx, fd, rd)); }
        bad += b2;
    }
    printf("proj %lld\n", bad);
    /* (4) scalar channel of BC7 modes 4/5: integer x in [-255,255] by (n + 0.001f), n in [-255,255] */
    bad = 0;
    for (int n = -255; n <= 255; n++) {
        float d = (float)n + 0.001f, r = 1.0f / d;
        for (int x = -255; x <= 255; x++) bad += ((float)x / d != dq((float)x, d, r));
    }
    printf("scalar %lld\n", bad);
    /* (5) BC4/BC5 ramp coefficients: integer 0..7 by 5 or 7 */
    bad = 0;
    for (int d = 5; d <= 7; d += 2)
        for (int x = 0; x <= 7; x++) bad += ((float)x / (float)d != dq((float)x, (float)d, 1.0f / (float)d));
    printf("ramp %lld\n", bad);
    /* (6) two-bit index search by integer thresholds (bc7_assign): q1 = clamp((int)(x / d * 4 + 0.5), 1, 3) against
       1 + [x >= ceil(3d/8)] + [x >= ceil(5d/8)], d in [1, 2^18] (the sum of <= 4 squared byte differences is < 2^18).
       Both are non-decreasing step functions of the integer x, so they are equal everywhere iff they are equal on both sides of
       every step and at the ends of the domain; small d are also swept exhaustively. */
    bad = 0;
    #pragma omp parallel for reduction(+:bad) schedule(dynamic, 64)
    for (int d = 1; d <= (1 << 18); d++) {
        const float fd = (float)d;
        const int t2 = (3 * d + 7) >> 3, t3 = (5 * d + 7) >> 3;
        const int probe[8] = {-(1 << 18), -1, 0, t2 - 1, t2, t3 - 1, t3, 1 << 18};
        const int lo = (d <= 4096) ? -d - 8 : 0, hi = (d <= 4096) ? 2 * d + 8 : 7;
        for (int i = lo; i <= hi; i++) {
            const int x = (d <= 4096) ? i : probe[i];
            volatile float p = (float)x / fd; volatile float p4 = p * 4.0f; volatile float y = p4 + 0.5f;
            int q = (int)y; q = q < 1 ? 1 : (q > 3 ? 3 : q);
            const int want = 1 + (x >= t2) + (x >= t3);
            bad += (q != want);
        }
    }
    printf("thresholds %lld\n", bad);
    return 0;
}
"""


def test_fma_corrected_quotient_is_exact_on_every_domain_it_is_used_on():
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "t.c")
        exe = os.path.join(tmp, "t")
        open(c, "w").write(SRC)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-mavx2", "-mfma", "-ffp-contract=off", c, "-o", exe, "-lm"])
        out = subprocess.check_output([exe], text=True).split()
    assert out == ["div255", "0", "count", "0", "proj", "0", "scalar", "0", "ramp", "0", "thresholds", "0"], out
