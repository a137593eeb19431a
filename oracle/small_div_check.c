/* Test infrastructure (not part of the product): exhaustive check of the identity the env-step kernels use for observable
 * row 1 (immediate reward / max local reward, spinsystem.py:490 + the driver's fp32 cast, experiments/utils.py:174):
 *     (float)((double)a / (double)m)  ==  fmaf(fmaf(-q0, m, a), y, q0),   q0 = a * y,   y = RN_f32(1 / m)
 * for every integer |m| <= MMAX (m != 0) and |a| <= AMAX, signed zeros included.
 * Usage: small_div_check MMAX AMAX  ->  prints "<cases> <mismatches>". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv) {
    const int MMAX = argc > 1 ? atoi(argv[1]) : 2048, AMAX = argc > 2 ? atoi(argv[2]) : 70000;
    long bad = 0, n = 0;
    for (int m = -MMAX; m <= MMAX; ++m) {
        if (m == 0) continue;
        const float mf = (float)m;
        volatile float y = 1.0f / mf; /* correctly rounded fp32 reciprocal (__frcp_rn on the device) */
        for (int a = -AMAX; a <= AMAX; ++a) {
            const float af = (float)a;
            volatile float q0 = af * y;
            const float r = fmaf(-q0, mf, af);
            const float q = fmaf(r, y, q0);
            const float ref = (float)((double)a / (double)m);
            ++n;
            if (q != ref || signbit(q) != signbit(ref)) ++bad;
        }
    }
    printf("%ld %ld\n", n, bad);
    return 0;
}
