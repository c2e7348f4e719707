// Host-side check for a planned k_rollup optimisation: x / s with s = dt/1000 (dt = window span in ms) through a
// cached reciprocal r = RN(1/s) and one Markstein correction step,
//     q = RN(x*r);  rem = fma(-q, s, x);  q' = fma(rem, r, q)
// compared with the IEEE division Go performs (rollup.go:1988).  Counts mismatches over every dt in [1, DT_MAX] x random
// numerators of the shapes rate() sees (differences of decimal values).  Build: gcc -O2 -march=native -fopenmp -o /tmp/mk/a scripts/exp_markstein_div.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double ms_to_s(int64_t dt) {  // same as the kernel's: exact dt/1e3
    const double x = (double)dt, r = 1e-3;
    double q = x * r;
    double rem = fma(-q, 1e3, x);
    return fma(rem, r, q);
}
int main(int argc, char** argv) {
    const int64_t dt_max = argc > 1 ? atoll(argv[1]) : (1 << 22);
    const int per = argc > 2 ? atoi(argv[2]) : 256;
    unsigned long long bad = 0, total = 0, bad_s = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : bad, total, bad_s)
    for (int64_t dt = 1; dt <= dt_max; dt++) {
        const double s = ms_to_s(dt);
        if (s != (double)dt / 1e3) bad_s++;
        const double r = 1.0 / s;
        uint64_t st = 0x1234567ull * (uint64_t)dt;
        for (int k = 0; k < per; k++) {
            uint64_t u = sm64(&st);
            double x;
            switch (k & 7) {
                case 0: x = (double)(int64_t)(u >> 20) / 100.0; break;                      // decimal hundredths
                case 1: x = (double)(u >> 11); break;                                        // 53-bit integers
                case 2: x = (double)(int64_t)(u >> 30) * 1e-3; break;
                case 3: x = ldexp((double)(u >> 11), (int)(u & 63) - 80); break;             // wide exponents
                case 4: x = -(double)(int64_t)(u >> 24) / 10.0; break;
                case 5: { uint64_t b = (u & 0x000FFFFFFFFFFFFFull) | ((uint64_t)(1023 + (int)(u >> 58)) << 52); memcpy(&x, &b, 8); } break;  // random mantissas
                case 6: x = (double)(int64_t)(u >> 34) * (double)dt / 1e3; break;            // near-exact quotients
                default: x = (double)((u >> 40) * (uint64_t)dt) / 1e3 + ((u & 1) ? 1e-9 : 0.0); break;
            }
            double q = x * r;
            double rem = fma(-q, s, x);
            double q2 = fma(rem, r, q);
            double want = x / s;
            if (memcmp(&q2, &want, 8) != 0) bad++;
            total++;
        }
    }
    printf("dt in [1, %lld], %d numerators each: %llu checks, %llu mismatches; ms_to_s mismatches %llu\n", (long long)dt_max, per, total, bad, bad_s);
    return bad != 0;
}
