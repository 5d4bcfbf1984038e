/* ttc_detexp.h — a DETERMINISTIC exp(x) built from + - * and integer/bit operations only.
 *
 * Why it exists: the exp-based integrands (stdnorm, test_crs_stdnorm.f90:154-170; MVN, lib/mvn_pdf.f90:63-83) call the
 * platform's libm, and CUDA's exp and glibc's exp differ in the last ulp of a few arguments.  On problems where pivot
 * candidates are (near-)ties -- the equicorrelated MVN of config E is symmetric under permutations of its variables --
 * that last ulp decides which candidate wins, so a GPU run and a CPU run drift apart although neither is wrong (two libm
 * builds of the reference itself would do the same).  In "parity mode" (ttc_set_exp_mode(h, 1) on the product side, the
 * same switch in the test oracle) both sides evaluate exp through THIS routine: identical operations in identical order,
 * no FMA contraction (the library is built with -fmad=false, the oracle with -ffp-contract=off), hence identical bits, and
 * the pivot tapes must then agree to the end.  Accuracy: Cody-Waite reduction by ln2 in two pieces + degree-13 Taylor
 * polynomial on |r| <= 0.35: error < 2 ulp.  The default mode of the product stays the platform exp (what a user of the
 * reference gets); parity mode is a test instrument.
 */
#ifndef TTC_DETEXP_H
#define TTC_DETEXP_H

#if defined(__CUDACC__)
#define TTC_HD __host__ __device__ __forceinline__
#else
#define TTC_HD static inline
#endif

TTC_HD double ttc_det_pow2(int k) {           /* 2^k for -1022 <= k <= 1023, by exponent-field construction */
    union { unsigned long long u; double d; } v;
    v.u = (unsigned long long)(k + 1023) << 52;
    return v.d;
}
TTC_HD double ttc_det_exp(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) { union { unsigned long long u; double d; } v; v.u = 0x7ff0000000000000ULL; return v.d; }
    if (x < -745.2) return 0.0;
    const double LOG2E = 1.44269504088896338700e+00;
    const double LN2HI = 6.93147180369123816490e-01;   /* 33 significant bits: k * LN2HI is exact */
    const double LN2LO = 1.90821492927058770002e-10;
    double t = x * LOG2E;
    long long ki = (long long)(t >= 0.0 ? t + 0.5 : t - 0.5);
    const double kf = (double)ki;
    const double r = (x - kf * LN2HI) - kf * LN2LO;
    /* exp(r) = sum r^n / n!, Horner, n <= 13 */
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    /* scale by 2^k in two steps so that results in the subnormal range and k = 1024 are formed by ordinary multiplications */
    const int k = (int)ki;
    const int k1 = k / 2, k2 = k - k1;
    return (p * ttc_det_pow2(k1)) * ttc_det_pow2(k2);
}
#endif
