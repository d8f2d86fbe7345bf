/* aar_acos.h — one arc cosine for both sides of the IPPE parity check.
 *
 * aruco's IPPERot2vec (3rdparty/aruco/aruco/ippe.cpp:355-378 of the reference) calls libm's acos; CUDA's acos and glibc's are
 * both accurate to about an ulp but not bit-identical, and the pose the Initializer keeps is rounded to float32 afterwards
 * (getRTMatrix(..., CV_32F), ippe.cpp:116-122), so a one-ulp difference can flip a float.  The kernel (csrc/aar_init.cuh) and
 * the oracle in its default mode therefore share this evaluation: the rational approximation of acos on [-1, 1] that Sun's
 * fdlibm uses (error below one ulp), written with +, -, *, /, sqrt only — IEEE operations that give the same bits on the
 * host (-ffp-contract=off) and on the device (-fmad=false).  The oracle's mode 0 keeps libm, and the tests measure how
 * often the two modes differ after the float32 rounding. */
#ifndef AAR_ACOS_H
#define AAR_ACOS_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef __CUDACC__
#define AAR_ACOS_HD __host__ __device__ inline
#else
#define AAR_ACOS_HD static inline
#endif

AAR_ACOS_HD double aar_acos_ratio(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}

AAR_ACOS_HD double aar_acos(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    const double ax = fabs(x);
    if (!(ax < 1.0)) {
        if (x == 1.0) return 0.0;
        if (x == -1.0) return pi + 2.0 * pio2_lo;
        return (x - x) / (x - x);                       /* NaN outside [-1, 1] */
    }
    if (ax < 0.5) {
        if (ax <= 6.938893903907228e-18) return pio2_hi + pio2_lo;      /* |x| < 2^-57 */
        const double r = aar_acos_ratio(x * x);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0) {
        const double z = (1.0 + x) * 0.5, s = sqrt(z), r = aar_acos_ratio(z), w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    {
        const double z = (1.0 - x) * 0.5, s = sqrt(z);
        uint64_t bits; double df;
        memcpy(&bits, &s, 8); bits &= 0xffffffff00000000ull; memcpy(&df, &bits, 8);      /* s with its low word cleared */
        const double c = (z - df * df) / (s + df), r = aar_acos_ratio(z), w = r * s + c;
        return 2.0 * (df + w);
    }
}
#endif
