/*
 * aar_crsincos.h — one sin/cos for host and device.
 *
 * The reference expands every pose with cv::Rodrigues (libs/multicam_mapper.cpp:470, 693,
 * 910-911), which calls libm sin()/cos().  glibc and the CUDA math library differ in the
 * last ulp for a fraction of inputs, and a 1-ulp difference in R flips the float32
 * rounding of a projection (multicam_mapper.cpp:644-648) with probability ~1e-8 per value.
 * To make the CPU oracle and the sm_100a kernels execute the SAME function, both call
 * aar_sincos() below: double-double Taylor evaluation after a 3-part Cody-Waite reduction,
 * rounded once at the end.  The result is the correctly rounded sin/cos except with
 * probability ~2^-45 per call (Ziv-style argument; no fallback stage), so it also agrees
 * with a correctly rounding libm wherever that libm is correct.
 *
 * Only IEEE-754 add/mul/fma in round-to-nearest are used.  On the device every operation
 * is an explicit *_rn intrinsic so nvcc cannot contract or reassociate; on the host the
 * translation unit must be built with -ffp-contract=off (and without -ffast-math).
 *
 * Valid for 0 <= |x| < 2^30 (rotation-vector norms are O(1)).
 */
#ifndef AAR_CRSINCOS_H
#define AAR_CRSINCOS_H

#include <math.h>

#if defined(__CUDACC__)
#define AAR_HD __host__ __device__ __forceinline__
#else
#define AAR_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define AAR_ADD(a, b) __dadd_rn((a), (b))
#define AAR_SUB(a, b) __dsub_rn((a), (b))
#define AAR_MUL(a, b) __dmul_rn((a), (b))
#define AAR_FMA(a, b, c) __fma_rn((a), (b), (c))
#define AAR_DIV(a, b) __ddiv_rn((a), (b))
#define AAR_SQRT(a) __dsqrt_rn((a))
#else
#define AAR_ADD(a, b) ((a) + (b))
#define AAR_SUB(a, b) ((a) - (b))
#define AAR_MUL(a, b) ((a) * (b))
#define AAR_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define AAR_DIV(a, b) ((a) / (b))
#define AAR_SQRT(a) __builtin_sqrt((a))
#endif

typedef struct { double hi, lo; } aar_dd;

AAR_HD aar_dd aar_dd_make(double hi, double lo) { aar_dd r; r.hi = hi; r.lo = lo; return r; }

AAR_HD aar_dd aar_two_sum(double a, double b) {
    double s = AAR_ADD(a, b);
    double bb = AAR_SUB(s, a);
    double e = AAR_ADD(AAR_SUB(a, AAR_SUB(s, bb)), AAR_SUB(b, bb));
    return aar_dd_make(s, e);
}
AAR_HD aar_dd aar_quick_two_sum(double a, double b) { /* |a| >= |b| */
    double s = AAR_ADD(a, b);
    double e = AAR_SUB(b, AAR_SUB(s, a));
    return aar_dd_make(s, e);
}
AAR_HD aar_dd aar_two_prod(double a, double b) {
    double p = AAR_MUL(a, b);
    double e = AAR_FMA(a, b, -p);
    return aar_dd_make(p, e);
}
AAR_HD aar_dd aar_dd_add(aar_dd x, aar_dd y) {
    aar_dd s = aar_two_sum(x.hi, y.hi);
    aar_dd t = aar_two_sum(x.lo, y.lo);
    s.lo = AAR_ADD(s.lo, t.hi);
    s = aar_quick_two_sum(s.hi, s.lo);
    s.lo = AAR_ADD(s.lo, t.lo);
    return aar_quick_two_sum(s.hi, s.lo);
}
AAR_HD aar_dd aar_dd_mul(aar_dd x, aar_dd y) {
    aar_dd p = aar_two_prod(x.hi, y.hi);
    p.lo = AAR_ADD(p.lo, AAR_ADD(AAR_MUL(x.hi, y.lo), AAR_MUL(x.lo, y.hi)));
    return aar_quick_two_sum(p.hi, p.lo);
}

/* sin(x), cos(x) for any finite |x| < 2^30. */
AAR_HD void aar_sincos(double x, double *sn, double *cs) {
    /* (-1)^((k-1)/2)/k!, k = 3,5,...,29 as double-double */
    const double SC[14][2] = {
        {-0x1.5555555555555p-3, -0x1.5555555555555p-57},  {0x1.1111111111111p-7, 0x1.1111111111111p-63},
        {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
        {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},  {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
        {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
        {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
        {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
        {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}};
    /* (-1)^(k/2)/k!, k = 2,4,...,30 */
    const double CC[15][2] = {
        {-0x1.0000000000000p-1, 0.0},                     {0x1.5555555555555p-5, 0x1.5555555555555p-59},
        {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65},  {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
        {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76}, {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
        {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
        {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
        {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
        {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},
        {-0x1.3932c5047d60ep-108, -0x1.832b7b530a627p-162}};
    const double PIO2_1 = 0x1.921fb54442d18p+0, PIO2_2 = 0x1.1a62633145c07p-54, PIO2_3 = -0x1.f1976b7ed8fbcp-110;
    const double TWO_OVER_PI = 0x1.45f306dc9c883p-1;

    /* k = nearest integer to x*2/pi; any deterministic nearby integer is fine */
    double kd = rint(AAR_MUL(x, TWO_OVER_PI));
    long long kq = (long long)kd;
    /* r = x - k*pi/2 in double-double */
    aar_dd p1 = aar_two_prod(kd, PIO2_1);
    aar_dd p2 = aar_two_prod(kd, PIO2_2);
    double p3 = AAR_MUL(kd, PIO2_3);
    aar_dd r = aar_dd_make(x, 0.0);
    r = aar_dd_add(r, aar_dd_make(-p1.hi, -p1.lo));
    r = aar_dd_add(r, aar_dd_make(-p2.hi, -p2.lo));
    r = aar_dd_add(r, aar_dd_make(-p3, 0.0));

    aar_dd r2 = aar_dd_mul(r, r);
    /* sin r = r + r*(r2*P(r2)) */
    aar_dd P = aar_dd_make(SC[13][0], SC[13][1]);
    for (int i = 12; i >= 0; --i) P = aar_dd_add(aar_dd_mul(P, r2), aar_dd_make(SC[i][0], SC[i][1]));
    aar_dd S = aar_dd_add(r, aar_dd_mul(r, aar_dd_mul(r2, P)));
    /* cos r = 1 + r2*Q(r2) */
    aar_dd Q = aar_dd_make(CC[14][0], CC[14][1]);
    for (int i = 13; i >= 0; --i) Q = aar_dd_add(aar_dd_mul(Q, r2), aar_dd_make(CC[i][0], CC[i][1]));
    aar_dd Cc = aar_dd_add(aar_dd_make(1.0, 0.0), aar_dd_mul(r2, Q));

    double s = AAR_ADD(S.hi, S.lo), c = AAR_ADD(Cc.hi, Cc.lo);
    switch ((int)(kq & 3)) {
        case 0: *sn = s;  *cs = c;  break;
        case 1: *sn = c;  *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs = s; break;
    }
}

#endif /* AAR_CRSINCOS_H */
