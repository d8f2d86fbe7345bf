// aar_device_math.cuh — the arithmetic kernel of the hot path, sm_100a.
//
// Reproduces, operation by operation, what MultiCamMapper::project_marker
// (/root/reference/libs/multicam_mapper.cpp:608-649) evaluates through OpenCV — but only the
// operations whose results can differ from zero/one for rigid transforms:
//   * cv::Mat operator* on 4x4 / 3x3*3x4 / 3x4*4x4 (cv::gemm small-matrix path): every output is
//     the left-associated sum of separately rounded products, k ascending.  With the last row of a
//     rigid transform equal to [0 0 0 1] the k=3 term of the rotation part is an exact +0 and the
//     k=3 term of the translation is the exact addend t_i, so the 4x4 product collapses to
//         R_ij = (Ra_i0*Rb_0j + Ra_i1*Rb_1j) + Ra_i2*Rb_2j
//         t_i  = ((Ra_i0*tb_0 + Ra_i1*tb_1) + Ra_i2*tb_2) + ta_i
//     with bit-identical results (x + 0 = x, x*1 = x).  Signs of zeros are not preserved; they never
//     reach a result (a zero coordinate is only ever compared or subtracted).
//   * the marker corner matrix X has z = 0 and w = 1 (multicam_mapper.cpp:261-270), so column 2 of
//     the composed transform is never needed and P_i = (A_i0*x + A_i1*y) + A_i3 with x,y = +-h.
//   * cv::Rodrigues (vector -> matrix) and cv::Mat::inv() (LU with partial pivoting) are restated in
//     full because they run once per camera / marker / frame, not per observation.
// The whole translation unit is compiled with -fmad=false: `a*b + c` below is two IEEE operations,
// exactly like the baseline-ISA OpenCV build the oracle is pinned to; fused multiply-adds appear only
// where written as fma().
#pragma once
#include <cfloat>
#include <cuda_runtime.h>
#include "../../include/aar_crsincos.h"

namespace aar {

// 3x4 rigid transform, row-major rotation + translation
struct Pose { double r[9]; double t[3]; };

__device__ __forceinline__ void load_pose(Pose &p, const double *__restrict__ src) {
#pragma unroll
    for (int i = 0; i < 9; i++) p.r[i] = src[i];
#pragma unroll
    for (int i = 0; i < 3; i++) p.t[i] = src[9 + i];
}

// cv::Rodrigues(r -> R) — mcm.cpp:470; same operation order as oracle rodrigues_vec2mat()
__device__ __forceinline__ void rodrigues(double rx, double ry, double rz, double *R) {
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double s, c;
    aar_sincos(theta, &s, &c);
    double c1 = 1. - c;
    double itheta = 1. / theta;
    rx *= itheta; ry *= itheta; rz *= itheta;
    // R = c*I + c1*r*rT + s*[r]x ; c*0 and s*0 terms are exact zeros
    R[0] = c + c1 * (rx * rx);            R[1] = c1 * (rx * ry) + s * (-rz);    R[2] = c1 * (rx * rz) + s * ry;
    R[3] = c1 * (rx * ry) + s * rz;       R[4] = c + c1 * (ry * ry);            R[5] = c1 * (ry * rz) + s * (-rx);
    R[6] = c1 * (rx * rz) + s * (-ry);    R[7] = c1 * (ry * rz) + s * rx;       R[8] = c + c1 * (rz * rz);
}

// cv::Mat::inv() of the 4x4 [R t; 0 0 0 1] — OpenCV hal LUImpl on [A | I], restated in full
// (mcm.cpp:619); returns the top 3x4 of the inverse.
__device__ inline void inv_rigid_lu(const Pose &in, Pose &out) {
    double A[16], b[16];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) A[i * 4 + j] = in.r[i * 3 + j];
        A[i * 4 + 3] = in.t[i];
    }
    A[12] = 0; A[13] = 0; A[14] = 0; A[15] = 1;
#pragma unroll
    for (int i = 0; i < 16; i++) b[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 4; i++) {
        int k = i;
        for (int j = i + 1; j < 4; j++)
            if (fabs(A[j * 4 + i]) > fabs(A[k * 4 + i])) k = j;
        if (k != i) {
            for (int j = i; j < 4; j++) { double t = A[i * 4 + j]; A[i * 4 + j] = A[k * 4 + j]; A[k * 4 + j] = t; }
            for (int j = 0; j < 4; j++) { double t = b[i * 4 + j]; b[i * 4 + j] = b[k * 4 + j]; b[k * 4 + j] = t; }
        }
        double d = -1 / A[i * 4 + i];
        for (int j = i + 1; j < 4; j++) {
            double alpha = A[j * 4 + i] * d;
            for (int kk = i + 1; kk < 4; kk++) A[j * 4 + kk] += alpha * A[i * 4 + kk];
            for (int kk = 0; kk < 4; kk++) b[j * 4 + kk] += alpha * b[i * 4 + kk];
        }
    }
    for (int i = 3; i >= 0; i--)
        for (int j = 0; j < 4; j++) {
            double s = b[i * 4 + j];
            for (int k = i + 1; k < 4; k++) s -= A[i * 4 + k] * b[k * 4 + j];
            b[i * 4 + j] = s / A[i * 4 + i];
        }
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) out.r[i * 3 + j] = b[i * 4 + j];
        out.t[i] = b[i * 4 + 3];
    }
}

// rotation part of A*B
__device__ __forceinline__ void compose_R(const double *Ra, const double *Rb, double *R) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R[i * 3 + j] = (Ra[i * 3 + 0] * Rb[0 * 3 + j] + Ra[i * 3 + 1] * Rb[1 * 3 + j]) + Ra[i * 3 + 2] * Rb[2 * 3 + j];
}
// translation part of A*B
__device__ __forceinline__ void compose_t(const double *Ra, const double *ta, const double *tb, double *t) {
#pragma unroll
    for (int i = 0; i < 3; i++)
        t[i] = ((Ra[i * 3 + 0] * tb[0] + Ra[i * 3 + 1] * tb[1]) + Ra[i * 3 + 2] * tb[2]) + ta[i];
}
// columns 0 and 1 of the rotation part of A*B (column 2 never reaches a projection: X has z = 0)
__device__ __forceinline__ void compose_R01(const double *Ra, const double *Rb, double *c0, double *c1) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        c0[i] = (Ra[i * 3 + 0] * Rb[0] + Ra[i * 3 + 1] * Rb[3]) + Ra[i * 3 + 2] * Rb[6];
        c1[i] = (Ra[i * 3 + 0] * Rb[1] + Ra[i * 3 + 1] * Rb[4]) + Ra[i * 3 + 2] * Rb[7];
    }
}

struct Intr { double fx, cx, fy, cy; };

// (K * T[0:3]) * X then the perspective divide, rounded to float32 (mcm.cpp:640-648).
// c0, c1: columns 0/1 of T's rotation, t: its translation, h: half marker size.
// out: x0 y0 x1 y1 x2 y2 x3 y3 for corners (-h,h) (h,h) (h,-h) (-h,-h)  (aruco marker.cpp:358-369).
__device__ __forceinline__ void project(const double *c0, const double *c1, const double *t, const Intr &k, double h, float *out) {
    // A = K*T34 with K = [fx 0 cx; 0 fy cy; 0 0 1]: the products with the structural zeros vanish exactly
    double a00 = k.fx * c0[0] + k.cx * c0[2], a01 = k.fx * c1[0] + k.cx * c1[2], a03 = k.fx * t[0] + k.cx * t[2];
    double a10 = k.fy * c0[1] + k.cy * c0[2], a11 = k.fy * c1[1] + k.cy * c1[2], a13 = k.fy * t[1] + k.cy * t[2];
    double a20 = c0[2], a21 = c1[2], a23 = t[2];
    double xa = a00 * h, xb = a01 * h, ya = a10 * h, yb = a11 * h, za = a20 * h, zb = a21 * h;
    // corner sums (A_i0*x + A_i1*y) + A_i3 ; (-a) + b == b - a and (-a) + (-b) == -(a + b) exactly
    double xs = xa + xb, xd = xb - xa, ys = ya + yb, yd = yb - ya, zs = za + zb, zd = zb - za;
    double X0 = xd + a03, X1 = xs + a03, X2 = a03 - xd, X3 = a03 - xs;
    double Y0 = yd + a13, Y1 = ys + a13, Y2 = a13 - yd, Y3 = a13 - ys;
    double Z0 = zd + a23, Z1 = zs + a23, Z2 = a23 - zd, Z3 = a23 - zs;
    out[0] = (float)(X0 / Z0); out[1] = (float)(Y0 / Z0);
    out[2] = (float)(X1 / Z1); out[3] = (float)(Y1 / Z1);
    out[4] = (float)(X2 / Z2); out[5] = (float)(Y2 / Z2);
    out[6] = (float)(X3 / Z3); out[7] = (float)(Y3 / Z3);
}

// MultiCamMapper.cpp:11-24 — note the float delta arithmetic
__device__ __forceinline__ double huber_weight(double sq, float delta) {
    if (sq == 0) return 1;
    float deltaSq = delta * delta;
    float delta2 = 2 * delta;
    double rho = (sq <= deltaSq) ? sq : delta2 * sqrt(sq) - deltaSq;
    return sqrt(rho / sq);
}

} // namespace aar
