/*
 * cv_shim.h — CPU ORACLE, TEST INFRASTRUCTURE ONLY (see mcm_oracle.cpp).
 *
 * The OpenCV arithmetic the reference reaches on the optimisation and initialisation paths, restated operation by
 * operation over a small 4x4 type: cv::Mat operator* (cv::gemm small-matrix path), cv::Mat::inv() (hal LUImpl),
 * cv::Rodrigues (both directions) and cv::undistortPoints.  Pinned bit-for-bit against cv2 4.13 known answers
 * (tests/golden/make_golden_cv2.py, tests/test_oracle_pin.py).  Shared by mcm_oracle.cpp and init_oracle.cpp.
 */
#ifndef AAR_ORACLE_CV_SHIM_H
#define AAR_ORACLE_CV_SHIM_H
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <utility>
#include "../include/aar_crsincos.h"

namespace {

/* ------------------------------------------------------------------ Mat4 shim ------------ */
struct M4 { double a[16]; };

M4 eye4() { M4 m; for (int i = 0; i < 16; i++) m.a[i] = (i % 5 == 0) ? 1.0 : 0.0; return m; }

/* cv::Mat operator* -> cv::gemm small-matrix path (inner length 2..4): every output element is
 * the left-associated sum a0*b0 + a1*b1 + a2*b2 + a3*b3 of separately rounded products.
 * Used at mcm.cpp:619-628, 640, 434, 704-705. */
M4 mul44(const M4 &A, const M4 &B) {
    M4 C;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            C.a[i * 4 + j] = A.a[i * 4 + 0] * B.a[0 * 4 + j] + A.a[i * 4 + 1] * B.a[1 * 4 + j] +
                             A.a[i * 4 + 2] * B.a[2 * 4 + j] + A.a[i * 4 + 3] * B.a[3 * 4 + j];
    return C;
}

/* cv::Mat::inv() on a 4x4 CV_64F = LU with partial pivoting on [A | I] (OpenCV hal LUImpl).
 * Used at mcm.cpp:294, 539, 619, 621, 704. */
M4 inv44(const M4 &Ain) {
    double A[16], b[16];
    std::memcpy(A, Ain.a, sizeof A);
    for (int i = 0; i < 16; i++) b[i] = (i % 5 == 0) ? 1.0 : 0.0;
    const int m = 4, n = 4;
    const double eps = DBL_EPSILON * 100;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j * 4 + i]) > std::fabs(A[k * 4 + i])) k = j;
        if (std::fabs(A[k * 4 + i]) < eps) { M4 z; std::memset(z.a, 0, sizeof z.a); return z; }
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * 4 + j], A[k * 4 + j]);
            for (int j = 0; j < n; j++) std::swap(b[i * 4 + j], b[k * 4 + j]);
        }
        double d = -1 / A[i * 4 + i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j * 4 + i] * d;
            for (int kk = i + 1; kk < m; kk++) A[j * 4 + kk] += alpha * A[i * 4 + kk];
            for (int kk = 0; kk < n; kk++) b[j * 4 + kk] += alpha * b[i * 4 + kk];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            double s = b[i * 4 + j];
            for (int k = i + 1; k < m; k++) s -= A[i * 4 + k] * b[k * 4 + j];
            b[i * 4 + j] = s / A[i * 4 + i];
        }
    M4 R; std::memcpy(R.a, b, sizeof b); return R;
}

/* 0 = libm (what cv::Rodrigues calls), 1 = aar_sincos (what the GPU runs); one variable for every translation unit of the library */
inline int &sincos_mode_ref() { static int mode = 0; return mode; }
#define g_sincos_mode (sincos_mode_ref())

/* cv::Rodrigues vector -> matrix (mcm.cpp:470, 693, 910-911). */
void rodrigues_vec2mat(const double r[3], double R[9]) {
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double c, s;
    if (g_sincos_mode == 0) { c = std::cos(theta); s = std::sin(theta); }
    else aar_sincos(theta, &s, &c);
    double c1 = 1. - c;
    double itheta = theta ? 1. / theta : 0.;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    /* R = cos(theta)*I + (1 - cos(theta))*r*rT + sin(theta)*[r_x] */
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
}

/* 3x3 one-sided Jacobi SVD (host-only helper for rodrigues_mat2vec; not parity critical:
 * R->r runs once before the solve, mcm.cpp:478, and both sides are fed the same z). */
void svd33(const double Ain[9], double U[9], double W[3], double Vt[9]) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(A, Ain, sizeof A);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; i++) {
                    alpha += A[i * 3 + p] * A[i * 3 + p];
                    beta += A[i * 3 + q] * A[i * 3 + q];
                    gamma += A[i * 3 + p] * A[i * 3 + q];
                }
                off = std::max(off, std::fabs(gamma) / std::sqrt(std::max(alpha * beta, 1e-300)));
                if (std::fabs(gamma) < 1e-300) continue;
                double zeta = (beta - alpha) / (2 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
                for (int i = 0; i < 3; i++) {
                    double ap = A[i * 3 + p], aq = A[i * 3 + q];
                    A[i * 3 + p] = cs * ap - sn * aq; A[i * 3 + q] = sn * ap + cs * aq;
                    double vp = V[i * 3 + p], vq = V[i * 3 + q];
                    V[i * 3 + p] = cs * vp - sn * vq; V[i * 3 + q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-16) break;
    }
    for (int j = 0; j < 3; j++) {
        double nrm = 0;
        for (int i = 0; i < 3; i++) nrm += A[i * 3 + j] * A[i * 3 + j];
        nrm = std::sqrt(nrm); W[j] = nrm;
        for (int i = 0; i < 3; i++) U[i * 3 + j] = nrm > 0 ? A[i * 3 + j] / nrm : (i == j);
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Vt[i * 3 + j] = V[j * 3 + i];
}

/* cv::Rodrigues matrix -> vector (mcm.cpp:478): SVD-orthonormalise then axis-angle. */
void rodrigues_mat2vec(const double Rin[9], double r[3]) {
    double U[9], W[3], Vt[9], R[9];
    svd33(Rin, U, W, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i * 3 + j] = U[i * 3 + 0] * Vt[0 * 3 + j] + U[i * 3 + 1] * Vt[1 * 3 + j] + U[i * 3 + 2] * Vt[2 * 3 + j];
    double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
    double s = std::sqrt((x * x + y * y + z * z) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) { x = y = z = 0; }
        else {
            double t;
            t = (R[0] + 1) * 0.5; x = std::sqrt(std::max(t, 0.));
            t = (R[4] + 1) * 0.5; y = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; z = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (std::fabs(x) < std::fabs(y) && std::fabs(x) < std::fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
            theta /= std::sqrt(x * x + y * y + z * z);
            x *= theta; y *= theta; z *= theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        x *= vth; y *= vth; z *= vth;
    }
    r[0] = x; r[1] = y; r[2] = z;
}

/* cv::undistortPoints(src CV_32FC2, K, dist(5), noArray(), P=K) for one point (mcm.cpp:570):
 * exactly 5 fixed-point iterations in double, output rounded to float32. */
void undistort_point(float u_in, float v_in, const double K[9], const double k[5], float *uo, float *vo) {
    double fx = K[0], fy = K[4], ifx = 1. / fx, ify = 1. / fy, cx = K[2], cy = K[5];
    /* RR = P * R with R = I, P = K (3x3 gemm: left-associated sums of products) */
    double RR[9];
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            RR[i * 3 + j] = K[i * 3 + 0] * I[0 * 3 + j] + K[i * 3 + 1] * I[1 * 3 + j] + K[i * 3 + 2] * I[2 * 3 + j];
    double x = u_in, y = v_in;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        double r2 = x * x + y * y;
        double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u_in - cx) * ifx; y = (v_in - cy) * ify; break; }
        double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
        double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0. * r2 + 0. * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    double xx = RR[0] * x + RR[1] * y + RR[2];
    double yy = RR[3] * x + RR[4] * y + RR[5];
    double ww = 1. / (RR[6] * x + RR[7] * y + RR[8]);
    x = xx * ww; y = yy * ww;
    *uo = (float)x; *vo = (float)y;
}

} // namespace
#endif
