// aar_host_math.h — host-only helpers of the C-ABI layer (no device code, no OpenCV).
// R -> r of cv::Rodrigues as used by MultiCamMapper::transformation_mat2vec
// (/root/reference/libs/multicam_mapper.cpp:475-486).  It runs once per pose before the solve and
// is not on the device path; the 3x3 SVD is a cyclic one-sided Jacobi (Hestenes) iteration.
#pragma once
#include <algorithm>
#include <cmath>

namespace aar_host {

// Polar factor U*Vt of a 3x3 matrix via one-sided Jacobi on the columns.
inline void nearest_rotation(const double *Rin, double *Rout) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) A[i] = Rin[i];
    for (int sweep = 0; sweep < 64; sweep++) {
        double worst = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double app = 0, aqq = 0, apq = 0;
                for (int i = 0; i < 3; i++) { app += A[i * 3 + p] * A[i * 3 + p]; aqq += A[i * 3 + q] * A[i * 3 + q]; apq += A[i * 3 + p] * A[i * 3 + q]; }
                double denom = std::sqrt(std::max(app * aqq, 1e-300));
                worst = std::max(worst, std::fabs(apq) / denom);
                if (std::fabs(apq) <= 1e-300) continue;
                double zeta = (aqq - app) / (2 * apq);
                double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int i = 0; i < 3; i++) {
                    double x = A[i * 3 + p], y = A[i * 3 + q];
                    A[i * 3 + p] = c * x - s * y; A[i * 3 + q] = s * x + c * y;
                    x = V[i * 3 + p]; y = V[i * 3 + q];
                    V[i * 3 + p] = c * x - s * y; V[i * 3 + q] = s * x + c * y;
                }
            }
        if (worst < 1e-17) break;
    }
    // A = U*diag(w): normalise the columns to get U, then R = U * V^T
    double U[9];
    for (int j = 0; j < 3; j++) {
        double n = std::sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
        for (int i = 0; i < 3; i++) U[i * 3 + j] = n > 0 ? A[i * 3 + j] / n : (i == j ? 1.0 : 0.0);
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rout[i * 3 + j] = U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
}

// cv::Rodrigues(matrix -> vector)
inline void rotation_to_vector(const double *Rin, double *rv) {
    double R[9];
    nearest_rotation(Rin, R);
    double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
    double s = std::sqrt((x * x + y * y + z * z) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : (c < -1. ? -1. : c);
    double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) { x = y = z = 0; }
        else {
            x = std::sqrt(std::max((R[0] + 1) * 0.5, 0.));
            y = std::sqrt(std::max((R[4] + 1) * 0.5, 0.)) * (R[1] < 0 ? -1. : 1.);
            z = std::sqrt(std::max((R[8] + 1) * 0.5, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (std::fabs(x) < std::fabs(y) && std::fabs(x) < std::fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
            double k = theta / std::sqrt(x * x + y * y + z * z);
            x *= k; y *= k; z *= k;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        x *= vth; y *= vth; z *= vth;
    }
    rv[0] = x; rv[1] = y; rv[2] = z;
}

} // namespace aar_host
