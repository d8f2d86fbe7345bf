// aar_init.cuh — device side of the initialisation path (include/aar_init.h), sm_100a.
//
//   k_ippe            one thread per detection: aruco::solvePnP_ (3rdparty/aruco/aruco/ippe.cpp:118-219) — undistortion to
//                     normalised coordinates, the closed-form homography of a centred square, the two IPPE rotations, their
//                     translations and reprojection errors, IPPERot2vec + getRTMatrix(CV_32F) (ippe.cpp:40-97, 332-358)
//   k_rig_tables      T and inv(T) of every camera / marker of the rig (initializer.cpp:83-100)
//   k_build_object    candidate triples of init_object_transforms (initializer.cpp:74-115)
//   k_build_pairs     candidate triples of fill_transformation_sets (initializer.cpp:117-146)
//   k_consensus       find_best_transformation (initializer.cpp:156-205): candidate i's error is the sum over every j of the
//                     corner displacement of T2inv_j * T_i * T1inv_j; one thread per candidate, the j-side operands staged
//                     in shared memory tile by tile, sums in list order (bit-exact with the sequential reference)
//   k_consensus_pick  first minimum per list (strict '<' of the reference's scan)
// Arithmetic: -fmad=false, IEEE division and square root, every matrix product in cv::gemm's order (aar_device_math.cuh);
// poses are 3x4 (rotation | translation) because every matrix on this path has the exact last row [0 0 0 1].
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/aar_acos.h"
#include "aar_device_math.cuh"

namespace aar {

constexpr int TRI_DOUBLES = 36;          // candidate triple: T | T1inv | T2inv, each r[9] t[3]
constexpr int CS_THREADS = 128;          // candidates per CTA of k_consensus
constexpr int CS_TILE = 32;              // j-side entries staged per shared-memory tile

__device__ __forceinline__ void store_pose(double *dst, const Pose &p) {
#pragma unroll
    for (int i = 0; i < 9; i++) dst[i] = p.r[i];
#pragma unroll
    for (int i = 0; i < 3; i++) dst[9 + i] = p.t[i];
}
__device__ __forceinline__ void compose(const Pose &a, const Pose &b, Pose &o) { compose_R(a.r, b.r, o.r); compose_t(a.r, a.t, b.t, o.t); }

// IPPComputeTranslation (ippe.cpp:380-424); hs = half marker size as float, corners (-hs,hs) (hs,hs) (hs,-hs) (-hs,-hs), z = 0
__device__ inline void ippe_translation(float hs, const float *q, const double *R, double *t) {
    const double ATA00 = 4, ATA11 = 4;
    double ATA02 = 0, ATA12 = 0, ATA20 = 0, ATA21 = 0, ATA22 = 0, ATb0 = 0, ATb1 = 0, ATb2 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float mx = (i == 0 || i == 3) ? -hs : hs, my = (i < 2) ? hs : -hs, mz = 0.f;
        const double rx = R[0] * mx + R[1] * my + R[2] * mz;
        const double ry = R[3] * mx + R[4] * my + R[5] * mz;
        const double rz = R[6] * mx + R[7] * my + R[8] * mz;
        const double a2 = -q[2 * i], b2 = -q[2 * i + 1];
        ATA02 = ATA02 + a2; ATA12 = ATA12 + b2; ATA20 = ATA20 + a2; ATA21 = ATA21 + b2;
        ATA22 = ATA22 + a2 * a2 + b2 * b2;
        const double bx = (q[2 * i]) * rz - rx, by = (q[2 * i + 1]) * rz - ry;
        ATb0 = ATb0 + bx; ATb1 = ATb1 + by;
        ATb2 = ATb2 + a2 * bx + b2 * by;
    }
    const double detAInv = 1.0 / (ATA00 * ATA11 * ATA22 - ATA00 * ATA12 * ATA21 - ATA02 * ATA11 * ATA20);
    const double S00 = ATA11 * ATA22 - ATA12 * ATA21, S01 = ATA02 * ATA21, S02 = -ATA02 * ATA11;
    const double S10 = ATA12 * ATA20, S11 = ATA00 * ATA22 - ATA02 * ATA20, S12 = -ATA00 * ATA12;
    const double S20 = -ATA11 * ATA20, S21 = -ATA00 * ATA21, S22 = ATA00 * ATA11;
    t[0] = detAInv * (S00 * ATb0 + S01 * ATb1 + S02 * ATb2);
    t[1] = detAInv * (S10 * ATb0 + S11 * ATb1 + S12 * ATb2);
    t[2] = detAInv * (S20 * ATb0 + S21 * ATb1 + S22 * ATb2);
}

// IPPEvalReprojectionError (ippe.cpp:296-330): float sums of double products narrowed to float
__device__ inline float ippe_reproj_error(float hs, const float *q, const double *R, const double *t) {
    float err = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float mx = (i == 0 || i == 3) ? -hs : hs, my = (i < 2) ? hs : -hs, mz = 0.f;
        const float px = (float)(R[0] * mx) + (float)(R[1] * my) + (float)(R[2] * mz + t[0]);
        const float py = (float)(R[3] * mx) + (float)(R[4] * my) + (float)(R[5] * mz + t[1]);
        const float pz = (float)(R[6] * mx) + (float)(R[7] * my) + (float)(R[8] * mz + t[2]);
        const float dx = px / pz - q[2 * i], dy = py / pz - q[2 * i + 1];
        err = err + sqrtf(dx * dx + dy * dy);
    }
    return err;
}

// IPPERot2vec (ippe.cpp:332-358) -> cv::Rodrigues -> CV_32F (getRTMatrix, ippe.cpp:40-97) -> CV_64F (initializer.cpp:403)
__device__ inline void ippe_pose_out(const double *R, const double *t, double *dst) {
    const double trace = R[0] + R[4] + R[8];
    const double w_norm = aar_acos((trace - 1.0) / 2.0);
    double sn, cs;
    aar_sincos(w_norm, &sn, &cs);
    const double d = 1 / (2 * sn) * w_norm;
    double rx = 0, ry = 0, rz = 0;
    if (!(w_norm < DBL_EPSILON)) { rx = d * (R[7] - R[5]); ry = d * (R[2] - R[6]); rz = d * (R[3] - R[1]); }
    double R33[9];
    rodrigues(rx, ry, rz, R33);
#pragma unroll
    for (int i = 0; i < 9; i++) dst[i] = (double)(float)R33[i];
#pragma unroll
    for (int i = 0; i < 3; i++) dst[9 + i] = (double)(float)t[i];
}

__global__ void __launch_bounds__(128) k_ippe(long long n, const float *__restrict__ xy, const int *__restrict__ det_cam, const uint8_t *__restrict__ active,
                                              const double *__restrict__ K9, const double *__restrict__ dist5, float size, double threshold,
                                              double *__restrict__ est /* [n][2][12] */, float *__restrict__ err /* [n][2] */, uint8_t *__restrict__ ncand) {
    const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    if (!active[d]) { ncand[d] = 0; return; }
    const int cam = det_cam[d];
    const double *K = K9 + 9 * (size_t)cam, *k = dist5 + 5 * (size_t)cam;
    // cv::undistortPoints without R / P (ippe.cpp:164): normalised coordinates, float32
    float q[8];
    {
        const double fx = K[0], fy = K[4], ifx = 1. / fx, ify = 1. / fy, cx = K[2], cy = K[5];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float u = xy[8 * d + 2 * i], v = xy[8 * d + 2 * i + 1];
            double x = u, y = v;
            x = (x - cx) * ifx; y = (y - cy) * ify;
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = 1 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = ((double)u - cx) * ifx; y = ((double)v - cy) * ify; break; }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
                x = (x0 - deltaX) * icdist; y = (y0 - deltaY) * icdist;
            }
            q[2 * i] = (float)x; q[2 * i + 1] = (float)y;
        }
    }
    const float hs = size / 2.0f;
    // homographyFromSquarePoints (ippe.cpp:535-579)
    double H0, H1, H2, H3, H4, H5, H6, H7;
    {
        const double hl = hs;
        const double ax = -q[0], ay = -q[1], bx = -q[2], by = -q[3], cx = -q[4], cy = -q[5], dx = -q[6], dy = -q[7];
        const double di = -1 / (hl * (ax * by - bx * ay - ax * dy + bx * cy - cx * by + dx * ay + cx * dy - dx * cy));
        H0 = di * (ax * cx * by - bx * cx * ay - ax * dx * by + bx * dx * ay - ax * cx * dy + ax * dx * cy + bx * cx * dy - bx * dx * cy);
        H1 = di * (ax * bx * cy - ax * cx * by - ax * bx * dy + bx * dx * ay + ax * cx * dy - cx * dx * ay - bx * dx * cy + cx * dx * by);
        H2 = di * hl * (ax * bx * cy - bx * cx * ay - ax * bx * dy + ax * dx * by - ax * dx * cy + cx * dx * ay + bx * cx * dy - cx * dx * by);
        H3 = di * (ax * by * cy - bx * ay * cy - ax * by * dy + bx * ay * dy - cx * ay * dy + dx * ay * cy + cx * by * dy - dx * by * cy);
        H4 = di * (bx * ay * cy - cx * ay * by - ax * by * dy + dx * ay * by + ax * cy * dy - dx * ay * cy - bx * cy * dy + cx * by * dy);
        H5 = di * hl * (ax * by * cy - cx * ay * by - bx * ay * dy + dx * ay * by - ax * cy * dy + cx * ay * dy + bx * cy * dy - dx * by * cy);
        H6 = -di * (ax * cy - cx * ay - ax * dy - bx * cy + cx * by + dx * ay + bx * dy - dx * by);
        H7 = di * (ax * by - bx * ay - ax * cy + cx * ay + bx * dy - dx * by - cx * dy + dx * cy);
    }
    // IPPComputeRotations (ippe.cpp:426-533)
    double Ra[9], Rb[9];
    {
        const double j00 = H0 - H6 * H2, j01 = H1 - H7 * H2, j10 = H3 - H6 * H5, j11 = H4 - H7 * H5, p = H2, qq = H5;
        const double s = sqrt(p * p + qq * qq + 1), t = sqrt(p * p + qq * qq);
        const double costh = 1 / s, sinth = sqrt(1 - 1 / (s * s));
        const double k0 = p / t, k1 = qq / t, k0s = k0 * k0, k1s = k1 * k1;
        double rv[9];
        rv[0] = (costh - 1) * k0s + 1;  rv[1] = k0 * k1 * (costh - 1);  rv[2] = k0 * sinth;
        rv[3] = k0 * k1 * (costh - 1);  rv[4] = (costh - 1) * k1s + 1;  rv[5] = k1 * sinth;
        rv[6] = -k0 * sinth;            rv[7] = -k1 * sinth;            rv[8] = (costh - 1) * (k0s + k1s) + 1;
        const double b00 = rv[0] - p * rv[6], b01 = rv[1] - p * rv[7], b10 = rv[3] - qq * rv[6], b11 = rv[4] - qq * rv[7];
        const double dtinv = 1.0 / ((b00 * b11 - b01 * b10));
        const double bi00 = dtinv * b11, bi01 = -dtinv * b01, bi10 = -dtinv * b10, bi11 = dtinv * b00;
        const double a00 = bi00 * j00 + bi01 * j10, a01 = bi00 * j01 + bi01 * j11, a10 = bi10 * j00 + bi11 * j10, a11 = bi10 * j01 + bi11 * j11;
        const double ata00 = a00 * a00 + a01 * a01, ata01 = a00 * a10 + a01 * a11, ata11 = a10 * a10 + a11 * a11;
        const double gamma = sqrt(0.5 * (ata00 + ata11 + sqrt((ata00 - ata11) * (ata00 - ata11) + 4.0 * ata01 * ata01)));
        const double r00 = a00 / gamma, r01 = a01 / gamma, r10 = a10 / gamma, r11 = a11 / gamma;
        const double b0 = sqrt(-(r00 * r00) - r10 * r10 + 1);
        double b1 = sqrt(-(r01 * r01) - r11 * r11 + 1);
        const double sp = (-r00 * r01 - r10 * r11);
        if (sp < 0) b1 = -b1;
        const double u1 = b1 * r10 - b0 * r11, v1 = b0 * r01 - b1 * r00, w = r00 * r11 - r01 * r10;
        const double u2 = b0 * r11 - b1 * r10, v2 = b1 * r00 - b0 * r01;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double x = rv[3 * i], y = rv[3 * i + 1], z = rv[3 * i + 2];
            Ra[3 * i + 0] = (r00) * x + (r10) * y + (b0) * z;
            Ra[3 * i + 1] = (r01) * x + (r11) * y + (b1) * z;
            Ra[3 * i + 2] = u1 * x + v1 * y + w * z;
            Rb[3 * i + 0] = (r00) * x + (r10) * y + (-b0) * z;
            Rb[3 * i + 1] = (r01) * x + (r11) * y + (-b1) * z;
            Rb[3 * i + 2] = u2 * x + v2 * y + w * z;
        }
    }
    double ta[3], tb[3];
    ippe_translation(hs, q, Ra, ta);
    ippe_translation(hs, q, Rb, tb);
    const float ea = ippe_reproj_error(hs, q, Ra, ta), eb = ippe_reproj_error(hs, q, Rb, tb);
    double *o = est + (size_t)d * 24;
    float e0, e1;
    if (ea < eb) { ippe_pose_out(Ra, ta, o); ippe_pose_out(Rb, tb, o + 12); e0 = ea; e1 = eb; }
    else         { ippe_pose_out(Rb, tb, o); ippe_pose_out(Ra, ta, o + 12); e0 = eb; e1 = ea; }
    err[2 * d] = e0; err[2 * d + 1] = e1;
    ncand[d] = ((double)e1 / (double)e0 < threshold) ? 2 : 1;             // initializer.cpp:408
}

// per camera / marker of the rig: [T | inv(T)], identity for ids without a transform (initializer.cpp:83-100)
__global__ void k_rig_tables(int n, const double *__restrict__ T12, const uint8_t *__restrict__ has, double *__restrict__ tab /* [n][24] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pose p, q;
    if (has[i]) { load_pose(p, T12 + 12 * (size_t)i); inv_rigid_lu(p, q); }
    else {
#pragma unroll
        for (int k = 0; k < 9; k++) p.r[k] = q.r[k] = (k % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) p.t[k] = q.t[k] = 0.0;
    }
    store_pose(tab + 24 * (size_t)i, p); store_pose(tab + 24 * (size_t)i + 12, q);
}

// fill_transformation_set (initializer.cpp:74-115): src = 2 * detection + candidate
__global__ void __launch_bounds__(128) k_build_object(long long n, const int *__restrict__ src, const double *__restrict__ est, const int *__restrict__ det_cam,
                                                      const int *__restrict__ det_midx /* marker index of the detection */, const double *__restrict__ cam_tab,
                                                      const double *__restrict__ mk_tab, double *__restrict__ tri) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int s = src[e], d = s >> 1;
    Pose T_mc, T_cm, T_cr, T_rc, T_mr, T_rm, a, b;
    load_pose(T_mc, est + (size_t)s * 12);
    inv_rigid_lu(T_mc, T_cm);
    const double *ct = cam_tab + 24 * (size_t)det_cam[d], *mt = mk_tab + 24 * (size_t)det_midx[d];
    load_pose(T_cr, ct); load_pose(T_rc, ct + 12); load_pose(T_mr, mt); load_pose(T_rm, mt + 12);
    compose(T_cr, T_mc, a); compose(a, T_rm, b);                     // T = T_cr * T_mc * T_rm
    double *o = tri + (size_t)e * TRI_DOUBLES;
    store_pose(o, b);
    compose(T_mr, T_cm, a);                                          // T1inv = T_mr * T_cm
    store_pose(o + 12, a);
    store_pose(o + 24, T_rc);                                        // T2inv = T_rc
}

// fill_transformation_sets (initializer.cpp:117-146): one candidate per (source a, source b)
__global__ void __launch_bounds__(128) k_build_pairs(long long n, int cams, const int2 *__restrict__ src, const double *__restrict__ est, double *__restrict__ tri) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int2 s = src[e];
    Pose p1, p2, i1, i2, t;
    load_pose(p1, est + (size_t)s.x * 12); load_pose(p2, est + (size_t)s.y * 12);
    double *o = tri + (size_t)e * TRI_DOUBLES;
    if (cams) { inv_rigid_lu(p1, i1); inv_rigid_lu(p2, i2); compose(p2, i1, t); store_pose(o, t); store_pose(o + 12, p1); store_pose(o + 24, i2); }
    else      { inv_rigid_lu(p1, i1); inv_rigid_lu(p2, i2); compose(i2, p1, t); store_pose(o, t); store_pose(o + 12, i1); store_pose(o + 24, p2); }
}

struct ConsJob { int seg, first; };       // one CTA: candidates [first, first + CS_THREADS) of list `seg`

__global__ void __launch_bounds__(CS_THREADS) k_consensus(const ConsJob *__restrict__ jobs, const long long *__restrict__ seg_begin, const double *__restrict__ tri,
                                                          double h, double *__restrict__ part_val, long long *__restrict__ part_idx) {
    __shared__ double sj[CS_TILE * 24];
    __shared__ double s_val[CS_THREADS / 32];
    __shared__ long long s_idx[CS_THREADS / 32];
    const ConsJob job = jobs[blockIdx.x];
    const long long e0 = seg_begin[job.seg], e1 = seg_begin[job.seg + 1];
    const long long m = e1 - e0;
    const long long ii = (long long)job.first + threadIdx.x;
    const bool live = ii < m;
    Pose Ti;
    if (live) load_pose(Ti, tri + (size_t)(e0 + ii) * TRI_DOUBLES);
    else { for (int k = 0; k < 9; k++) Ti.r[k] = 0; for (int k = 0; k < 3; k++) Ti.t[k] = 0; }
    double curr = 0;
    for (long long j0 = 0; j0 < m; j0 += CS_TILE) {
        const int nj = (int)min((long long)CS_TILE, m - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < nj * 24; e += CS_THREADS) sj[e] = tri[(size_t)(e0 + j0 + e / 24) * TRI_DOUBLES + 12 + e % 24];
        __syncthreads();
        if (live)
            for (int j = 0; j < nj; j++) {
                const double *R1 = sj + j * 24, *t1 = R1 + 9, *R2 = R1 + 12, *t2 = R1 + 21;
                // p2 = ((T2inv_j * T_i) * T1inv_j) * points ; points = [x; y; 0; 1] with x, y = +-h
                double Ra[9], ta[3], c0[3], c1[3], tb[3];
                compose_R(R2, Ti.r, Ra); compose_t(R2, t2, Ti.t, ta);
                compose_R01(Ra, R1, c0, c1); compose_t(Ra, ta, t1, tb);
                double xa[3], ya[3];
#pragma unroll
                for (int r = 0; r < 3; r++) { xa[r] = c0[r] * h; ya[r] = c1[r] * h; }
                double e = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const double sx = (c == 0 || c == 3) ? -1.0 : 1.0, sy = (c < 2) ? 1.0 : -1.0;
                    // (B_r0 * x + B_r1 * y) + t_r : products with -h are the exact negatives of the products with h
                    const double p0 = (sx * xa[0] + sy * ya[0]) + tb[0], p1 = (sx * xa[1] + sy * ya[1]) + tb[1], p2 = (sx * xa[2] + sy * ya[2]) + tb[2];
                    const double d0 = sx * h - p0, d1 = sy * h - p1, d2 = 0.0 - p2;
                    const double ec = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
                    e = (c == 0) ? ec : e + ec;                      // cv::sum: ((e0 + e1) + e2) + e3
                }
                curr += e;
            }
    }
    // first minimum among the live candidates whose error is below DBL_MAX (the reference's initial min_error)
    double v = (live && curr < DBL_MAX) ? curr : DBL_MAX;
    long long idx = (live && curr < DBL_MAX) ? ii : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const long long oi = __shfl_down_sync(0xffffffffu, idx, o);
        if (oi >= 0 && (idx < 0 || ov < v || (ov == v && oi < idx))) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = v; s_idx[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CS_THREADS / 32; w++)
            if (s_idx[w] >= 0 && (idx < 0 || s_val[w] < v)) { v = s_val[w]; idx = s_idx[w]; }
        part_val[blockIdx.x] = v; part_idx[blockIdx.x] = idx;
    }
}

// per list: the first minimum over its CTAs (ascending candidate ranges), and the winner's T
__global__ void k_consensus_pick(int nseg, const long long *__restrict__ job_begin, const long long *__restrict__ seg_begin, const double *__restrict__ part_val,
                                 const long long *__restrict__ part_idx, const double *__restrict__ tri, long long *__restrict__ best, double *__restrict__ weight,
                                 double *__restrict__ best_T /* [nseg][12] */) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    double v = DBL_MAX; long long idx = -1;
    for (long long j = job_begin[s]; j < job_begin[s + 1]; j++)
        if (part_idx[j] >= 0 && (idx < 0 || part_val[j] < v)) { v = part_val[j]; idx = part_idx[j]; }
    best[s] = idx; weight[s] = v;
    if (idx >= 0) for (int k = 0; k < 12; k++) best_T[(size_t)s * 12 + k] = tri[(size_t)(seg_begin[s] + idx) * TRI_DOUBLES + k];
}

} // namespace aar
