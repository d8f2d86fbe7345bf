/*
 * mcm_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (automatic-ar_b200/) never links or calls it.
 *
 * A from-scratch restatement of the optimisation part of the reference's MultiCamMapper
 * (/root/reference/libs/multicam_mapper.cpp, cited per function below as mcm.cpp:LINES) with
 * a small Mat4 shim standing in for OpenCV.  The OpenCV arithmetic the reference reaches on
 * this path (cv::Rodrigues, cv::Mat::inv, cv::gemm small-matrix path, cv::undistortPoints)
 * is restated operation by operation and pinned bit-for-bit against cv2 4.13 known answers
 * in tests/golden/ (see tests/golden/make_golden_cv2.py).  The reference has no tests, golden
 * vectors or fixtures of its own and cannot be built here (OpenCV C++ is absent), so apart
 * from those cv2 vectors and the verbatim sparselevmarq.h below: PARITY UNPINNED.
 *
 * Two builds (oracle/Makefile):
 *   oracle/_ref/libaar_oracle_ref.so  -DAAR_ORACLE_WITH_REFERENCE_SLM : drives the UNMODIFIED
 *        reference solver /root/reference/libs/sparselevmarq.h with the vendored Eigen 3.2.92
 *        (included from where they lie; nothing is copied into this repo).
 *   oracle/libaar_oracle.so           : no reference includes; the LM loop is the restatement
 *        of sparselevmarq.h:237-249, 348-472 in lm_port() with an exact block-arrow solver.
 *
 * Build flags matter: -ffp-contract=off (no FMA contraction) so that every product/sum is
 * the separately rounded IEEE operation OpenCV's baseline build performs.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <utility>
#include <vector>
#include <omp.h>

#include "../include/aar_crsincos.h"
#include "../include/aar_analytic.h"

#ifdef AAR_ORACLE_WITH_REFERENCE_SLM
#include <Eigen/Sparse>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <sstream>
/* expose mu / currErr of the reference solver for the per-iteration trace (test infra only) */
#define private public
#include "sparselevmarq.h"
#undef private
#endif

#include "cv_shim.h"

namespace {

/* ------------------------------------------------------------------ MCM restatement ------ */
struct Marker { int id; float x[4], y[4]; }; /* aruco::Marker = id + 4 cv::Point2f (marker.h:47-59) */

typedef std::map<int, std::map<int, std::vector<Marker>>> FCM;
typedef std::map<int, std::map<int, std::map<int, std::pair<Marker, size_t>>>> Index3;

/* MultiCamMapper::MatArray (multicam_mapper.h:86-136): id -> index = rank of the id */
struct MatArray {
    std::vector<M4> v;
    std::map<int, int> m;
    std::vector<int> id;
    void assign(const std::map<int, M4> &ma) { /* operator=(std::map) h:95-105 */
        v.resize(ma.size()); id.resize(ma.size()); m.clear();
        size_t index = 0;
        for (auto it = ma.begin(); it != ma.end(); ++it) { m[it->first] = (int)index; id[index] = it->first; v[index++] = it->second; }
    }
    void init(const MatArray &ma) { m = ma.m; id = ma.id; v.assign(ma.v.size(), eye4()); } /* h:122-126 */
    const M4 &at_id(int i) const { return v[m.at(i)]; }
};
struct MatArrays { /* h:138-179; cam_mats hold the 3x3 in the top-left of an M4, dist in a[0..4] */
    MatArray object_to_global, transforms_to_root_marker, transforms_to_root_cam, transforms_to_local_cam, cam_mats, dist_coeffs;
};
struct Config { bool optimize_cam_poses = true, optimize_object_poses = true, optimize_marker_poses = true, optimize_cam_intrinsics = true; };
enum param_type { camera, marker, object, intrinsics };

/* stands in for the reference's deep-copied, one-entry-modified MatArrays ma_a / ma_s
 * (mcm.cpp:804-813): "use matrix *T instead of entry `index` of group `type`". */
struct Override { int type; size_t index; const M4 *T; const M4 *dist; };

typedef std::vector<double> eVec;
struct Triplet { int row, col; double val; };

struct Mcm {
    bool with_huber = false;
    /* analytic-Jacobian / full-FP64 variant (include/aar_analytic.h): residual in double with no float32
     * rounding, Jacobian by differentiation instead of central differences.  NOT the reference's arithmetic. */
    bool analytic = false;
    MatArrays mat_arrays;
    Config config;
    double J_delta = 0.001;
    double marker_size = 0;
    size_t root_cam = 0, root_marker = 0, num_cameras = 0, num_markers = 0, num_frames = 0, num_point_xys = 0, num_vars = 0;
    float hubberDelta = 2.5;
    M4 marker_points_3d_mat; /* 4x4: rows x,y,z,1 ; columns = corners */
    std::map<int, M4> marker_object_points_3d_mats;
    FCM frame_cam_markers;
    Index3 cam_marker_frame, marker_cam_frame, frame_cam_marker;
    eVec io_vec;
    /* LM params (mcm.cpp:326-330) */
    int maxIters = 10000; double minError = 1e-5, min_step_error_diff = 0, min_average_step_error_diff = 1e-4, tau = 1, der_epsilon = 1e-3;

    /* mcm.cpp:239-250 */
    size_t get_num_vars(const Config &conf) const {
        size_t n = 0;
        if (conf.optimize_cam_poses) n += (num_cameras - 1) * 6;
        if (conf.optimize_marker_poses) n += (num_markers - 1) * 6;
        if (conf.optimize_object_poses) n += num_frames * 6;
        if (conf.optimize_cam_intrinsics) n += num_cameras * 9;
        return n;
    }
    void set_config(const Config &c) { config = c; num_vars = get_num_vars(config); } /* :81-84 */

    /* mcm.cpp:261-270 + aruco marker.cpp:358-369 (half size computed in float) */
    void init_marker_points_3d() {
        float halfSize = (float)marker_size / 2.f;
        const float px[4] = {-halfSize, halfSize, halfSize, -halfSize};
        const float py[4] = {halfSize, halfSize, -halfSize, -halfSize};
        for (int i = 0; i < 4; i++) {
            marker_points_3d_mat.a[0 * 4 + i] = px[i];
            marker_points_3d_mat.a[1 * 4 + i] = py[i];
            marker_points_3d_mat.a[2 * 4 + i] = 0;
            marker_points_3d_mat.a[3 * 4 + i] = 1;
        }
    }

    /* mcm.cpp:281-335 (8-argument init); K/dist indexed by camera id like cam_configs[cam_id] */
    void init(size_t root_c, const std::map<int, M4> &T_to_root_cam, size_t root_m, const std::map<int, M4> &T_to_root_marker,
              const std::map<int, M4> &object_poses, const FCM &fcm, float m_size, const std::map<int, M4> &Ks, const std::map<int, M4> &dists) {
        root_cam = root_c; root_marker = root_m; frame_cam_markers = fcm;
        num_cameras = T_to_root_cam.size();
        mat_arrays.transforms_to_root_cam.assign(T_to_root_cam);
        mat_arrays.transforms_to_local_cam.init(mat_arrays.transforms_to_root_cam);
        for (size_t i = 0; i < num_cameras; i++) mat_arrays.transforms_to_local_cam.v[i] = inv44(mat_arrays.transforms_to_root_cam.v[i]);
        num_markers = T_to_root_marker.size();
        mat_arrays.transforms_to_root_marker.assign(T_to_root_marker);
        num_frames = object_poses.size();
        mat_arrays.object_to_global.assign(object_poses);
        num_vars = (num_frames + num_cameras - 1 + num_markers - 1) * 6 + num_cameras * 9;
        fill_iteration_arrays(); /* :304 — BEFORE the undistortion at :322 */
        std::map<int, M4> km, dm;
        for (auto &p : T_to_root_cam) { km[p.first] = Ks.at(p.first); dm[p.first] = dists.at(p.first); }
        mat_arrays.cam_mats.assign(km);
        mat_arrays.dist_coeffs.assign(dm);
        remove_distortions();
        marker_size = m_size;
        init_marker_points_3d();
    }
    /* mcm.cpp:272-279 (re-init for tracking) */
    void init_track(const std::map<int, M4> &object_poses, const FCM &fcm) {
        num_frames = object_poses.size();
        mat_arrays.object_to_global.assign(object_poses);
        frame_cam_markers = fcm;
        num_vars = get_num_vars(config);
        fill_iteration_arrays();
        remove_distortions();
    }

    /* mcm.cpp:345-377 */
    void fill_iteration_arrays() {
        num_point_xys = 0;
        cam_marker_frame.clear(); marker_cam_frame.clear(); frame_cam_marker.clear();
        for (auto frame_it = frame_cam_markers.begin(); frame_it != frame_cam_markers.end(); ++frame_it) {
            int frame_id = frame_it->first;
            auto &cam_markers = frame_it->second;
            for (auto cam_it = cam_markers.begin(); cam_it != cam_markers.end();) {
                int cam_id = cam_it->first;
                if (mat_arrays.transforms_to_root_cam.m.find(cam_id) == mat_arrays.transforms_to_root_cam.m.end()) { cam_it = cam_markers.erase(cam_it); continue; }
                auto &markers = cam_it->second;
                for (auto marker_it = markers.begin(); marker_it != markers.end();) {
                    int marker_id = marker_it->id;
                    if (mat_arrays.transforms_to_root_marker.m.find(marker_id) == mat_arrays.transforms_to_root_marker.m.end()) { marker_it = markers.erase(marker_it); continue; }
                    cam_marker_frame[cam_id][marker_id][frame_id] = std::make_pair(*marker_it, num_point_xys);
                    marker_cam_frame[marker_id][cam_id][frame_id] = std::make_pair(*marker_it, num_point_xys);
                    frame_cam_marker[frame_id][cam_id][marker_id] = std::make_pair(*marker_it, num_point_xys);
                    num_point_xys += 8;
                    ++marker_it;
                }
                ++cam_it;
            }
        }
    }

    /* mcm.cpp:554-578 */
    void remove_distortions() {
        for (auto &f : frame_cam_markers)
            for (auto &c : f.second) {
                int cam_id = c.first;
                const M4 &K4 = mat_arrays.cam_mats.at_id(cam_id);
                const M4 &d4 = mat_arrays.dist_coeffs.at_id(cam_id);
                double K[9] = {K4.a[0], K4.a[1], K4.a[2], K4.a[4], K4.a[5], K4.a[6], K4.a[8], K4.a[9], K4.a[10]};
                for (auto &mk : c.second)
                    for (int j = 0; j < 4; j++) undistort_point(mk.x[j], mk.y[j], K, d4.a, &mk.x[j], &mk.y[j]);
            }
    }

    /* mcm.cpp:463-473 */
    void vec2transformation_mat(size_t &vec_index, const eVec &vec, M4 &mat) const {
        double rv[3], R[9];
        for (int i = 0; i < 3; i++) { rv[i] = vec[vec_index + i]; mat.a[i * 4 + 3] = vec[vec_index + 3 + i]; }
        rodrigues_vec2mat(rv, R);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) mat.a[i * 4 + j] = R[i * 3 + j];
        vec_index += 6;
    }
    /* mcm.cpp:475-486 */
    void transformation_mat2vec(const M4 &mat, size_t &vec_index) {
        double R[9], rv[3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i * 3 + j] = mat.a[i * 4 + j];
        rodrigues_mat2vec(R, rv);
        for (int i = 0; i < 3; i++) { io_vec[vec_index + i] = rv[i]; io_vec[vec_index + 3 + i] = mat.a[i * 4 + 3]; }
        vec_index += 6;
    }
    /* mcm.cpp:445-461, 488-522 */
    void mats2eVec(const Config &conf) {
        io_vec.assign(get_num_vars(conf), 0.0);
        size_t vi = 0;
        if (conf.optimize_cam_poses)
            for (size_t i = 0; i < num_cameras; i++) { if ((int)i == mat_arrays.transforms_to_root_cam.m.at((int)root_cam)) continue; transformation_mat2vec(mat_arrays.transforms_to_root_cam.v[i], vi); }
        if (conf.optimize_marker_poses)
            for (size_t i = 0; i < num_markers; i++) { if ((int)i == mat_arrays.transforms_to_root_marker.m.at((int)root_marker)) continue; transformation_mat2vec(mat_arrays.transforms_to_root_marker.v[i], vi); }
        if (conf.optimize_object_poses)
            for (size_t i = 0; i < num_frames; i++) transformation_mat2vec(mat_arrays.object_to_global.v[i], vi);
        if (conf.optimize_cam_intrinsics)
            for (size_t i = 0; i < num_cameras; i++) {
                const M4 &K = mat_arrays.cam_mats.v[i];
                io_vec[vi++] = K.a[0]; io_vec[vi++] = K.a[2]; io_vec[vi++] = K.a[5]; io_vec[vi++] = K.a[6];
                for (int j = 0; j < 5; j++) io_vec[vi++] = mat_arrays.dist_coeffs.v[i].a[j];
            }
    }
    /* mcm.cpp:595-606 with :524-552, 580-593 */
    void eVec2Mats(const eVec &e, MatArrays &ma) const {
        size_t vi = 0;
        if (config.optimize_cam_poses)
            for (size_t i = 0; i < num_cameras; i++) {
                ma.transforms_to_root_cam.v[i] = eye4(); ma.transforms_to_local_cam.v[i] = eye4();
                if ((int)i == ma.transforms_to_root_cam.m.at((int)root_cam)) continue;
                vec2transformation_mat(vi, e, ma.transforms_to_root_cam.v[i]);
                ma.transforms_to_local_cam.v[i] = inv44(ma.transforms_to_root_cam.v[i]);
            }
        if (config.optimize_marker_poses)
            for (size_t i = 0; i < num_markers; i++) {
                ma.transforms_to_root_marker.v[i] = eye4();
                if ((int)i == ma.transforms_to_root_marker.m.at((int)root_marker)) continue;
                vec2transformation_mat(vi, e, ma.transforms_to_root_marker.v[i]);
            }
        if (config.optimize_object_poses)
            for (size_t i = 0; i < num_frames; i++) { ma.object_to_global.v[i] = eye4(); vec2transformation_mat(vi, e, ma.object_to_global.v[i]); }
        if (config.optimize_cam_intrinsics)
            for (size_t i = 0; i < num_cameras; i++) {
                M4 K; std::memset(K.a, 0, sizeof K.a); K.a[0] = K.a[5] = K.a[10] = 1; /* eye(3,3) in the 4x4 container */
                K.a[0] = e[vi++]; K.a[2] = e[vi++]; K.a[5] = e[vi++]; K.a[6] = e[vi++];
                ma.cam_mats.v[i] = K;
                M4 d; std::memset(d.a, 0, sizeof d.a);
                for (int j = 0; j < 5; j++) d.a[j] = e[vi++];
                ma.dist_coeffs.v[i] = d;
            }
    }
    /* MatArrays::init(ma, conf) h:145-161 */
    void ma_init(MatArrays &ma) const {
        if (config.optimize_cam_poses) { ma.transforms_to_root_cam.init(mat_arrays.transforms_to_root_cam); ma.transforms_to_local_cam.init(mat_arrays.transforms_to_local_cam); }
        if (config.optimize_marker_poses) ma.transforms_to_root_marker.init(mat_arrays.transforms_to_root_marker);
        if (config.optimize_object_poses) ma.object_to_global.init(mat_arrays.object_to_global);
        if (config.optimize_cam_intrinsics) { ma.cam_mats.init(mat_arrays.cam_mats); ma.dist_coeffs.init(mat_arrays.dist_coeffs); }
    }

    /* mcm.cpp:608-649.  NOTE cam_mat layout: 3x3 in rows 0..2 / cols 0..2 of the M4 with stride 4. */
    void project_marker(const MatArrays &ma, const Override *ov, size_t frame_id, size_t marker_id, size_t cam_id, float px[4], float py[4]) const {
        M4 transform;
        auto pick = [&](int type, const MatArray &opt, const MatArray &fixed, bool optimized, size_t id) -> const M4 & {
            const MatArray &src = optimized ? opt : fixed;
            size_t idx = (size_t)src.m.at((int)id);
            if (ov && optimized && ov->type == type && ov->index == idx) return *ov->T;
            return src.v[idx];
        };
        transform = pick(object, ma.object_to_global, mat_arrays.object_to_global, config.optimize_object_poses, frame_id);
        if (cam_id != root_cam)
            transform = mul44(inv44(pick(camera, ma.transforms_to_root_cam, mat_arrays.transforms_to_root_cam, config.optimize_cam_poses, cam_id)), transform);
        if (marker_id != root_marker)
            transform = mul44(transform, pick(marker, ma.transforms_to_root_marker, mat_arrays.transforms_to_root_marker, config.optimize_marker_poses, marker_id));
        const M4 &cam_mat = pick(intrinsics, ma.cam_mats, mat_arrays.cam_mats, config.optimize_cam_intrinsics, cam_id);
        /* cam_mat(3x3) * transform.rowRange(0,3)(3x4) * marker_points_3d_mat(4x4), left to right */
        double KT[12];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++)
                KT[i * 4 + j] = cam_mat.a[i * 4 + 0] * transform.a[0 * 4 + j] + cam_mat.a[i * 4 + 1] * transform.a[1 * 4 + j] + cam_mat.a[i * 4 + 2] * transform.a[2 * 4 + j];
        double P[12];
        const double *X = marker_points_3d_mat.a;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++)
                P[i * 4 + j] = KT[i * 4 + 0] * X[0 * 4 + j] + KT[i * 4 + 1] * X[1 * 4 + j] + KT[i * 4 + 2] * X[2 * 4 + j] + KT[i * 4 + 3] * X[3 * 4 + j];
        for (int c = 0; c < 4; c++) { px[c] = (float)(P[0 * 4 + c] / P[2 * 4 + c]); py[c] = (float)(P[1 * 4 + c] / P[2 * 4 + c]); }
    }

    /* mcm.cpp:11-24 */
    static double hubberMono(double e, float delta) {
        float deltaSq = delta * delta;
        float delta2 = 2 * delta;
        if (e <= deltaSq) return e;
        return delta2 * std::sqrt(e) - deltaSq;
    }
    static double getHubberMonoWeight(double SqErr, float delta) {
        if (SqErr == 0) return 1;
        return std::sqrt(hubberMono(SqErr, delta) / SqErr);
    }


    /* ---- analytic variant: the per-observation inputs of aar_an_observation (include/aar_analytic.h) */
    struct AnObs { double Ri[9], ti[3], dRc[27], tc[3], Ro[9], to[3], dRo[27], Rm[9], tm[3], dRm[27], fx, cx, fy, cy, h; bool act_c, act_m, act_f; long col_c, col_m, col_f; };
    static void split34(const M4 &T, double *R, double *t) { for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i * 3 + j] = T.a[i * 4 + j]; t[i] = T.a[i * 4 + 3]; } }
    /* `input` may be null when only the residual is wanted (no derivative tables, no active blocks) */
    void an_gather(const MatArrays &ma, const eVec *input, int frame_id, int cam_id, int marker_id, AnObs &q) const {
        const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        std::memset(&q, 0, sizeof q);
        size_t off = 0;
        { /* camera: Xc = inv(Tc) Xo, skipped for the root camera (mcm.cpp:617-621) */
            const MatArray &src = config.optimize_cam_poses ? ma.transforms_to_root_cam : mat_arrays.transforms_to_root_cam;
            const size_t idx = (size_t)src.m.at(cam_id), ridx = (size_t)src.m.at((int)root_cam);
            if ((size_t)cam_id == root_cam) { std::memcpy(q.Ri, I9, sizeof I9); }
            else { M4 inv = inv44(src.v[idx]); split34(inv, q.Ri, q.ti); }
            q.act_c = input && config.optimize_cam_poses && (size_t)cam_id != root_cam;
            q.col_c = (long)(off + 6 * (idx - (idx > ridx ? 1 : 0)));
            if (q.act_c) { double R[9]; split34(src.v[idx], R, q.tc); aar_an_rodrigues_derivs(&(*input)[(size_t)q.col_c], R, q.dRc); }
            if (config.optimize_cam_poses) off += (num_cameras - 1) * 6;
        }
        { /* marker, skipped for the root marker (mcm.cpp:624-628) */
            const MatArray &src = config.optimize_marker_poses ? ma.transforms_to_root_marker : mat_arrays.transforms_to_root_marker;
            const size_t idx = (size_t)src.m.at(marker_id), ridx = (size_t)src.m.at((int)root_marker);
            if ((size_t)marker_id == root_marker) std::memcpy(q.Rm, I9, sizeof I9); else split34(src.v[idx], q.Rm, q.tm);
            q.act_m = input && config.optimize_marker_poses && (size_t)marker_id != root_marker;
            q.col_m = (long)(off + 6 * (idx - (idx > ridx ? 1 : 0)));
            if (q.act_m) aar_an_rodrigues_derivs(&(*input)[(size_t)q.col_m], q.Rm, q.dRm);
            if (config.optimize_marker_poses) off += (num_markers - 1) * 6;
        }
        { /* frame */
            const MatArray &src = config.optimize_object_poses ? ma.object_to_global : mat_arrays.object_to_global;
            const size_t idx = (size_t)src.m.at(frame_id);
            split34(src.v[idx], q.Ro, q.to);
            q.act_f = input && config.optimize_object_poses;
            q.col_f = (long)(off + 6 * idx);
            if (q.act_f) aar_an_rodrigues_derivs(&(*input)[(size_t)q.col_f], q.Ro, q.dRo);
        }
        const M4 &K = mat_arrays.cam_mats.at_id(cam_id);
        q.fx = K.a[0]; q.cx = K.a[2]; q.fy = K.a[5]; q.cy = K.a[6];
        q.h = marker_points_3d_mat.a[1]; /* +halfSize, computed in float (marker.cpp:358-369) */
    }
    struct AnTripletSink {
        std::vector<Triplet> *out; size_t row0; long col_c, col_m, col_f;
        void put(int col, int corner, double jx, double jy) {
            const int blk = col / 6, d = col % 6;
            const long c = (blk == 0 ? col_c : blk == 1 ? col_m : col_f) + d;
            out->push_back({(int)(row0 + 2 * corner), (int)c, jx}); out->push_back({(int)(row0 + 2 * corner + 1), (int)c, jy});
        }
    };
    void an_residual8(const AnObs &q, const Marker &mk, double *e) const {
        float und[8]; for (int i = 0; i < 4; i++) { und[2 * i] = mk.x[i]; und[2 * i + 1] = mk.y[i]; }
        aar_an_null_sink ns;
        aar_an_observation(q.Ri, q.ti, q.dRc, q.tc, q.Ro, q.to, q.dRo, q.Rm, q.tm, q.dRm, q.fx, q.cx, q.fy, q.cy, q.h, und, false, false, false, e, ns);
    }
    /* the same projection through the restated OpenCV chain of project_marker (mul44 / inv44, K * T34 * X), kept in double:
     * written independently of include/aar_analytic.h; tests/test_analytic_cpu.py differentiates THIS numerically */
    void project_marker_fp64(const MatArrays &ma, size_t frame_id, size_t marker_id, size_t cam_id, double px[4], double py[4]) const {
        auto pick = [&](const MatArray &opt, const MatArray &fixed, bool optimized, size_t id) -> const M4 & { const MatArray &src = optimized ? opt : fixed; return src.v[(size_t)src.m.at((int)id)]; };
        M4 transform = pick(ma.object_to_global, mat_arrays.object_to_global, config.optimize_object_poses, frame_id);
        if (cam_id != root_cam) transform = mul44(inv44(pick(ma.transforms_to_root_cam, mat_arrays.transforms_to_root_cam, config.optimize_cam_poses, cam_id)), transform);
        if (marker_id != root_marker) transform = mul44(transform, pick(ma.transforms_to_root_marker, mat_arrays.transforms_to_root_marker, config.optimize_marker_poses, marker_id));
        const M4 &cam_mat = mat_arrays.cam_mats.at_id((int)cam_id);
        double KT[12], P[12];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) KT[i * 4 + j] = cam_mat.a[i * 4 + 0] * transform.a[0 * 4 + j] + cam_mat.a[i * 4 + 1] * transform.a[1 * 4 + j] + cam_mat.a[i * 4 + 2] * transform.a[2 * 4 + j];
        const double *X = marker_points_3d_mat.a;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) P[i * 4 + j] = KT[i * 4 + 0] * X[0 * 4 + j] + KT[i * 4 + 1] * X[1 * 4 + j] + KT[i * 4 + 2] * X[2 * 4 + j] + KT[i * 4 + 3] * X[3 * 4 + j];
        for (int c = 0; c < 4; c++) { px[c] = P[0 * 4 + c] / P[2 * 4 + c]; py[c] = P[1 * 4 + c] / P[2 * 4 + c]; }
    }
    void error_function_fp64_chain(const eVec &input, eVec &error) const {
        MatArrays ma; ma_init(ma); eVec2Mats(input, ma);
        error.assign(num_point_xys, 0.0);
        size_t index = 0;
        for (auto &f : frame_cam_markers) for (auto &c : f.second) for (auto &mk : c.second) {
            double px[4], py[4]; project_marker_fp64(ma, (size_t)f.first, (size_t)mk.id, (size_t)c.first, px, py);
            for (int i = 0; i < 4; i++) { error[index++] = (double)mk.x[i] - px[i]; error[index++] = (double)mk.y[i] - py[i]; }
        }
    }

    /* mcm.cpp:996-1028 */
    void eval_curr_solution(const MatArrays &ma, eVec &error) const {
        error.assign(num_point_xys, 0.0);
        size_t index = 0;
        for (auto frame_it = frame_cam_markers.begin(); frame_it != frame_cam_markers.end(); ++frame_it) {
            int frame_id = frame_it->first;
            for (auto it = frame_it->second.begin(); it != frame_it->second.end(); ++it) {
                size_t cam_id = it->first;
                const std::vector<Marker> &markers = it->second;
                for (size_t m = 0; m < markers.size(); m++) {
                    float px[4], py[4]; double e8[8];
                    if (analytic) { AnObs q; an_gather(ma, nullptr, frame_id, (int)cam_id, markers[m].id, q); an_residual8(q, markers[m], e8); }
                    else project_marker(ma, nullptr, frame_id, markers[m].id, cam_id, px, py);
                    for (int i = 0; i < 4; i++) {
                        double ex = markers[m].x[i] - px[i]; /* float - float, widened afterwards */
                        double ey = markers[m].y[i] - py[i];
                        if (analytic) { ex = e8[2 * i]; ey = e8[2 * i + 1]; }
                        if (with_huber) {
                            double e = ex * ex + ey * ey;
                            double w = getHubberMonoWeight(e, hubberDelta);
                            error[index++] = w * ex; error[index++] = w * ey;
                        } else { error[index++] = ex; error[index++] = ey; }
                    }
                }
            }
        }
    }
    /* mcm.cpp:731-737 */
    void error_function(const eVec &input, eVec &error) const {
        MatArrays ma; ma_init(ma); eVec2Mats(input, ma); eval_curr_solution(ma, error);
    }

    /* mcm.cpp:976-994 */
    void obtain_marker_derivs(const MatArrays &ma, const Override &oa, const Override &os, const Marker &mk, size_t frame_id, size_t marker_id, size_t cam_id,
                              size_t error_vec_ind, size_t param_index, std::vector<Triplet> &elems) const {
        float ax[4], ay[4], sx[4], sy[4];
        project_marker(ma, &oa, frame_id, marker_id, cam_id, ax, ay);
        project_marker(ma, &os, frame_id, marker_id, cam_id, sx, sy);
        for (int p = 0; p < 4; p++) {
            double error_x_a = mk.x[p] - ax[p];
            double error_x_s = mk.x[p] - sx[p];
            double error_x_deriv = (error_x_a - error_x_s) / (2 * J_delta);
            elems.push_back({(int)(error_vec_ind + p * 2), (int)param_index, error_x_deriv});
            double error_y_a = mk.y[p] - ay[p];
            double error_y_s = mk.y[p] - sy[p];
            double error_y_deriv = (error_y_a - error_y_s) / (2 * J_delta);
            elems.push_back({(int)(error_vec_ind + p * 2 + 1), (int)param_index, error_y_deriv});
        }
    }

    /* mcm.cpp:803-974 */
    void obtain_transformation_derivs(const MatArrays &ma, const eVec &input, param_type parameter_type, size_t transform_index, size_t param_offset, std::vector<Triplet> &elems) const {
        M4 T, T_a, T_s;
        size_t transform_param_index;
        if (parameter_type == camera) {
            T = ma.transforms_to_root_cam.v[transform_index];
            transform_param_index = param_offset + transform_index * 6;
            if (transform_index > (size_t)ma.transforms_to_root_cam.m.at((int)root_cam)) transform_param_index -= 6;
        } else if (parameter_type == marker) {
            T = ma.transforms_to_root_marker.v[transform_index];
            transform_param_index = param_offset + transform_index * 6;
            if (transform_index > (size_t)ma.transforms_to_root_marker.m.at((int)root_marker)) transform_param_index -= 6;
        } else if (parameter_type == object) {
            T = ma.object_to_global.v[transform_index];
            transform_param_index = param_offset + transform_index * 6;
        } else { /* intrinsics, :835-893 */
            M4 K = ma.cam_mats.v[transform_index], d = ma.dist_coeffs.v[transform_index];
            param_offset += transform_index * 9;
            for (int i = 0; i < 9; i++) {
                M4 K_a = K, K_s = K, d_a = d, d_s = d;
                switch (i) {
                    case 0: K_a.a[0] += J_delta; K_s.a[0] -= J_delta; break;
                    case 1: K_a.a[2] += J_delta; K_s.a[2] -= J_delta; break;
                    case 2: K_a.a[5] += J_delta; K_s.a[5] -= J_delta; break;
                    case 3: K_a.a[6] += J_delta; K_s.a[6] -= J_delta; break;
                    default: d_a.a[i - 4] += J_delta; d_s.a[i - 4] -= J_delta;
                }
                Override oa{intrinsics, transform_index, &K_a, &d_a}, os{intrinsics, transform_index, &K_s, &d_s};
                size_t cam_id = ma.cam_mats.id[transform_index];
                auto cmf = cam_marker_frame.find((int)cam_id);
                if (cmf == cam_marker_frame.end()) continue;
                for (auto &mit : cmf->second)
                    for (auto &fit : mit.second)
                        obtain_marker_derivs(ma, oa, os, fit.second.first, fit.first, mit.first, cam_id, fit.second.second, param_offset + i, elems);
            }
            return;
        }
        T_a = T; T_s = T;
        double rot_vec[3];
        for (int i = 0; i < 3; i++) rot_vec[i] = input[transform_param_index + i];
        for (int i = 0; i < 6; i++) {
            if (i < 3) {
                double rv_a[3] = {rot_vec[0], rot_vec[1], rot_vec[2]}, rv_s[3] = {rot_vec[0], rot_vec[1], rot_vec[2]};
                rv_a[i] += J_delta; rv_s[i] -= J_delta;
                double Ra[9], Rs[9];
                rodrigues_vec2mat(rv_a, Ra); rodrigues_vec2mat(rv_s, Rs);
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { T_a.a[r * 4 + c] = Ra[r * 3 + c]; T_s.a[r * 4 + c] = Rs[r * 3 + c]; }
            } else { T_a.a[(i - 3) * 4 + 3] += J_delta; T_s.a[(i - 3) * 4 + 3] -= J_delta; }
            Override oa{(int)parameter_type, transform_index, &T_a, nullptr}, os{(int)parameter_type, transform_index, &T_s, nullptr};
            if (parameter_type == camera) {
                size_t cam_id = ma.transforms_to_root_cam.id[transform_index];
                auto cmf = cam_marker_frame.find((int)cam_id);
                if (cmf != cam_marker_frame.end())
                    for (auto &mit : cmf->second)
                        for (auto &fit : mit.second)
                            obtain_marker_derivs(ma, oa, os, fit.second.first, fit.first, mit.first, cam_id, fit.second.second, transform_param_index + i, elems);
            } else if (parameter_type == marker) {
                size_t marker_id = ma.transforms_to_root_marker.id[transform_index];
                auto mcf = marker_cam_frame.find((int)marker_id);
                if (mcf != marker_cam_frame.end())
                    for (auto &cit : mcf->second)
                        for (auto &fit : cit.second)
                            obtain_marker_derivs(ma, oa, os, fit.second.first, fit.first, marker_id, cit.first, fit.second.second, transform_param_index + i, elems);
            } else {
                size_t frame_id = ma.object_to_global.id[transform_index];
                auto fcm = frame_cam_marker.find((int)frame_id);
                if (fcm != frame_cam_marker.end())
                    for (auto &cit : fcm->second)
                        for (auto &mit : cit.second)
                            obtain_marker_derivs(ma, oa, os, mit.second.first, frame_id, mit.first, cit.first, mit.second.second, transform_param_index + i, elems);
            }
            /* restore, :965-972 */
            if (i < 3) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { T_a.a[r * 4 + c] = T.a[r * 4 + c]; T_s.a[r * 4 + c] = T.a[r * 4 + c]; } }
            else { for (int r = 0; r < 3; r++) { T_a.a[r * 4 + 3] = T.a[r * 4 + 3]; T_s.a[r * 4 + 3] = T.a[r * 4 + 3]; } }
        }
    }

    /* mcm.cpp:739-801 — triplets in the reference's emission order (block index order) */
    void jacobian_triplets(const eVec &input, std::vector<Triplet> &all) const {
        MatArrays ma; ma_init(ma); eVec2Mats(input, ma);
        all.clear();
        if (analytic) { /* one 8 x 18 block per observation that owns Jacobian rows (the last of a repeated (frame, cam, marker), mcm.cpp:368-370) */
            size_t row0 = 0;
            for (auto &f : frame_cam_markers) for (auto &c : f.second) for (auto &mk : c.second) {
                if (frame_cam_marker.at(f.first).at(c.first).at(mk.id).second == row0) {
                    AnObs q; an_gather(ma, &input, f.first, c.first, mk.id, q);
                    float und[8]; for (int i = 0; i < 4; i++) { und[2 * i] = mk.x[i]; und[2 * i + 1] = mk.y[i]; }
                    double e8[8]; AnTripletSink sink{&all, row0, q.col_c, q.col_m, q.col_f};
                    aar_an_observation(q.Ri, q.ti, q.dRc, q.tc, q.Ro, q.to, q.dRo, q.Rm, q.tm, q.dRm, q.fx, q.cx, q.fy, q.cy, q.h, und, q.act_c, q.act_m, q.act_f, e8, sink);
                }
                row0 += 8;
            }
            return;
        }
        size_t param_offset = 0;
        auto run = [&](param_type t, long long n, std::function<bool(long long)> skip) {
            std::vector<std::vector<Triplet>> elems((size_t)n);
#pragma omp parallel for schedule(dynamic, 8)
            for (long long i = 0; i < n; i++) { if (skip(i)) continue; obtain_transformation_derivs(ma, input, t, (size_t)i, param_offset, elems[(size_t)i]); }
            size_t tot = all.size(); for (auto &e : elems) tot += e.size(); all.reserve(tot);
            for (auto &e : elems) all.insert(all.end(), e.begin(), e.end());
        };
        if (config.optimize_cam_poses) { run(camera, (long long)num_cameras, [&](long long i) { return (size_t)ma.transforms_to_root_cam.id[(size_t)i] == root_cam; }); param_offset += (num_cameras - 1) * 6; }
        if (config.optimize_marker_poses) { run(marker, (long long)num_markers, [&](long long i) { return (size_t)ma.transforms_to_root_marker.id[(size_t)i] == root_marker; }); param_offset += (num_markers - 1) * 6; }
        if (config.optimize_object_poses) { run(object, (long long)num_frames, [](long long) { return false; }); param_offset += num_frames * 6; }
        if (config.optimize_cam_intrinsics) { run(intrinsics, (long long)num_cameras, [](long long) { return false; }); param_offset += num_cameras * 9; }
    }

    /* mcm.cpp:678-729 */
    void error_function_tracking(const eVec &input, eVec &error) const {
        M4 object_to_global = eye4();
        double r[3], R[9];
        for (int i = 0; i < 3; i++) { r[i] = input[i]; object_to_global.a[i * 4 + 3] = input[i + 3]; }
        rodrigues_vec2mat(r, R);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) object_to_global.a[i * 4 + j] = R[i * 3 + j];
        error.assign(num_point_xys, 0.0);
        size_t index = 0;
        if (frame_cam_markers.empty()) return;
        const auto &cams_markers = frame_cam_markers.begin()->second;
        for (auto &cm : cams_markers) {
            size_t cam_id = cm.first;
            const std::vector<Marker> &markers = cm.second;
            for (size_t m = 0; m < markers.size(); m++) {
                M4 tp = mul44(mul44(inv44(mat_arrays.transforms_to_root_cam.at_id((int)cam_id)), object_to_global), marker_object_points_3d_mats.at(markers[m].id));
                const M4 &K = mat_arrays.cam_mats.at_id((int)cam_id);
                double q[12];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 4; j++)
                        q[i * 4 + j] = K.a[i * 4 + 0] * tp.a[0 * 4 + j] + K.a[i * 4 + 1] * tp.a[1 * 4 + j] + K.a[i * 4 + 2] * tp.a[2 * 4 + j];
                for (int i = 0; i < 4; i++) {
                    double pxs = q[0 * 4 + i] / q[2 * 4 + i], pys = q[1 * 4 + i] / q[2 * 4 + i];
                    double ex = markers[m].x[i] - pxs; /* float corner minus DOUBLE projection */
                    double ey = markers[m].y[i] - pys;
                    if (with_huber) {
                        double e = ex * ex + ey * ey;
                        double w = getHubberMonoWeight(e, hubberDelta);
                        error[index++] = w * ex; error[index++] = w * ey;
                    } else { error[index++] = ex; error[index++] = ey; }
                }
            }
        }
    }
    /* the part of mcm.cpp:430-443 before the solve */
    void track_prepare() {
        mats2eVec(config);
        marker_object_points_3d_mats.clear();
        for (auto &p : mat_arrays.transforms_to_root_marker.m)
            marker_object_points_3d_mats[p.first] = mul44(mat_arrays.transforms_to_root_marker.v[(size_t)p.second], marker_points_3d_mat);
        hubberDelta = 10;
    }
    /* mcm.cpp:412-417 */
    void optCallBack() { if (hubberDelta > 2.5) hubberDelta -= 7.5 / 500; }
};

/* --------------------------------------------------------- triplets -> CSC (setFromTriplets) */
void triplets_to_csc(const std::vector<Triplet> &t, int rows, int cols, std::vector<int64_t> &colptr, std::vector<int> &rowidx, std::vector<double> &vals) {
    (void)rows;
    std::vector<size_t> order(t.size());
    for (size_t i = 0; i < t.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return t[a].col != t[b].col ? t[a].col < t[b].col : t[a].row < t[b].row; });
    colptr.assign((size_t)cols + 1, 0); rowidx.clear(); vals.clear();
    int lastc = -1, lastr = -1;
    for (size_t k : order) {
        if (t[k].col == lastc && t[k].row == lastr) { vals.back() += t[k].val; continue; } /* duplicates are summed */
        rowidx.push_back(t[k].row); vals.push_back(t[k].val); colptr[(size_t)t[k].col + 1]++;
        lastc = t[k].col; lastr = t[k].row;
    }
    for (int c = 0; c < cols; c++) colptr[(size_t)c + 1] += colptr[(size_t)c];
}

/* ------------------------------------------ exact block-arrow normal-equation tools (port) --
 * JtJ/B from triplets; frames eliminated by 6x6 Cholesky, reduced system by dense Cholesky.
 * Mathematically the same linear solve SimplicialLDLT performs (sparselevmarq.h:394-400). */
struct Normal { int n, n_r; std::vector<double> Hrr, Hff, W, B; /* Hrr n_r*n_r ; Hff F*36 ; W F*n_r*6 ; B n */ };

bool chol_inplace(std::vector<double> &A, int n) { /* lower */
    for (int j = 0; j < n; j++) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0)) return false;
        d = std::sqrt(d); A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    return true;
}
void chol_solve(const std::vector<double> &L, int n, double *x) {
    for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * x[k]; x[i] = s / L[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k]; x[i] = s / L[(size_t)i * n + i]; }
}

} // namespace

/* =================================================================== C API (ctypes) ======== */
struct OracleHandle {
    Mcm mcm;
    std::vector<int64_t> colptr; std::vector<int> rowidx; std::vector<double> vals; /* last Jacobian */
    std::vector<double> trace; /* per iteration: cost, mu, huber, accepted */
};

static M4 m4_from(const double *p) { M4 m; std::memcpy(m.a, p, 16 * sizeof(double)); return m; }

extern "C" {

void aar_oracle_set_sincos_mode(int mode) { g_sincos_mode = mode; }

/* cams/markers/frames: ids + row-major 4x4 doubles.  K: 9 doubles per cam, dist: 5 per cam (same
 * order as cam_ids).  Detections in file order: frame id, cam id, marker id, 8 floats x0 y0 .. */
OracleHandle *aar_oracle_create(int C, const int *cam_ids, const double *cam_T, const double *K, const double *dist,
                                int M, const int *marker_ids, const double *marker_T,
                                int F, const int *frame_ids, const double *frame_T,
                                int root_cam, int root_marker, float marker_size,
                                int64_t ndet, const int *det_frame, const int *det_cam, const int *det_marker, const float *det_xy) {
    OracleHandle *h = new OracleHandle();
    std::map<int, M4> tc, tm, tf, Ks, ds;
    for (int i = 0; i < C; i++) {
        tc[cam_ids[i]] = m4_from(cam_T + 16 * (size_t)i);
        M4 k; std::memset(k.a, 0, sizeof k.a);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) k.a[r * 4 + c] = K[9 * (size_t)i + r * 3 + c];
        Ks[cam_ids[i]] = k;
        M4 d; std::memset(d.a, 0, sizeof d.a);
        for (int j = 0; j < 5; j++) d.a[j] = dist[5 * (size_t)i + j];
        ds[cam_ids[i]] = d;
    }
    for (int i = 0; i < M; i++) tm[marker_ids[i]] = m4_from(marker_T + 16 * (size_t)i);
    for (int i = 0; i < F; i++) tf[frame_ids[i]] = m4_from(frame_T + 16 * (size_t)i);
    FCM fcm;
    for (int64_t i = 0; i < ndet; i++) {
        Marker mk; mk.id = det_marker[i];
        for (int c = 0; c < 4; c++) { mk.x[c] = det_xy[8 * i + 2 * c]; mk.y[c] = det_xy[8 * i + 2 * c + 1]; }
        fcm[det_frame[i]][det_cam[i]].push_back(mk);
    }
    h->mcm.init((size_t)root_cam, tc, (size_t)root_marker, tm, tf, fcm, marker_size, Ks, ds);
    return h;
}
void aar_oracle_destroy(OracleHandle *h) { delete h; }

void aar_oracle_set_config(OracleHandle *h, int cams, int markers, int objects, int intrinsics, int with_huber, float huber_delta) {
    Config c; c.optimize_cam_poses = cams; c.optimize_marker_poses = markers; c.optimize_object_poses = objects; c.optimize_cam_intrinsics = intrinsics;
    h->mcm.set_config(c); h->mcm.with_huber = with_huber != 0; h->mcm.hubberDelta = huber_delta;
}
/* SparseLevMarq::Params::maxIters as MultiCamMapper sets it (mcm.cpp:326-330); tests lower it to compare single LM steps */
void aar_oracle_set_max_iters(OracleHandle *h, int max_iters) { h->mcm.maxIters = max_iters; }
/* analytic-Jacobian / full-FP64 variant (include/aar_analytic.h); requires intrinsics off */
void aar_oracle_set_analytic(OracleHandle *h, int on) { h->mcm.analytic = on != 0; }
/* m - p of the restated OpenCV chain kept in double (no float32 rounding, no Huber): the function the analytic Jacobian is
 * differentiated against numerically in tests/test_analytic_cpu.py */
void aar_oracle_error_fp64_chain(OracleHandle *h, const double *z, double *r) {
    eVec e(z, z + h->mcm.num_vars), err; h->mcm.error_function_fp64_chain(e, err);
    std::memcpy(r, err.data(), err.size() * sizeof(double));
}
void aar_oracle_rodrigues_derivs(const double *r, double *dR) { double R[9]; rodrigues_vec2mat(r, R); aar_an_rodrigues_derivs(r, R, dR); }
int64_t aar_oracle_num_vars(OracleHandle *h) { return (int64_t)h->mcm.num_vars; }
int64_t aar_oracle_num_rows(OracleHandle *h) { return (int64_t)h->mcm.num_point_xys; }

/* row map: observation o (row0 = 8*o) -> ids, raw corners (as captured by fill_iteration_arrays
 * BEFORE the undistortion) and undistorted corners.  has_jac = 0 when a later duplicate of the
 * same (frame,cam,marker) replaced this one in the inverted indices (mcm.cpp:368-370). */
void aar_oracle_get_observations(OracleHandle *h, int *frame_id, int *cam_id, int *marker_id, float *und_xy, float *raw_xy, int *has_jac) {
    const Mcm &m = h->mcm;
    size_t o = 0;
    for (auto &f : m.frame_cam_markers)
        for (auto &c : f.second)
            for (auto &mk : c.second) {
                frame_id[o] = f.first; cam_id[o] = c.first; marker_id[o] = mk.id;
                for (int k = 0; k < 4; k++) { und_xy[8 * o + 2 * k] = mk.x[k]; und_xy[8 * o + 2 * k + 1] = mk.y[k]; }
                const auto &rec = m.frame_cam_marker.at(f.first).at(c.first).at(mk.id);
                has_jac[o] = rec.second == 8 * o;
                for (int k = 0; k < 4; k++) { raw_xy[8 * o + 2 * k] = rec.first.x[k]; raw_xy[8 * o + 2 * k + 1] = rec.first.y[k]; }
                o++;
            }
}

void aar_oracle_mats2evec(OracleHandle *h, double *z) { h->mcm.mats2eVec(h->mcm.config); std::memcpy(z, h->mcm.io_vec.data(), h->mcm.io_vec.size() * sizeof(double)); }
/* z -> 4x4 matrices of every camera / marker / frame in index order (eVec2Mats into mat_arrays, mcm.cpp:427) */
void aar_oracle_evec2mats(OracleHandle *h, const double *z, double *cam_T, double *marker_T, double *frame_T) {
    Mcm &m = h->mcm;
    eVec e(z, z + m.num_vars);
    m.eVec2Mats(e, m.mat_arrays);
    for (size_t i = 0; i < m.num_cameras; i++) std::memcpy(cam_T + 16 * i, m.mat_arrays.transforms_to_root_cam.v[i].a, 128);
    for (size_t i = 0; i < m.num_markers; i++) std::memcpy(marker_T + 16 * i, m.mat_arrays.transforms_to_root_marker.v[i].a, 128);
    for (size_t i = 0; i < m.num_frames; i++) std::memcpy(frame_T + 16 * i, m.mat_arrays.object_to_global.v[i].a, 128);
}

void aar_oracle_error(OracleHandle *h, const double *z, double *r) {
    eVec e(z, z + h->mcm.num_vars), err; h->mcm.error_function(e, err);
    std::memcpy(r, err.data(), err.size() * sizeof(double));
}
/* Jacobian in compressed-column form exactly as Eigen's setFromTriplets leaves it (mcm.cpp:800);
 * returns nnz; fetch with aar_oracle_get_jacobian. */
int64_t aar_oracle_jacobian(OracleHandle *h, const double *z) {
    eVec e(z, z + h->mcm.num_vars);
    std::vector<Triplet> t; h->mcm.jacobian_triplets(e, t);
    triplets_to_csc(t, (int)h->mcm.num_point_xys, (int)h->mcm.num_vars, h->colptr, h->rowidx, h->vals);
    return (int64_t)h->vals.size();
}
void aar_oracle_get_jacobian(OracleHandle *h, int64_t *colptr, int *rowidx, double *vals) {
    std::memcpy(colptr, h->colptr.data(), h->colptr.size() * sizeof(int64_t));
    std::memcpy(rowidx, h->rowidx.data(), h->rowidx.size() * sizeof(int));
    std::memcpy(vals, h->vals.data(), h->vals.size() * sizeof(double));
}

/* pieces exposed for the cv2 pin tests */
void aar_oracle_rodrigues(const double *r, double *R) { rodrigues_vec2mat(r, R); }
void aar_oracle_rodrigues_inv(const double *R, double *r) { rodrigues_mat2vec(R, r); }
void aar_oracle_inv44(const double *A, double *B) { M4 r = inv44(m4_from(A)); std::memcpy(B, r.a, 128); }
void aar_oracle_mul44(const double *A, const double *B, double *Cc) { M4 r = mul44(m4_from(A), m4_from(B)); std::memcpy(Cc, r.a, 128); }
void aar_oracle_undistort(int64_t n, const float *xy, const double *K, const double *dist, float *out) {
    for (int64_t i = 0; i < n; i++) undistort_point(xy[2 * i], xy[2 * i + 1], K, dist, &out[2 * i], &out[2 * i + 1]);
}
void aar_oracle_sincos(double x, double *s, double *c) { aar_sincos(x, s, c); }

/* normal equations at z with the CURRENT residual convention of sparselevmarq.h:353-367:
 * A = JtJ, B = -Jt r.  Returns the frame-eliminated reduced system for frames [f_lo, f_hi):
 *   S = sum_f ( Hrr_f - W_f (Hff_f + mu I)^-1 W_f^T )   (mu I on the reduced diagonal NOT added)
 *   b = sum_f ( Br_f  - W_f (Hff_f + mu I)^-1 Bf )      , cost = sum r^2 over those frames.
 * Requires config = cams+markers+objects optimised, intrinsics off. */
int aar_oracle_reduced_system(OracleHandle *h, const double *z, double mu, int f_lo, int f_hi, double *S, double *b, double *cost) {
    Mcm &m = h->mcm;
    if (!m.config.optimize_object_poses || m.config.optimize_cam_intrinsics) return -1;
    const int n_r = (int)((m.config.optimize_cam_poses ? (m.num_cameras - 1) * 6 : 0) + (m.config.optimize_marker_poses ? (m.num_markers - 1) * 6 : 0));
    eVec e(z, z + m.num_vars), r; m.error_function(e, r);
    std::vector<Triplet> t; m.jacobian_triplets(e, t);
    /* rows grouped by frame: row -> frame index */
    std::vector<int> row_frame(m.num_point_xys / 8);
    { size_t o = 0; int fi = 0; for (auto &f : m.frame_cam_markers) { fi = m.mat_arrays.object_to_global.m.at(f.first); for (auto &c : f.second) for (size_t k = 0; k < c.second.size(); k++) row_frame[o++] = fi; } }
    std::vector<int64_t> colptr; std::vector<int> rowidx; std::vector<double> vals;
    triplets_to_csc(t, (int)m.num_point_xys, (int)m.num_vars, colptr, rowidx, vals);
    /* per-row sparse lists */
    std::vector<std::vector<std::pair<int, double>>> rows(m.num_point_xys);
    for (int c = 0; c < (int)m.num_vars; c++) for (int64_t k = colptr[c]; k < colptr[c + 1]; k++) rows[(size_t)rowidx[k]].push_back({c, vals[k]});
    std::fill(S, S + (size_t)n_r * n_r, 0.0); std::fill(b, b + n_r, 0.0); *cost = 0;
    const int F = (int)m.num_frames;
    std::vector<double> Hff((size_t)36), W((size_t)n_r * 6), Bf(6);
    /* bucket rows per frame */
    std::vector<std::vector<int>> frame_rows((size_t)F);
    for (size_t row = 0; row < m.num_point_xys; row++) frame_rows[(size_t)row_frame[row / 8]].push_back((int)row);
    for (int f = f_lo; f < f_hi && f < F; f++) {
        std::fill(Hff.begin(), Hff.end(), 0.0); std::fill(W.begin(), W.end(), 0.0); std::fill(Bf.begin(), Bf.end(), 0.0);
        const int fcol = n_r + 6 * f;
        for (int row : frame_rows[(size_t)f]) {
            const auto &R = rows[(size_t)row];
            double rv = r[(size_t)row]; *cost += rv * rv;
            for (auto &a : R) {
                if (a.first < n_r) b[a.first] -= a.second * rv; else Bf[(size_t)(a.first - fcol)] -= a.second * rv;
                for (auto &c : R) {
                    if (a.first < n_r && c.first < n_r) S[(size_t)a.first * n_r + c.first] += a.second * c.second;
                    else if (a.first < n_r && c.first >= n_r) W[(size_t)a.first * 6 + (c.first - fcol)] += a.second * c.second;
                    else if (a.first >= n_r && c.first >= n_r) Hff[(size_t)(a.first - fcol) * 6 + (c.first - fcol)] += a.second * c.second;
                }
            }
        }
        for (int i = 0; i < 6; i++) Hff[(size_t)i * 6 + i] += mu;
        std::vector<double> L = Hff;
        if (!chol_inplace(L, 6)) return -2;
        /* Y = Hff^-1 [W^T | Bf] */
        std::vector<double> col(6);
        std::vector<double> Yw((size_t)n_r * 6);
        for (int i = 0; i < n_r; i++) { for (int k = 0; k < 6; k++) col[(size_t)k] = W[(size_t)i * 6 + k]; chol_solve(L, 6, col.data()); for (int k = 0; k < 6; k++) Yw[(size_t)i * 6 + k] = col[(size_t)k]; }
        std::vector<double> yb = Bf; chol_solve(L, 6, yb.data());
        for (int i = 0; i < n_r; i++) {
            bool nz = false; for (int k = 0; k < 6; k++) nz |= W[(size_t)i * 6 + k] != 0;
            if (!nz) continue;
            for (int j = 0; j < n_r; j++) { double s = 0; for (int k = 0; k < 6; k++) s += W[(size_t)i * 6 + k] * Yw[(size_t)j * 6 + k]; S[(size_t)i * n_r + j] -= s; }
            double s = 0; for (int k = 0; k < 6; k++) s += W[(size_t)i * 6 + k] * yb[(size_t)k]; b[i] -= s;
        }
    }
    return 0;
}

/* ----------------------------------------------------------------- LM: restated port ------ *
 * sparselevmarq.h:237-249 (init), 348-430 (step), 439-472 (solve).  Linear solve: frames
 * eliminated block-wise, dense Cholesky on the rest — exact, like SimplicialLDLT.
 * `v` is uninitialised in the reference until the first accepted step (h:133, 411, 416); 2 here.
 * track = 1 uses the 2-argument solve: FD Jacobian calcDerivates (h:198-220), der_epsilon,
 * entries |d| <= 1e-4 dropped, on error_function_tracking.
 * trace: per iteration [cost, mu, gain, tries, accepted, huberDelta].  Returns iterations. */
static void fd_tracking_jacobian(const Mcm &m, const eVec &z, std::vector<Triplet> &t) {
    t.clear();
    for (int i = 0; i < (int)z.size(); i++) {
        eVec zp(z), zm(z), xp, xm;
        zp[(size_t)i] += m.der_epsilon; zm[(size_t)i] -= m.der_epsilon;
        m.error_function_tracking(zp, xp); m.error_function_tracking(zm, xm);
        for (int r = 0; r < (int)xp.size(); r++) {
            double d = (xp[(size_t)r] - xm[(size_t)r]) / (2.f * m.der_epsilon);
            if (std::fabs(d) > 1e-4) t.push_back({r, i, d});
        }
    }
}

int aar_oracle_lm_port(OracleHandle *h, double *z_io, int track, int max_trace, double *trace, double *final_cost) {
    Mcm &m = h->mcm;
    const int n = track ? 6 : (int)m.num_vars;
    auto f = [&](const eVec &z, eVec &x) { if (track) m.error_function_tracking(z, x); else m.error_function(z, x); };
    eVec curr_z(z_io, z_io + n), x64;
    f(curr_z, x64);
    double currErr = 0; for (double v : x64) currErr += v * v;
    double prevErr = currErr, mu = -1, v = 2;
    const int rows = (int)x64.size();
    int it = 0, mustExit = 0;
    /* layout of the arrow system */
    const int n_r = track ? 0 : (int)((m.config.optimize_cam_poses ? (m.num_cameras - 1) * 6 : 0) + (m.config.optimize_marker_poses ? (m.num_markers - 1) * 6 : 0));
    const bool arrow = !track && m.config.optimize_object_poses && !m.config.optimize_cam_intrinsics;
    const int F = arrow ? (int)m.num_frames : 0;
    const int n_dense = arrow ? n_r : n;
    for (it = 0; it < m.maxIters && !mustExit; it++) {
        std::vector<Triplet> t;
        if (track) fd_tracking_jacobian(m, curr_z, t); else m.jacobian_triplets(curr_z, t);
        std::vector<int64_t> colptr; std::vector<int> rowidx; std::vector<double> vals;
        triplets_to_csc(t, rows, n, colptr, rowidx, vals);
        std::vector<std::vector<std::pair<int, double>>> rws((size_t)rows);
        for (int c = 0; c < n; c++) for (int64_t k = colptr[(size_t)c]; k < colptr[(size_t)c + 1]; k++) rws[(size_t)rowidx[(size_t)k]].push_back({c, vals[(size_t)k]});
        /* A = JtJ in arrow blocks, B = -Jt x */
        std::vector<double> Hd((size_t)n_dense * n_dense, 0.0), Hff((size_t)F * 36, 0.0), W((size_t)F * n_r * 6, 0.0), B((size_t)n, 0.0);
        for (int r = 0; r < rows; r++) {
            const auto &R = rws[(size_t)r];
            for (auto &a : R) {
                B[(size_t)a.first] -= a.second * x64[(size_t)r];
                for (auto &c : R) {
                    if (a.first < n_dense && c.first < n_dense) Hd[(size_t)a.first * n_dense + c.first] += a.second * c.second;
                    else if (a.first < n_dense && c.first >= n_dense) { int fi = (c.first - n_dense) / 6; W[((size_t)fi * n_r + a.first) * 6 + (c.first - n_dense) % 6] += a.second * c.second; }
                    else if (a.first >= n_dense && c.first >= n_dense) { int fi = (a.first - n_dense) / 6; Hff[(size_t)fi * 36 + ((a.first - n_dense) % 6) * 6 + (c.first - n_dense) % 6] += a.second * c.second; }
                }
            }
        }
        if (mu < 0) {
            double maxv = std::numeric_limits<double>::lowest();
            for (int i = 0; i < n_dense; i++) maxv = std::max(maxv, Hd[(size_t)i * n_dense + i]);
            for (int fi = 0; fi < F; fi++) for (int i = 0; i < 6; i++) maxv = std::max(maxv, Hff[(size_t)fi * 36 + i * 7]);
            mu = maxv * m.tau;
        }
        double gain = 0; int ntries = 0; bool isStepAccepted = false;
        do {
            /* solve (A + mu I) delta = B */
            std::vector<double> S = Hd, bb(B.begin(), B.begin() + n_dense), delta((size_t)n, 0.0);
            for (int i = 0; i < n_dense; i++) S[(size_t)i * n_dense + i] += mu;
            std::vector<std::vector<double>> Lf((size_t)F);
            for (int fi = 0; fi < F; fi++) {
                std::vector<double> L(Hff.begin() + (size_t)fi * 36, Hff.begin() + (size_t)fi * 36 + 36);
                for (int i = 0; i < 6; i++) L[(size_t)i * 7] += mu;
                chol_inplace(L, 6); Lf[(size_t)fi] = L;
                std::vector<double> yb(B.begin() + n_dense + 6 * fi, B.begin() + n_dense + 6 * fi + 6); chol_solve(L, 6, yb.data());
                const double *Wf = &W[(size_t)fi * n_r * 6];
                std::vector<int> nzrows; for (int i = 0; i < n_r; i++) { bool nz = false; for (int k = 0; k < 6; k++) nz |= Wf[(size_t)i * 6 + k] != 0; if (nz) nzrows.push_back(i); }
                std::vector<double> Yw(nzrows.size() * 6);
                for (size_t a = 0; a < nzrows.size(); a++) { double col[6]; for (int k = 0; k < 6; k++) col[k] = Wf[(size_t)nzrows[a] * 6 + k]; chol_solve(L, 6, col); for (int k = 0; k < 6; k++) Yw[a * 6 + k] = col[k]; }
                for (size_t a = 0; a < nzrows.size(); a++) {
                    int i = nzrows[a];
                    for (size_t c = 0; c < nzrows.size(); c++) { double s = 0; for (int k = 0; k < 6; k++) s += Wf[(size_t)i * 6 + k] * Yw[c * 6 + k]; S[(size_t)i * n_dense + nzrows[c]] -= s; }
                    double s = 0; for (int k = 0; k < 6; k++) s += Wf[(size_t)i * 6 + k] * yb[(size_t)k]; bb[(size_t)i] -= s;
                }
            }
            if (n_dense > 0) { chol_inplace(S, n_dense); chol_solve(S, n_dense, bb.data()); }
            for (int i = 0; i < n_dense; i++) delta[(size_t)i] = bb[(size_t)i];
            for (int fi = 0; fi < F; fi++) {
                double rhs[6];
                const double *Wf = &W[(size_t)fi * n_r * 6];
                for (int k = 0; k < 6; k++) { double s = B[(size_t)n_dense + 6 * fi + k]; for (int i = 0; i < n_r; i++) s -= Wf[(size_t)i * 6 + k] * delta[(size_t)i]; rhs[k] = s; }
                chol_solve(Lf[(size_t)fi], 6, rhs);
                for (int k = 0; k < 6; k++) delta[(size_t)n_dense + 6 * fi + k] = rhs[k];
            }
            eVec estimated_z(curr_z);
            for (int i = 0; i < n; i++) estimated_z[(size_t)i] += delta[(size_t)i];
            f(estimated_z, x64);
            double err = 0; for (double q : x64) err += q * q;
            double L = 0; for (int i = 0; i < n; i++) L += delta[(size_t)i] * (mu * delta[(size_t)i] - B[(size_t)i]);
            L *= 0.5;
            gain = (err - prevErr) / L;
            if (gain > 0 && ((err - prevErr) < 0)) {
                mu = mu * std::max(double(0.33), double(1. - std::pow(2 * gain - 1, 3)));
                v = 2.f; currErr = err; curr_z = estimated_z; isStepAccepted = true;
            } else { mu = mu * v; v = v * 5; }
        } while (gain <= 0 && ntries++ < 5 && !isStepAccepted);
        if (currErr < m.minError) mustExit = 1;
        if (std::fabs(prevErr - currErr) <= m.min_step_error_diff || std::fabs((prevErr - currErr) / rows) <= m.min_average_step_error_diff || !isStepAccepted) mustExit = 2;
        if (currErr > prevErr) mustExit = 3;
        if (trace && it < max_trace) { double *tr = trace + 6 * (size_t)it; tr[0] = currErr; tr[1] = mu; tr[2] = gain; tr[3] = ntries + (isStepAccepted ? 1 : 0); tr[4] = isStepAccepted; tr[5] = m.hubberDelta; }
        if (!track) m.optCallBack();   /* the callback is installed by solve() only (mcm.cpp:422); apps/track.cpp goes init -> track() */
        prevErr = currErr;
    }
    std::memcpy(z_io, curr_z.data(), (size_t)n * sizeof(double));
    *final_cost = currErr;
    return it;
}

/* MultiCamMapper::solve() (mcm.cpp:419-428) / track() (:430-443) around the port LM. */
int aar_oracle_solve_port(OracleHandle *h, double *z_io, int max_trace, double *trace, double *final_cost) {
    h->mcm.hubberDelta = 10;
    return aar_oracle_lm_port(h, z_io, 0, max_trace, trace, final_cost);
}
/* track: object pose + detections of ONE frame against the fixed rig; config must be
 * cams=0 markers=0 objects=1 intrinsics=0 (apps/track.cpp:95-97). */
int aar_oracle_track_init(OracleHandle *h, int frame_id, const double *T, int64_t ndet, const int *det_cam, const int *det_marker, const float *det_xy) {
    std::map<int, M4> poses; poses[frame_id] = m4_from(T);
    FCM fcm;
    for (int64_t i = 0; i < ndet; i++) {
        Marker mk; mk.id = det_marker[i];
        for (int c = 0; c < 4; c++) { mk.x[c] = det_xy[8 * i + 2 * c]; mk.y[c] = det_xy[8 * i + 2 * c + 1]; }
        fcm[frame_id][det_cam[i]].push_back(mk);
    }
    h->mcm.init_track(poses, fcm);
    h->mcm.track_prepare();
    return (int)h->mcm.num_point_xys;
}
void aar_oracle_track_get_z(OracleHandle *h, double *z) { std::memcpy(z, h->mcm.io_vec.data(), 6 * sizeof(double)); }
void aar_oracle_track_error(OracleHandle *h, const double *z, double *r) { eVec e(z, z + 6), err; h->mcm.error_function_tracking(e, err); std::memcpy(r, err.data(), err.size() * sizeof(double)); }
int aar_oracle_track_port(OracleHandle *h, double *z_io, int max_trace, double *trace, double *final_cost) {
    return aar_oracle_lm_port(h, z_io, 1, max_trace, trace, final_cost);
}

int aar_oracle_has_reference_slm(void) {
#ifdef AAR_ORACLE_WITH_REFERENCE_SLM
    return 1;
#else
    return 0;
#endif
}

#ifdef AAR_ORACLE_WITH_REFERENCE_SLM
/* MultiCamMapper::solve() / track() driving the UNMODIFIED reference sparselevmarq.h. */
static int run_reference_slm(OracleHandle *h, double *z_io, int track, int max_trace, double *trace, double *final_cost, int verbose) {
    Mcm &m = h->mcm;
    typedef ucoslam::SparseLevMarq<double> SLM;
    SLM solver;
    SLM::Params p; p.verbose = verbose != 0; p.maxIters = m.maxIters; p.min_average_step_error_diff = m.min_average_step_error_diff; /* mcm.cpp:326-330 */
    solver.setParams(p);
    solver.v = 2; /* uninitialised in the reference (h:133); pinned to the documented choice */
    const int n = track ? 6 : (int)m.num_vars;
    SLM::eVector z(n);
    for (int i = 0; i < n; i++) z(i) = z_io[i];
    int it = 0;
    double last_cost = -1;
    auto f = [&](const SLM::eVector &in, SLM::eVector &out) {
        eVec e(in.data(), in.data() + in.size()), err;
        if (track) m.error_function_tracking(e, err); else m.error_function(e, err);
        out.resize((long)err.size());
        std::memcpy(out.data(), err.data(), err.size() * sizeof(double));
    };
    auto J = [&](const SLM::eVector &in, Eigen::SparseMatrix<double> &Jm) {
        eVec e(in.data(), in.data() + in.size());
        std::vector<Triplet> t; m.jacobian_triplets(e, t);
        Jm.resize((long)m.num_point_xys, (long)m.num_vars);
        std::vector<Eigen::Triplet<double>> et; et.reserve(t.size());
        for (auto &q : t) et.push_back(Eigen::Triplet<double>(q.row, q.col, q.val));
        Jm.setFromTriplets(et.begin(), et.end());
    };
    solver.setStepCallBackFunc([&](const SLM::eVector &) {
        if (trace && it < max_trace) { double *tr = trace + 6 * (size_t)it; tr[0] = solver.currErr; tr[1] = solver.mu; tr[2] = 0; tr[3] = 0; tr[4] = (last_cost < 0 || solver.currErr < last_cost); tr[5] = m.hubberDelta; }
        last_cost = solver.currErr;
        it++;
        if (!track) m.optCallBack();   /* MultiCamMapper::track() never installs optCallBack (mcm.cpp:422 is in solve() only): hubberDelta stays 10 */
    });
    m.hubberDelta = 10;
    double fc = track ? solver.solve(z, f) : solver.solve(z, f, J);
    for (int i = 0; i < n; i++) z_io[i] = z(i);
    *final_cost = fc;
    return it;
}
int aar_oracle_solve_ref(OracleHandle *h, double *z_io, int max_trace, double *trace, double *final_cost, int verbose) { return run_reference_slm(h, z_io, 0, max_trace, trace, final_cost, verbose); }
int aar_oracle_track_ref(OracleHandle *h, double *z_io, int max_trace, double *trace, double *final_cost) { return run_reference_slm(h, z_io, 1, max_trace, trace, final_cost, 0); }

/* One reference LM iteration's phases timed like sparselevmarq.h:352-366 for the CPU baseline:
 * runs `iters` calls of SLM::step from z and returns seconds. */
double aar_oracle_time_ref_steps(OracleHandle *h, const double *z0, int iters) {
    Mcm &m = h->mcm;
    typedef ucoslam::SparseLevMarq<double> SLM;
    SLM solver; SLM::Params p; p.maxIters = iters; p.min_average_step_error_diff = 0; p.minError = 0; solver.setParams(p); solver.v = 2;
    SLM::eVector z((long)m.num_vars);
    for (size_t i = 0; i < m.num_vars; i++) z((long)i) = z0[i];
    auto f = [&](const SLM::eVector &in, SLM::eVector &out) { eVec e(in.data(), in.data() + in.size()), err; m.error_function(e, err); out.resize((long)err.size()); std::memcpy(out.data(), err.data(), err.size() * sizeof(double)); };
    auto J = [&](const SLM::eVector &in, Eigen::SparseMatrix<double> &Jm) {
        eVec e(in.data(), in.data() + in.size()); std::vector<Triplet> t; m.jacobian_triplets(e, t);
        Jm.resize((long)m.num_point_xys, (long)m.num_vars);
        std::vector<Eigen::Triplet<double>> et; et.reserve(t.size());
        for (auto &q : t) et.push_back(Eigen::Triplet<double>(q.row, q.col, q.val));
        Jm.setFromTriplets(et.begin(), et.end());
    };
    m.hubberDelta = 10;
    solver.init(z, f);
    double t0 = omp_get_wtime();
    for (int i = 0; i < iters; i++) { solver.step(f, J); solver.prevErr = solver.currErr; }
    return omp_get_wtime() - t0;
}
#endif

int aar_oracle_omp_threads(void) { return omp_get_max_threads(); }
/* torch.distributed.run exports OMP_NUM_THREADS=1: the reference arm of bench.py sets its thread count itself */
void aar_oracle_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }

} /* extern "C" */
