/*
 * init_oracle.cpp — CPU ORACLE of the initialisation path.  TEST INFRASTRUCTURE ONLY (same rules as mcm_oracle.cpp:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product never does).
 *
 * A from-scratch restatement of what the reference runs between `aruco.detections` + `calib.yml` and the
 * MultiCamMapper constructor (cited per function as init.cpp = /root/reference/libs/initializer.cpp and
 * ippe.cpp = /root/reference/3rdparty/aruco/aruco/ippe.cpp):
 *   Initializer::obtain_pose_estimations   init.cpp:364-419   -> aruco::solvePnP_ (ippe.cpp:118-126) per detection
 *   solvePoseOfCentredSquare               ippe.cpp:141-219   (homography, two rotations, translations, errors)
 *   fill_transformation_sets / _set        init.cpp:74-146
 *   find_best_transformation               init.cpp:156-205   (the O(n^2) consensus)
 *   make_mst / find_transforms_to_root     init.cpp:237-314
 *   init_transforms_cam / _marker / init_object_transforms  init.cpp:422-463
 * OpenCV calls on this path — cv::undistortPoints (no P), cv::Rodrigues, cv::Mat::inv, operator*, cv::reduce, cv::sqrt,
 * cv::sum — are restated over cv_shim.h and pinned against cv2 4.13 (tests/golden/make_golden_init.py,
 * tests/test_init_oracle_pin.py).  The reference ships no tests for this path either: PARITY UNPINNED beyond those vectors.
 *
 * One documented extension: `consensus_max` > 0 restricts a consensus of n > consensus_max candidate
 * transformations to the consensus_max candidates at list positions floor(k * n / consensus_max); 0 is the reference.
 */
#include <cstdint>
#include <cstdio>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <tuple>
#include <vector>

#include "cv_shim.h"
#include "../include/aar_acos.h"

namespace {

struct PoseEst { M4 T; double err; };                                      /* std::pair<cv::Mat, double> */
typedef std::map<int, std::map<int, std::vector<PoseEst>>> PoseMap;       /* [marker][cam] or [cam][marker] -> candidates */
struct Tri { M4 T, T1inv, T2inv; double err; };                           /* std::tuple<cv::Mat, cv::Mat, cv::Mat, double> */
typedef std::map<int, std::map<int, std::vector<Tri>>> TriSets;
typedef std::map<int, std::map<int, std::pair<M4, double>>> Best;

double sin_mode(double x) { if (g_sincos_mode == 0) return std::sin(x); double s, c; aar_sincos(x, &s, &c); return s; }
double acos_mode(double x) { return g_sincos_mode == 0 ? std::acos(x) : aar_acos(x); }

/* cv::undistortPoints(src, dst, K, dist) with R and P empty (ippe.cpp:164): normalised coordinates, float32 out */
void undistort_normalised(float u_in, float v_in, const double K[9], const double k[5], float *uo, float *vo) {
    double fx = K[0], fy = K[4], ifx = 1. / fx, ify = 1. / fy, cx = K[2], cy = K[5];
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
    /* RR = I: xx = 1*x + 0*y + 0, ww = 1/(0*x + 0*y + 1) — exact for finite x, y */
    *uo = (float)x; *vo = (float)y;
}

/* homographyFromSquarePoints (ippe.cpp:535-579): q = corners (normalised, float), hl = half length */
void square_homography(const float *q, double hl, double H[9]) {
    const double ax = -q[0], ay = -q[1], bx = -q[2], by = -q[3], cx = -q[4], cy = -q[5], dx = -q[6], dy = -q[7];
    const double di = -1 / (hl * (ax * by - bx * ay - ax * dy + bx * cy - cx * by + dx * ay + cx * dy - dx * cy));
    H[0] = di * (ax * cx * by - bx * cx * ay - ax * dx * by + bx * dx * ay - ax * cx * dy + ax * dx * cy + bx * cx * dy - bx * dx * cy);
    H[1] = di * (ax * bx * cy - ax * cx * by - ax * bx * dy + bx * dx * ay + ax * cx * dy - cx * dx * ay - bx * dx * cy + cx * dx * by);
    H[2] = di * hl * (ax * bx * cy - bx * cx * ay - ax * bx * dy + ax * dx * by - ax * dx * cy + cx * dx * ay + bx * cx * dy - cx * dx * by);
    H[3] = di * (ax * by * cy - bx * ay * cy - ax * by * dy + bx * ay * dy - cx * ay * dy + dx * ay * cy + cx * by * dy - dx * by * cy);
    H[4] = di * (bx * ay * cy - cx * ay * by - ax * by * dy + dx * ay * by + ax * cy * dy - dx * ay * cy - bx * cy * dy + cx * by * dy);
    H[5] = di * hl * (ax * by * cy - cx * ay * by - bx * ay * dy + dx * ay * by - ax * cy * dy + cx * ay * dy + bx * cy * dy - dx * by * cy);
    H[6] = -di * (ax * cy - cx * ay - ax * dy - bx * cy + cx * by + dx * ay + bx * dy - dx * by);
    H[7] = di * (ax * by - bx * ay - ax * cy + cx * ay + bx * dy - dx * by - cx * dy + dx * cy);
    H[8] = 1.0;
}

/* IPPComputeRotations (ippe.cpp:426-533) */
void ippe_rotations(double j00, double j01, double j10, double j11, double p, double q, double R1[9], double R2[9]) {
    const double s = std::sqrt(p * p + q * q + 1), t = std::sqrt(p * p + q * q);
    const double costh = 1 / s, sinth = std::sqrt(1 - 1 / (s * s));
    const double k0 = p / t, k1 = q / t, k0s = k0 * k0, k1s = k1 * k1;
    double rv[9];
    rv[0] = (costh - 1) * k0s + 1;  rv[1] = k0 * k1 * (costh - 1);     rv[2] = k0 * sinth;
    rv[3] = k0 * k1 * (costh - 1);  rv[4] = (costh - 1) * k1s + 1;     rv[5] = k1 * sinth;
    rv[6] = -k0 * sinth;            rv[7] = -k1 * sinth;               rv[8] = (costh - 1) * (k0s + k1s) + 1;
    const double b00 = rv[0] - p * rv[6], b01 = rv[1] - p * rv[7], b10 = rv[3] - q * rv[6], b11 = rv[4] - q * rv[7];
    const double dtinv = 1.0 / ((b00 * b11 - b01 * b10));
    const double bi00 = dtinv * b11, bi01 = -dtinv * b01, bi10 = -dtinv * b10, bi11 = dtinv * b00;
    const double a00 = bi00 * j00 + bi01 * j10, a01 = bi00 * j01 + bi01 * j11, a10 = bi10 * j00 + bi11 * j10, a11 = bi10 * j01 + bi11 * j11;
    const double ata00 = a00 * a00 + a01 * a01, ata01 = a00 * a10 + a01 * a11, ata11 = a10 * a10 + a11 * a11;
    const double gamma = std::sqrt(0.5 * (ata00 + ata11 + std::sqrt((ata00 - ata11) * (ata00 - ata11) + 4.0 * ata01 * ata01)));
    const double r00 = a00 / gamma, r01 = a01 / gamma, r10 = a10 / gamma, r11 = a11 / gamma;
    const double b0 = std::sqrt(-(r00 * r00) - r10 * r10 + 1);
    double b1 = std::sqrt(-(r01 * r01) - r11 * r11 + 1);
    const double sp = (-r00 * r01 - r10 * r11);
    if (sp < 0) b1 = -b1;
    /* third column: two cross-product coefficient sets (R1 with (b0, b1), R2 with (-b0, -b1)) and the shared in-plane determinant */
    const double u1 = b1 * r10 - b0 * r11, v1 = b0 * r01 - b1 * r00, w = r00 * r11 - r01 * r10;
    const double u2 = b0 * r11 - b1 * r10, v2 = b1 * r00 - b0 * r01;
    for (int i = 0; i < 3; i++) {
        const double x = rv[3 * i], y = rv[3 * i + 1], z = rv[3 * i + 2];
        R1[3 * i + 0] = (r00) * x + (r10) * y + (b0) * z;
        R1[3 * i + 1] = (r01) * x + (r11) * y + (b1) * z;
        R1[3 * i + 2] = u1 * x + v1 * y + w * z;
        R2[3 * i + 0] = (r00) * x + (r10) * y + (-b0) * z;
        R2[3 * i + 1] = (r01) * x + (r11) * y + (-b1) * z;
        R2[3 * i + 2] = u2 * x + v2 * y + w * z;
    }
}

/* IPPComputeTranslation (ippe.cpp:380-424): mp = model points (float x, y, z), q = normalised image points */
void ippe_translation(const float mp[4][3], const float *q, const double R[9], double t[3]) {
    const double ATA00 = 4, ATA11 = 4;
    double ATA02 = 0, ATA12 = 0, ATA20 = 0, ATA21 = 0, ATA22 = 0, ATb0 = 0, ATb1 = 0, ATb2 = 0;
    for (int i = 0; i < 4; i++) {
        const double rx = R[0] * mp[i][0] + R[1] * mp[i][1] + R[2] * mp[i][2];
        const double ry = R[3] * mp[i][0] + R[4] * mp[i][1] + R[5] * mp[i][2];
        const double rz = R[6] * mp[i][0] + R[7] * mp[i][1] + R[8] * mp[i][2];
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

/* IPPEvalReprojectionError (ippe.cpp:296-330): float arithmetic on double products narrowed to float */
float ippe_reproj_error(const double R[9], const double t[3], const float mp[4][3], const float *q) {
    float err = 0;
    for (int i = 0; i < 4; i++) {
        const float px = static_cast<float>(R[0] * mp[i][0]) + static_cast<float>(R[1] * mp[i][1]) + static_cast<float>(R[2] * mp[i][2] + t[0]);
        const float py = static_cast<float>(R[3] * mp[i][0]) + static_cast<float>(R[4] * mp[i][1]) + static_cast<float>(R[5] * mp[i][2] + t[1]);
        const float pz = static_cast<float>(R[6] * mp[i][0]) + static_cast<float>(R[7] * mp[i][1]) + static_cast<float>(R[8] * mp[i][2] + t[2]);
        const float dx = px / pz - q[2 * i], dy = py / pz - q[2 * i + 1];
        err = err + std::sqrt(dx * dx + dy * dy);
    }
    return err;
}

/* IPPERot2vec (ippe.cpp:332-358) followed by getRTMatrix(..., CV_32F) (ippe.cpp:40-97) and the Initializer's
 * convertTo(CV_64FC1) (init.cpp:403, 409): a 4x4 double matrix whose entries are float32 values */
M4 ippe_pose_matrix(const double R[9], const double t[3]) {
    const double trace = R[0] + R[4] + R[8];
    const double w_norm = acos_mode((trace - 1.0) / 2.0);
    const double d = 1 / (2 * sin_mode(w_norm)) * w_norm;
    double rvec[3] = {0, 0, 0};
    if (!(w_norm < std::numeric_limits<double>::epsilon())) {
        rvec[0] = d * (R[7] - R[5]); rvec[1] = d * (R[2] - R[6]); rvec[2] = d * (R[3] - R[1]);
    }
    double R33[9];
    rodrigues_vec2mat(rvec, R33);
    M4 m = eye4();
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m.a[4 * i + j] = R33[3 * i + j]; m.a[4 * i + 3] = t[i]; }
    for (int i = 0; i < 16; i++) m.a[i] = (double)(float)m.a[i];
    return m;
}

/* aruco::solvePnP_(float size, imgPoints, K, dist) (ippe.cpp:118-126) -> solvePoseOfCentredSquare (ippe.cpp:141-219) */
void solve_pnp(float size, const float raw[8], const double K[9], const double dist[5], PoseEst out[2]) {
    float mp[4][3] = {{-size / 2.0f, size / 2.0f, 0}, {size / 2.0f, size / 2.0f, 0}, {size / 2.0f, -size / 2.0f, 0}, {-size / 2.0f, -size / 2.0f, 0}};
    float q[8];
    for (int i = 0; i < 4; i++) undistort_normalised(raw[2 * i], raw[2 * i + 1], K, dist, &q[2 * i], &q[2 * i + 1]);
    double H[9];
    square_homography(q, size / 2.0f, H);
    const double j00 = H[0] - H[6] * H[2], j01 = H[1] - H[7] * H[2], j10 = H[3] - H[6] * H[5], j11 = H[4] - H[7] * H[5];
    double Ra[9], Rb[9], ta[3], tb[3];
    ippe_rotations(j00, j01, j10, j11, H[2], H[5], Ra, Rb);
    ippe_translation(mp, q, Ra, ta);
    ippe_translation(mp, q, Rb, tb);
    const float ea = ippe_reproj_error(Ra, ta, mp, q), eb = ippe_reproj_error(Rb, tb, mp, q);
    if (ea < eb) { out[0].T = ippe_pose_matrix(Ra, ta); out[0].err = ea; out[1].T = ippe_pose_matrix(Rb, tb); out[1].err = eb; }
    else         { out[0].T = ippe_pose_matrix(Rb, tb); out[0].err = eb; out[1].T = ippe_pose_matrix(Ra, ta); out[1].err = ea; }
}

/* Initializer::find_best_transformation (init.cpp:156-205) */
int consensus(double marker_size, const std::vector<Tri> &sol, int consensus_max, double &weight) {
    const double h = marker_size / 2;
    M4 pts;
    const double px[4] = {-h, h, h, -h}, py[4] = {h, h, -h, -h};
    for (int c = 0; c < 4; c++) { pts.a[0 * 4 + c] = px[c]; pts.a[1 * 4 + c] = py[c]; pts.a[2 * 4 + c] = 0; pts.a[3 * 4 + c] = 1; }
    const int64_t n = (int64_t)sol.size();
    const int64_t m = (consensus_max > 0 && n > consensus_max) ? consensus_max : n;
    double min_error = std::numeric_limits<double>::max();
    int min_index = -1;
    for (int64_t ii = 0; ii < m; ii++) {
        const int64_t i = (m == n) ? ii : ii * n / m;
        double curr = 0;
        for (int64_t jj = 0; jj < m; jj++) {
            const int64_t j = (m == n) ? jj : jj * n / m;
            const M4 p2 = mul44(mul44(mul44(sol[j].T2inv, sol[i].T), sol[j].T1inv), pts);
            double e[4];
            for (int c = 0; c < 4; c++) {                       /* diff.mul(diff), cv::reduce(REDUCE_SUM over rows 0..2), cv::sqrt */
                const double d0 = pts.a[c] - p2.a[c], d1 = pts.a[4 + c] - p2.a[4 + c], d2 = pts.a[8 + c] - p2.a[8 + c];
                e[c] = std::sqrt((d0 * d0 + d1 * d1) + d2 * d2);
            }
            curr += ((e[0] + e[1]) + e[2]) + e[3];              /* cv::sum of 4 doubles */
        }
        if (curr < min_error) { min_index = (int)i; min_error = curr; weight = min_error; }
    }
    return min_index;
}

struct Init {
    int num_cams = 0, num_frames = 0, consensus_max = 0, min_detections = 2;
    double marker_size = 0, threshold = 2.0;
    std::vector<double> K, dist;                               /* cam_configs[cam] */
    std::set<int> excluded;
    /* detections[frame][cam] = indices into the flat detection arrays, detection order */
    std::vector<std::vector<std::vector<int64_t>>> detections;
    std::vector<int> det_marker; std::vector<float> det_xy;
    /* per flat detection: what obtain_pose_estimations produced (ncand 0 = frame skipped / camera excluded) */
    std::vector<PoseEst> est; std::vector<uint8_t> ncand;
    std::set<int> cam_ids, marker_ids;
    std::map<int, PoseMap> frame_poses_cam, frame_poses_marker;
    std::map<int, std::map<int, std::vector<int64_t>>> frame_cam_markers;
    int root_cam = -1, root_marker = -1;
    std::map<int, M4> to_root_cam, to_root_marker, object_T;
    std::map<int, std::map<int, std::pair<int64_t, double>>> best_cam_info, best_marker_info;   /* [id1][id2] -> (list length, weight) */

    /* init.cpp:364-419 */
    void obtain_pose_estimations() {
        frame_cam_markers.clear(); frame_poses_cam.clear(); frame_poses_marker.clear();
        est.assign(det_marker.size() * 2, PoseEst()); ncand.assign(det_marker.size(), 0);
        for (int f = 0; f < (int)detections.size(); f++) {
            PoseMap pe_marker, pe_cam;
            int num = 0;
            for (int cam = 0; cam < (int)detections[f].size(); cam++) if (!excluded.count(cam)) num += (int)detections[f][cam].size();
            if (!(num >= min_detections)) continue;
            for (int cam = 0; cam < (int)detections[f].size(); cam++) {
                if (excluded.count(cam)) continue;
                if (detections[f][cam].size() < 1) continue;
                cam_ids.insert(cam);
                auto &cm = frame_cam_markers[f][cam];
                for (int64_t d : detections[f][cam]) {
                    const int id = det_marker[d];
                    marker_ids.insert(id);
                    cm.push_back(d);
                    PoseEst s[2];
                    solve_pnp((float)marker_size, &det_xy[8 * d], &K[9 * cam], &dist[5 * cam], s);
                    est[2 * d] = s[0]; est[2 * d + 1] = s[1]; ncand[d] = 1;
                    pe_cam[id][cam].push_back(s[0]); pe_marker[cam][id].push_back(s[0]);
                    if (s[1].err / s[0].err < threshold) { ncand[d] = 2; pe_cam[id][cam].push_back(s[1]); pe_marker[cam][id].push_back(s[1]); }
                }
            }
            frame_poses_cam[f] = pe_cam; frame_poses_marker[f] = pe_marker;
        }
    }

    /* init.cpp:117-146 */
    static void fill_transformation_sets(bool cams, const PoseMap &pe, TriSets &sets) {
        for (auto it = pe.begin(); it != pe.end(); ++it) {
            const auto &objects = it->second;
            if (objects.size() > 1)
                for (auto it1 = objects.begin(); it1 != objects.end(); ++it1)
                    for (size_t i = 0; i < it1->second.size(); i++)
                        for (auto it2 = std::next(it1); it2 != objects.end(); ++it2)
                            for (size_t j = 0; j < it2->second.size(); j++) {
                                const PoseEst &p1 = it1->second[i], &p2 = it2->second[j];
                                Tri t; t.err = p1.err * p2.err;
                                if (cams) { t.T = mul44(p2.T, inv44(p1.T)); t.T1inv = p1.T; t.T2inv = inv44(p2.T); }
                                else      { t.T = mul44(inv44(p2.T), p1.T); t.T1inv = inv44(p1.T); t.T2inv = p2.T; }
                                sets[it1->first][it2->first].push_back(t);
                            }
        }
    }

    /* init.cpp:207-235 */
    void find_best_transformations(const TriSets &sets, Best &best, std::map<int, std::map<int, std::pair<int64_t, double>>> &info) const {
        for (auto it1 = sets.begin(); it1 != sets.end(); ++it1)
            for (auto it2 = it1->second.begin(); it2 != it1->second.end(); ++it2) {
                double w = 0;
                const int idx = consensus(marker_size, it2->second, consensus_max, w);
                best[it1->first][it2->first] = std::make_pair(it2->second[idx].T, w);
                info[it1->first][it2->first] = std::make_pair((int64_t)it2->second.size(), w);
            }
    }

    /* init.cpp:237-288 — the reference's Prim variant on edge weights (a disconnected node keeps parent -1) */
    static void make_mst(int start, const std::set<int> &ids, const Best &adj, std::map<int, std::set<int>> &children) {
        struct Node { int id; double distance; int parent; };
        std::map<int, Node> outside;
        for (int id : ids) outside[id] = Node{id, id == start ? 0.0 : std::numeric_limits<double>::max(), -1};
        while (!outside.empty()) {
            auto mn = outside.begin();
            for (auto it = outside.begin(); it != outside.end(); ++it) if (it->second.distance < mn->second.distance) mn = it;
            for (auto it = outside.begin(); it != outside.end(); ++it) {
                bool have = false; double error = std::numeric_limits<double>::max();
                const int a = mn->first, b = it->first;
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                if (a != b) {
                    auto r = adj.find(lo);
                    if (r != adj.end()) { auto c = r->second.find(hi); if (c != r->second.end()) { have = true; error = c->second.second; } }
                }
                if (have && error < it->second.distance) {
                    it->second.distance = error;
                    if (it->second.parent != -1) children[it->second.parent].erase(b);
                    children[a].insert(b);
                    it->second.parent = a;
                }
            }
            outside.erase(mn);
        }
    }

    /* init.cpp:290-314 */
    static void find_transforms_to_root(int root, const std::map<int, std::set<int>> &children, const Best &best, std::map<int, M4> &out) {
        out[root] = eye4();
        std::queue<int> q; q.push(root);
        while (!q.empty()) {
            const int parent = q.front();
            auto ch = children.find(parent);
            if (ch != children.end())
                for (int child : ch->second) {
                    if (child < parent) out[child] = best.at(child).at(parent).first;
                    else out[child] = inv44(best.at(parent).at(child).first);
                    if (parent != root) out[child] = mul44(out[parent], out[child]);
                    q.push(child);
                }
            q.pop();
        }
    }

    /* init.cpp:422-449 */
    void init_transforms_rig() {
        TriSets sets_cam, sets_marker;
        for (int f = 0; f < (int)detections.size(); f++) { auto it = frame_poses_cam.find(f); if (it != frame_poses_cam.end()) fill_transformation_sets(true, it->second, sets_cam); }
        Best best_cam; find_best_transformations(sets_cam, best_cam, best_cam_info);
        std::map<int, std::set<int>> cam_tree;
        root_cam = *cam_ids.begin();
        make_mst(root_cam, cam_ids, best_cam, cam_tree);
        find_transforms_to_root(root_cam, cam_tree, best_cam, to_root_cam);
        for (int f = 0; f < (int)detections.size(); f++) { auto it = frame_poses_marker.find(f); if (it != frame_poses_marker.end()) fill_transformation_sets(false, it->second, sets_marker); }
        Best best_marker; find_best_transformations(sets_marker, best_marker, best_marker_info);
        std::map<int, std::set<int>> marker_tree;
        root_marker = *marker_ids.begin();
        make_mst(root_marker, marker_ids, best_marker, marker_tree);
        find_transforms_to_root(root_marker, marker_tree, best_marker, to_root_marker);
    }

    /* init.cpp:74-115 + 451-463 */
    void init_object_transforms() {
        object_T.clear();
        for (auto it = frame_poses_cam.begin(); it != frame_poses_cam.end(); ++it) {
            std::vector<Tri> set;
            for (auto mk = it->second.begin(); mk != it->second.end(); ++mk)
                for (auto cm = mk->second.begin(); cm != mk->second.end(); ++cm) {
                    M4 T_mr = eye4(), T_rm = eye4(), T_cr = eye4(), T_rc = eye4();
                    auto fm = to_root_marker.find(mk->first);
                    if (fm != to_root_marker.end()) { T_mr = fm->second; T_rm = inv44(T_mr); }
                    auto fc = to_root_cam.find(cm->first);
                    if (fc != to_root_cam.end()) { T_cr = fc->second; T_rc = inv44(T_cr); }
                    for (const PoseEst &pe : cm->second) {
                        const M4 T_cm = inv44(pe.T);
                        Tri t; t.T = mul44(mul44(T_cr, pe.T), T_rm); t.T1inv = mul44(T_mr, T_cm); t.T2inv = T_rc; t.err = pe.err;
                        set.push_back(t);
                    }
                }
            double w = 0;
            const int idx = consensus(marker_size, set, consensus_max, w);
            if (idx >= 0) object_T[it->first] = set[idx].T;
        }
    }
};

} // namespace

extern "C" {

/* Initializer(double marker_s, cam_configs, excluded) + set_detections (init.cpp:58-62, 22-24); detections are flat, in
 * aruco.detections file order (frame, camera, detection order) */
void *aar_init_oracle_create(int num_cams, const double *K, const double *dist, double marker_size, int num_frames, int64_t ndet,
                             const int *det_frame, const int *det_cam, const int *det_marker, const float *det_xy,
                             const uint8_t *excluded, double threshold, int consensus_max) {
    Init *h = new Init;
    h->num_cams = num_cams; h->num_frames = num_frames; h->marker_size = marker_size; h->threshold = threshold; h->consensus_max = consensus_max;
    h->K.assign(K, K + 9 * num_cams); h->dist.assign(dist, dist + 5 * num_cams);
    if (excluded) for (int c = 0; c < num_cams; c++) if (excluded[c]) h->excluded.insert(c);
    h->detections.assign(num_frames, std::vector<std::vector<int64_t>>(num_cams));
    h->det_marker.assign(det_marker, det_marker + ndet); h->det_xy.assign(det_xy, det_xy + 8 * ndet);
    for (int64_t d = 0; d < ndet; d++) h->detections[det_frame[d]][det_cam[d]].push_back(d);
    return h;
}
void aar_init_oracle_destroy(void *hv) { delete (Init *)hv; }
void aar_init_oracle_obtain_pose_estimations(void *hv) { ((Init *)hv)->obtain_pose_estimations(); }
void aar_init_oracle_init_transforms(void *hv) { Init *h = (Init *)hv; h->init_transforms_rig(); h->init_object_transforms(); }   /* init.cpp:465-469 */
void aar_init_oracle_init_object_transforms(void *hv) { ((Init *)hv)->init_object_transforms(); }

/* per flat detection: candidate poses [ndet][2][16], errors [ndet][2], candidates kept (0, 1 or 2) */
void aar_init_oracle_get_estimations(void *hv, double *T, double *err, uint8_t *ncand) {
    Init *h = (Init *)hv;
    for (size_t d = 0; d < h->ncand.size(); d++) {
        ncand[d] = h->ncand[d];
        for (int k = 0; k < 2; k++) { std::memcpy(T + (2 * d + k) * 16, h->est[2 * d + k].T.a, 128); err[2 * d + k] = h->est[2 * d + k].err; }
    }
}
static int copy_map(const std::map<int, M4> &m, int cap, int *ids, double *T) {
    int n = 0;
    for (auto &kv : m) { if (n < cap) { if (ids) ids[n] = kv.first; if (T) std::memcpy(T + 16 * n, kv.second.a, 128); } n++; }
    return n;
}
static int copy_set(const std::set<int> &s, int cap, int *ids) { int n = 0; for (int v : s) { if (n < cap && ids) ids[n] = v; n++; } return n; }
int aar_init_oracle_cam_ids(void *hv, int cap, int *ids) { return copy_set(((Init *)hv)->cam_ids, cap, ids); }
int aar_init_oracle_marker_ids(void *hv, int cap, int *ids) { return copy_set(((Init *)hv)->marker_ids, cap, ids); }
int aar_init_oracle_root_cam(void *hv) { return ((Init *)hv)->root_cam; }
int aar_init_oracle_root_marker(void *hv) { return ((Init *)hv)->root_marker; }
int aar_init_oracle_transforms_to_root_cam(void *hv, int cap, int *ids, double *T) { return copy_map(((Init *)hv)->to_root_cam, cap, ids, T); }
int aar_init_oracle_transforms_to_root_marker(void *hv, int cap, int *ids, double *T) { return copy_map(((Init *)hv)->to_root_marker, cap, ids, T); }
int aar_init_oracle_object_transforms(void *hv, int cap, int *ids, double *T) { return copy_map(((Init *)hv)->object_T, cap, ids, T); }
/* set_transforms_to_root_cam / _marker (init.cpp:14-20) — the track app's flow */
void aar_init_oracle_set_rig(void *hv, int nc, const int *cam_ids, const double *cam_T, int nm, const int *marker_ids, const double *marker_T) {
    Init *h = (Init *)hv; h->to_root_cam.clear(); h->to_root_marker.clear();
    for (int i = 0; i < nc; i++) { M4 m; std::memcpy(m.a, cam_T + 16 * i, 128); h->to_root_cam[cam_ids[i]] = m; }
    for (int i = 0; i < nm; i++) { M4 m; std::memcpy(m.a, marker_T + 16 * i, 128); h->to_root_marker[marker_ids[i]] = m; }
}
/* edges of the consensus graphs: (id1, id2, list length, weight) */
int aar_init_oracle_edges(void *hv, int cams, int cap, int *id1, int *id2, int64_t *len, double *weight) {
    Init *h = (Init *)hv; int n = 0;
    for (auto &a : (cams ? h->best_cam_info : h->best_marker_info)) for (auto &b : a.second) {
        if (n < cap) { id1[n] = a.first; id2[n] = b.first; len[n] = b.second.first; weight[n] = b.second.second; }
        n++;
    }
    return n;
}

/* pinning hooks: one IPPE solve and one consensus */
void aar_init_oracle_solve_pnp(float size, const float *raw8, const double *K, const double *dist, double *T /* [2][16] */, double *err /* [2] */) {
    PoseEst s[2]; solve_pnp(size, raw8, K, dist, s);
    for (int k = 0; k < 2; k++) { std::memcpy(T + 16 * k, s[k].T.a, 128); err[k] = s[k].err; }
}
void aar_init_oracle_ippe_raw(float size, const float *raw8, const double *K, const double *dist, float *q8, double *H9, double *Ra, double *ta, double *Rb, double *tb, float *errs) {
    float mp[4][3] = {{-size / 2.0f, size / 2.0f, 0}, {size / 2.0f, size / 2.0f, 0}, {size / 2.0f, -size / 2.0f, 0}, {-size / 2.0f, -size / 2.0f, 0}};
    for (int i = 0; i < 4; i++) undistort_normalised(raw8[2 * i], raw8[2 * i + 1], K, dist, &q8[2 * i], &q8[2 * i + 1]);
    square_homography(q8, size / 2.0f, H9);
    ippe_rotations(H9[0] - H9[6] * H9[2], H9[1] - H9[7] * H9[2], H9[3] - H9[6] * H9[5], H9[4] - H9[7] * H9[5], H9[2], H9[5], Ra, Rb);
    ippe_translation(mp, q8, Ra, ta); ippe_translation(mp, q8, Rb, tb);
    errs[0] = ippe_reproj_error(Ra, ta, mp, q8); errs[1] = ippe_reproj_error(Rb, tb, mp, q8);
}
int aar_init_oracle_consensus(double marker_size, int64_t n, const double *T, const double *T1inv, const double *T2inv, int consensus_max, double *weight) {
    std::vector<Tri> s(n);
    for (int64_t i = 0; i < n; i++) { std::memcpy(s[i].T.a, T + 16 * i, 128); std::memcpy(s[i].T1inv.a, T1inv + 16 * i, 128); std::memcpy(s[i].T2inv.a, T2inv + 16 * i, 128); s[i].err = 0; }
    double w = 0; const int idx = consensus(marker_size, s, consensus_max, w); *weight = w; return idx;
}
double aar_init_oracle_acos(double x) { return aar_acos(x); }

} /* extern "C" */
