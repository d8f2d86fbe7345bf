// aar_intrinsics.cuh — the camera-intrinsics block of the Jacobian: MultiCamMapper::obtain_transformation_derivs,
// parameter_type == intrinsics (/root/reference/libs/multicam_mapper.cpp:835-893).  The reference perturbs fx, cx, fy, cy and the
// five distortion coefficients of a camera by +-J_delta and differences the projections of every observation of that camera;
// project_marker (:608-649) never reads the distortion coefficients, so their five columns are structurally present and
// identically zero.  find_solution switches the block off (apps/find_solution.cpp:140); it is on in the reference's default Config.
//
// Layout.  The nine parameters of camera c are kept as TWO 6-wide pseudo-blocks of the reduced system, behind the camera and marker
// pose blocks: A = [fx cx fy cy k1 k2] (block nrc + nrm + 2c) and B = [p1 p2 k3 . . .] (block + 1; three padding columns).  Every
// column without a Jacobian — the distortion coefficients and the padding — is a zero row / column of J^T J whose diagonal is the
// damping mu alone, so its step is exactly 0 (as in the reference, whose LDLT sees the same zero columns).  With 6-wide blocks the
// whole Schur / Cholesky / back-substitution machinery runs unchanged; block A owns one more W slot per (frame, camera) pair.
// The host converts between this internal order and the reference's io_vec (intrinsics last, 9 per camera) at the ABI.
//
// Not a tuned path: one warp per (frame, camera) pair, 8 projections per observation, shuffles and atomics.  It runs next to the
// tensor-core assembly of the pose blocks and reads the same staged rows.
#pragma once

namespace aar {

// internal reduced vector -> camera matrix entries used by the projections (intrinsics_vec2mats, mcm.cpp:580-593)
__global__ void k_expand_intr(DevProblem p, const double *__restrict__ z, double *__restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    const double *s = z + p.col_intr0 + 12 * (size_t)c;
    dst[4 * c] = s[0]; dst[4 * c + 1] = s[1]; dst[4 * c + 2] = s[2]; dst[4 * c + 3] = s[3];
}

// central-difference numerators float(m - p+) - float(m - p-) of observation o for fx, cx, fy, cy: n[i][row]
__device__ __forceinline__ void intr_numerators(const DevProblem &p, long long o, int pair, int c, int m, double (&n)[4][8]) {
    Pose T1, Tm;
    load_pose(T1, p.pair_tab + (size_t)pair * PAIR_TAB);
    const bool mk_root = m == p.root_marker;
    if (!mk_root) load_pose(Tm, p.mk_tab + (size_t)m * MK_TAB);
    double c0[3], c1[3], t[3];
    if (mk_root) { c0[0] = T1.r[0]; c0[1] = T1.r[3]; c0[2] = T1.r[6]; c1[0] = T1.r[1]; c1[1] = T1.r[4]; c1[2] = T1.r[7]; t[0] = T1.t[0]; t[1] = T1.t[1]; t[2] = T1.t[2]; }
    else { compose_R01(T1.r, Tm.r, c0, c1); compose_t(T1.r, T1.t, Tm.t, t); }
    float raw[8];
    load8(p.raw_a, p.raw_b, o, raw);
    const double K0[4] = {p.intr[4 * c], p.intr[4 * c + 1], p.intr[4 * c + 2], p.intr[4 * c + 3]};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float pa[8], ps[8];
        Intr ka, ks;
        ka.fx = K0[0] + (i == 0 ? p.J_delta : 0.0); ka.cx = K0[1] + (i == 1 ? p.J_delta : 0.0); ka.fy = K0[2] + (i == 2 ? p.J_delta : 0.0); ka.cy = K0[3] + (i == 3 ? p.J_delta : 0.0);
        ks.fx = K0[0] - (i == 0 ? p.J_delta : 0.0); ks.cx = K0[1] - (i == 1 ? p.J_delta : 0.0); ks.fy = K0[2] - (i == 2 ? p.J_delta : 0.0); ks.cy = K0[3] - (i == 3 ? p.J_delta : 0.0);
        project(c0, c1, t, ka, p.h, pa);
        project(c0, c1, t, ks, p.h, ps);
#pragma unroll
        for (int r = 0; r < 8; r++) n[i][r] = (double)(raw[r] - pa[r]) - (double)(raw[r] - ps[r]);
    }
}

// parity hook: the intrinsics columns of every observation, Ji[o][i][row] = numerator / (2 delta)
__global__ void k_intr_dump(DevProblem p, double *__restrict__ Ji) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.N) return;
    const int cm = p.obs_cm[o];
    double n[4][8];
    if (obs_nojac(cm)) { for (int i = 0; i < 32; i++) Ji[32 * o + i] = 0.0; return; }
    intr_numerators(p, o, p.obs_pair[o], obs_cam(cm), obs_marker(cm), n);
    for (int i = 0; i < 4; i++) for (int r = 0; r < 8; r++) Ji[32 * o + 8 * i + r] = n[i][r] / (2 * p.J_delta);
}

// J^T J blocks and J^T r of the intrinsics columns: one warp per (frame, camera) pair, lanes over its observations.
//   W_i  = Ji^T Jf  -> the pair's own W slot (plain store, zero rows for k1, k2)        H_ii = Ji^T Ji, g_i = Ji^T r  -> reduced system
//   H_ci = Jc^T Ji  -> reduced system (camera pose x intrinsics of the same camera)      H_mi = Jm^T Ji -> reduced system, per observation
template <typename JT>
__global__ void __launch_bounds__(128) k_intr_assemble(DevProblem p, const int4 *__restrict__ pair_info, const int *__restrict__ pair_cam, const int *__restrict__ pair_slot_i,
                                                       const JT *__restrict__ Jn /* [N][JROW] */, const double *__restrict__ Hw /* [N][4] Huber weights or null */,
                                                       double s1, double s2, double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    const int lane = threadIdx.x & 31;
    const long long pr = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pr >= p.npairs) return;
    const int4 pi = pair_info[pr];
    const int cam = pair_cam[pr], n_r = p.n_r;
    const int bi = 6 * (p.nrc + p.nrm + 2 * cam);                                         // first column of pseudo-block A
    const bool act_c = p.opt_c && cam != p.root_cam;
    const int cb = 6 * (cam - (cam > p.root_cam ? 1 : 0));
    double Wi[4][6], Hii[10], gi[4], Hci[6][4];
#pragma unroll
    for (int i = 0; i < 4; i++) { gi[i] = 0; for (int d = 0; d < 6; d++) { Wi[i][d] = 0; Hci[d][i] = 0; } }
#pragma unroll
    for (int i = 0; i < 10; i++) Hii[i] = 0;
    for (int k = lane; k < pi.y; k += 32) {
        const long long o = pi.x + k;
        const int cm = p.obs_cm[o];
        if (obs_nojac(cm)) continue;
        const int m = obs_marker(cm);
        double n[4][8];
        intr_numerators(p, o, (int)pr, cam, m, n);
        const JT *row = Jn + (size_t)o * JROW;
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; q++) r[q] = (double)row[144 + q] * (Hw ? Hw[4 * o + (q >> 1)] : 1.0);
        int e = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int j = i; j < 4; j++, e++) { double s = 0; for (int q = 0; q < 8; q++) s = fma(n[i][q], n[j][q], s); Hii[e] += s; }
            double s = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) s = fma(n[i][q], r[q], s);
            gi[i] += s;
        }
#pragma unroll
        for (int d = 0; d < 6; d++) {
            double jf[8], jc[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { jf[q] = (double)row[96 + 8 * d + q]; jc[q] = (double)row[8 * d + q]; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double sf = 0, sc = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { sf = fma(n[i][q], jf[q], sf); sc = fma(jc[q], n[i][q], sc); }
                Wi[i][d] += sf; Hci[d][i] += sc;
            }
        }
        if (p.opt_m && m != p.root_marker) {
            double *dst = Hrr + (size_t)(6 * (p.nrc + m - (m > p.root_marker ? 1 : 0))) * n_r + bi;
#pragma unroll
            for (int d = 0; d < 6; d++) {
                double jm[8];
#pragma unroll
                for (int q = 0; q < 8; q++) jm[q] = (double)row[48 + 8 * d + q];
#pragma unroll
                for (int i = 0; i < 4; i++) { double s = 0; for (int q = 0; q < 8; q++) s = fma(jm[q], n[i][q], s); atomicAdd(dst + (size_t)d * n_r + i, s * s2); }
            }
        }
    }
    // ---- the pair's sums: butterfly over the lanes, then lane 0 publishes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) { gi[i] += __shfl_xor_sync(0xffffffffu, gi[i], o); for (int d = 0; d < 6; d++) { Wi[i][d] += __shfl_xor_sync(0xffffffffu, Wi[i][d], o); Hci[d][i] += __shfl_xor_sync(0xffffffffu, Hci[d][i], o); } }
#pragma unroll
        for (int i = 0; i < 10; i++) Hii[i] += __shfl_xor_sync(0xffffffffu, Hii[i], o);
    }
    if (lane != 0) return;
    const int slot = pair_slot_i[pr];
    if (p.opt_f && slot >= 0) {
        double *w = W + (size_t)slot * 36;
#pragma unroll
        for (int i = 0; i < 6; i++) for (int d = 0; d < 6; d++) w[i * 6 + d] = i < 4 ? Wi[i][d] * s2 : 0.0;
    }
    int e = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = i; j < 4; j++, e++) {
            atomicAdd(Hrr + (size_t)(bi + i) * n_r + bi + j, Hii[e] * s2);
            if (j != i) atomicAdd(Hrr + (size_t)(bi + j) * n_r + bi + i, Hii[e] * s2);
        }
        atomicAdd(gr + bi + i, gi[i] * s1);
        if (act_c) for (int d = 0; d < 6; d++) atomicAdd(Hrr + (size_t)(cb + d) * n_r + bi + i, Hci[d][i] * s2);
    }
}

} // namespace aar
