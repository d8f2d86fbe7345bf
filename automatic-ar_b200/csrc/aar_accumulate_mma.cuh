// aar_accumulate_mma.cuh — J^T J / J^T r assembly on the FP64 tensor cores (mma.sync.m8n8k4.f64): two OPT-IN formulations of
// k_jac_accumulate (aar_jacobian.cuh; sparselevmarq.h:362-367 in /root/reference/libs/), selected with AAR_ACC_MMA=1 / =2 at
// aar_problem_create.  Both are parity-tested (tests/test_gpu_parity.py) and, at the end of round 1, within 10 % of the
// lane-per-observation kernel's time — profiles/r1_notes.md has the measurements and what round 2 should do with them.
// They read the central-difference numerators as one row of 144 per observation (k_jac_project<JT, ROWS = true>).
#pragma once

namespace aar {

// ------------------------------------------------------------------------------------------------
// K2'  k_jac_accumulate_mma: the same sums on the FP64 tensor cores (mma.sync.m8n8k4.f64), one WARP per observation at
// a time instead of one lane.  A block product P = Xa^T Xb (6x6, 8 residual rows deep) is two DMMA k-steps with
// fragment element (g = lane / 4, q = lane % 4, step s) = X[dof g][row 2q + s] — the same register serves as A fragment
// (row g of Xa) and as B fragment (column g of Xb); fragment row / column 6 of a B operand carries the residual, so the
// gradient J^T r falls out of the same instruction as column 6 of the result.  What this buys over k_jac_accumulate:
//   * the result fragment IS the "lane v owns value v" layout the atomics want: no transposition scratch, no warp reduce;
//   * sums over runs of observations (frame, (frame, camera)) accumulate inside the DMMA accumulators;
//   * ~90 registers instead of 255: 16 warps per SM, every branch warp-uniform.
// A CTA owns whole frames (batches of consecutive frames, <= ACC2_OBS_CAP observations, planned on the host): Hff / gf and
// the W blocks of a batch are summed in shared-memory windows and leave with plain coalesced stores — no global atomics and
// no zeroing pass for them.  Hcc / Hmm / Hcm live in CTA-lifetime shared accumulators as before (pairs that do not fit: RED).
constexpr int ACC2_THREADS = 512, ACC2_WARPS = ACC2_THREADS / 32;
constexpr int ACC2_OBS_CAP = 512, ACC2_FRAME_CAP = 64, ACC2_SLOT_MIN = 160;

struct Acc2Plan { int hcm_smem, nbatch, win_slots, win_frames; const int *batch_f; /* [nbatch + 1] first frame of each batch */ const int *frame_obs_ptr; double s1, s2; };

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// Shared-memory double accumulation by 32-bit shared addresses.  addr[b] += val[b] for the entries of this lane with
// on[b] != 0 (distinct addresses): loads, adds and compare-and-swaps go out as straight-line batches, predicated per entry
// inside the PTX (no branches, no shared-memory traffic for a lane without an element); a lost race — another warp on the
// same accumulator — is finished by a CAS loop.
__device__ __forceinline__ unsigned long long lds_b64_if(unsigned a, int on) {
    unsigned long long v = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.shared.b64 %0, [%1];\n\t}" : "+l"(v) : "r"(a), "r"(on) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long cas_b64_if(unsigned a, unsigned long long cmp, unsigned long long val, int on) {
    unsigned long long o = cmp;                      // a lane that does not take part reports "no race lost"
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %4, 0;\n\t@p atom.shared.cas.b64 %0, [%1], %2, %3;\n\t}" : "+l"(o) : "r"(a), "l"(cmp), "l"(val), "r"(on) : "memory");
    return o;
}
__device__ __noinline__ void smem_add_slow(unsigned a, double v) {
    unsigned long long old = lds_b64_if(a, 1);
    for (;;) {
        const unsigned long long want = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)old) + v);
        const unsigned long long got = cas_b64_if(a, old, want, 1);
        if (got == old) return;
        old = got;
    }
}
template <int NB>
__device__ __forceinline__ void smem_add_batch(const unsigned (&addr)[NB], const double (&val)[NB], const int (&on)[NB]) {
    unsigned long long old[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) old[b] = lds_b64_if(addr[b], on[b]);
    unsigned long long got[NB], diff = 0;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const unsigned long long want = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)old[b]) + val[b]);
        got[b] = cas_b64_if(addr[b], old[b], want, on[b]);
        diff |= got[b] ^ old[b];
    }
    if (diff) {                                      // rare
#pragma unroll
        for (int b = 0; b < NB; b++) if (got[b] != old[b]) smem_add_slow(addr[b], val[b]);
    }
}
template <typename JT> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

template <typename JT>
__global__ void __launch_bounds__(ACC2_THREADS, 1) k_jac_accumulate_mma(DevProblem p, Acc2Plan pl, const JT *__restrict__ Jn /* [N][144] */, const double *__restrict__ Rv /* [N][8] */,
                                                                       double *__restrict__ Hf, double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    typedef typename Vec2<JT>::type V2;
    extern __shared__ __align__(16) double sAcc[];
    double *sHcc = sAcc;                                                      // [nrc][27]   (sHmm follows: blocks nrc.. are markers)
    double *sHmm = sHcc + p.nrc * 27;                                         // [nrm][27]
    double *sHcm = sHmm + p.nrm * 27;                                         // [hcm_smem][36]: the first pairs in (camera, marker) order
    double *sW = sHcm + (size_t)pl.hcm_smem * 36;                             // [win_slots][36]  W blocks of the current batch
    double *sHf = sW + (size_t)pl.win_slots * 36;                             // [win_frames][27] frame blocks of the current batch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int n_all = (int)(sHf - sAcc) + pl.win_frames * 27;
    for (int i = tid; i < n_all; i += ACC2_THREADS) sAcc[i] = 0.0;
    __syncthreads();
    const int n_r = p.n_r;
    const double s1 = pl.s1, s2 = pl.s2;
    // Shared addresses of this lane's two result elements (row g, columns 2q and 2q + 1) inside block 0 of each table, and
    // whether the element exists there (rows 6, 7, columns 6 / 7 and the lower triangle of a packed symmetric block do not)
    const unsigned aHcc = (unsigned)__cvta_generic_to_shared(sHcc), aHmm = (unsigned)__cvta_generic_to_shared(sHmm), aHcm = (unsigned)__cvta_generic_to_shared(sHcm);
    const unsigned aW = (unsigned)__cvta_generic_to_shared(sW), aHf = (unsigned)__cvta_generic_to_shared(sHf);
    unsigned eW[2], eHmm[2], eHcm[2], eHf[2], eHccU[2], eHccG[2]; int m36[2], m27[2], mU[2], mG[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int j = 2 * q + e;
        const int idx36 = (g < 6 && j < 6) ? g * 6 + j : -1;                                          // full 6x6 block, row-major
        const int idxU = (g < 6 && j < 6 && j >= g) ? g * 6 - g * (g - 1) / 2 + (j - g) : -1;       // packed upper triangle, as prod27
        const int idxG = (g < 6 && j == 6) ? 21 + g : -1;                                             // gradient: column 6 of a product with [X | r]
        const int idx27 = idxU >= 0 ? idxU : idxG;
        m36[e] = idx36 >= 0; m27[e] = idx27 >= 0; mU[e] = idxU >= 0; mG[e] = idxG >= 0;
        eW[e] = aW + 8u * max(idx36, 0); eHcm[e] = aHcm + 8u * max(idx36, 0);
        eHmm[e] = aHmm + 8u * max(idx27, 0); eHf[e] = aHf + 8u * max(idx27, 0);
        eHccU[e] = aHcc + 8u * max(idxU, 0); eHccG[e] = aHcc + 8u * max(idxG, 0);
    }
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0, opt_f = p.opt_f != 0;
    for (int b = blockIdx.x; b < pl.nbatch; b += gridDim.x) {
        const int f0 = pl.batch_f[b], f1 = pl.batch_f[b + 1];
        const int o0 = pl.frame_obs_ptr[f0], o1 = pl.frame_obs_ptr[f1];
        const int sl0 = p.frame_slot_ptr[f0], sl1 = p.frame_slot_ptr[f1];
        const int per = (o1 - o0 + ACC2_WARPS - 1) / ACC2_WARPS;
        const int wa = o0 + warp * per, wb = min(o1, wa + per);
        if (wa < wb) {
            double Tff[2] = {0, 0}, Tcf[2] = {0, 0}, Tcc[2] = {0, 0};
            int cur_f = -1, cur_c = -1, cur_slc = -1;
            auto emit_cam_run = [&]() {           // W_c, gc (column 6 of the same product) and Hcc of the (frame, camera) run that just ended
                if (cur_c >= 0 && opt_c && cur_c != p.root_cam) {
                    const int cb = cur_c - (cur_c > p.root_cam ? 1 : 0);
                    unsigned ad[4]; double val[4]; int on[4];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        // element (g, 2q + e) of Jc^T [Jf | r]: a W_c entry (if the pair has a slot), or gc, or nothing
                        ad[e] = mG[e] ? eHccG[e] + 8u * 27u * cb : eW[e] + 8u * 36u * max(cur_slc - sl0, 0);
                        on[e] = mG[e] | (m36[e] & (cur_slc >= 0)); val[e] = Tcf[e];
                        ad[2 + e] = eHccU[e] + 8u * 27u * cb; on[2 + e] = mU[e]; val[2 + e] = Tcc[e];
                    }
                    smem_add_batch<4>(ad, val, on);
                }
                Tcf[0] = Tcf[1] = Tcc[0] = Tcc[1] = 0.0;
            };
            auto emit_frame_run = [&]() {         // Hff + gf of the frame run that just ended
                if (cur_f >= 0 && opt_f) {
                    unsigned ad[2]; double val[2]; int on[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) { ad[e] = eHf[e] + 8u * 27u * (cur_f - f0); on[e] = m27[e]; val[e] = Tff[e]; }
                    smem_add_batch<2>(ad, val, on);
                }
                Tff[0] = Tff[1] = 0.0;
            };
            // this lane's fragment elements of one observation: rows 2q, 2q + 1 of dof g of the three column groups; the lanes
            // of fragment row 6 carry the residual instead, those of row 7 nothing (their registers stay zero)
            auto load_frag = [&](int o, V2 &xc, V2 &xm, V2 &xf, double2 &rr) {
                if (g < 6) {
                    const JT *row = Jn + (size_t)o * 144 + g * 8 + 2 * q;
                    xc = *reinterpret_cast<const V2 *>(row); xm = *reinterpret_cast<const V2 *>(row + 48); xf = *reinterpret_cast<const V2 *>(row + 96);
                } else if (g == 6) rr = *reinterpret_cast<const double2 *>(Rv + (size_t)o * 8 + 2 * q);
            };
            // two fragment buffers used alternately (current / in flight): the loop over observations is unrolled by two so
            // that no registers are copied between them
            V2 axc, axm, axf, bxc, bxm, bxf; double2 arr, brr;
            axc.x = axc.y = axm.x = axm.y = axf.x = axf.y = bxc.x = bxc.y = bxm.x = bxm.y = bxf.x = bxf.y = (JT)0; arr.x = arr.y = brr.x = brr.y = 0.0;
            load_frag(wa, axc, axm, axf, arr);
            int cm_l = 0, f_l = 0, sl_l = 0, base = wa;
            auto one_obs = [&](int t, V2 &xc, V2 &xm, V2 &xf, double2 &rr, V2 &nxc, V2 &nxm, V2 &nxf, double2 &nrr) {
                    const int o = base + t;
                    const int cm = __shfl_sync(0xffffffffu, cm_l, t), f = __shfl_sync(0xffffffffu, f_l, t), sl = __shfl_sync(0xffffffffu, sl_l, t);
                    const int slc = (sl & 0xffff) - 1 + sl0, slm = (sl >> 16) - 1;          // slm relative to the batch, -1 = none
                    if (o + 1 < wb) load_frag(o + 1, nxc, nxm, nxf, nrr);          // next observation's fragments in flight during this one
                    const int c = obs_cam(cm), m = obs_marker(cm);
                    if (f != cur_f || c != cur_c) {
                        emit_cam_run();
                        if (f != cur_f) { emit_frame_run(); cur_f = f; }
                        cur_c = c; cur_slc = (sl & 0xffff) ? slc : -1;
                    }
                    const bool use = !obs_nojac(cm);
                    const bool uc = use && opt_c && c != p.root_cam, um = use && opt_m && m != p.root_marker, uf = use && opt_f;
                    if (!(uc && um && uf)) {        // rare: a block of this observation has no columns (root camera / marker, erased duplicate)
                        if (!uc) { xc.x = (JT)0; xc.y = (JT)0; }
                        if (!um) { xm.x = (JT)0; xm.y = (JT)0; }
                        if (!uf) { xf.x = (JT)0; xf.y = (JT)0; }
                        if (!use) { rr.x = 0.0; rr.y = 0.0; }
                    }
                    const double ac0 = (double)xc.x, ac1 = (double)xc.y, am0 = (double)xm.x, am1 = (double)xm.y, af0 = (double)xf.x, af1 = (double)xf.y;
                    const double bfr0 = g == 6 ? rr.x : af0, bfr1 = g == 6 ? rr.y : af1;        // [Jf | r]
                    if (opt_f) { dmma884(Tff, af0, bfr0); dmma884(Tff, af1, bfr1); }
                    if (opt_c) { dmma884(Tcf, ac0, bfr0); dmma884(Tcf, ac1, bfr1); dmma884(Tcc, ac0, ac0); dmma884(Tcc, ac1, ac1); }
                    if (opt_m) {
                        // marker-keyed sums: no runs in row order, every observation leaves its blocks right away
                        double Tmm[2] = {0, 0}, Tcm[2] = {0, 0}, Tmf[2] = {0, 0};
                        const double bmr0 = g == 6 ? rr.x : am0, bmr1 = g == 6 ? rr.y : am1;    // [Jm | r]
                        dmma884(Tmm, am0, bmr0); dmma884(Tmm, am1, bmr1);
                        if (opt_c) { dmma884(Tcm, ac0, am0); dmma884(Tcm, ac1, am1); }
                        if (opt_f) { dmma884(Tmf, am0, af0); dmma884(Tmf, am1, af1); }
                        if (um) {
                            const int mb = m - (m > p.root_marker ? 1 : 0), cb = c - (c > p.root_cam ? 1 : 0);
                            const int pair = uc ? cb * p.nrm + mb : -1;
                            const bool pair_sm = pair >= 0 && pair < pl.hcm_smem, wm = uf && slm >= 0;
                            unsigned ad[6]; double val[6]; int on[6];
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                ad[e] = eHmm[e] + 8u * 27u * mb; on[e] = m27[e]; val[e] = Tmm[e];
                                ad[2 + e] = eW[e] + 8u * 36u * max(slm, 0); on[2 + e] = m36[e] & (int)wm; val[2 + e] = Tmf[e];
                                ad[4 + e] = eHcm[e] + 8u * 36u * max(pair, 0); on[4 + e] = m36[e] & (int)pair_sm; val[4 + e] = Tcm[e];
                            }
                            smem_add_batch<6>(ad, val, on);
                            if (pair >= 0 && !pair_sm) {
#pragma unroll
                                for (int e = 0; e < 2; e++)
                                    if (m36[e]) atomicAdd(Hrr + (size_t)(6 * cb + g) * n_r + 6 * p.nrc + 6 * mb + 2 * q + e, Tcm[e] * s2);
                            }
                        }
                    }
            };
            for (; base < wb; base += 32) {
                const int my = base + lane;
                cm_l = (int)0x80000000u; f_l = 0; sl_l = 0;
                if (my < wb) {      // W slots relative to the batch, + 1 (0 = none), camera | marker << 16
                    cm_l = p.obs_cm[my]; f_l = p.obs_f[my];
                    const int a = p.obs_slot_c[my], c2 = p.obs_slot_m[my];
                    sl_l = (a >= 0 ? a - sl0 + 1 : 0) | ((c2 >= 0 ? c2 - sl0 + 1 : 0) << 16);
                }
                const int cnt = min(32, wb - base);           // 32 (even) in every group but the last
                for (int t = 0; t < cnt; t += 2) {
                    one_obs(t, axc, axm, axf, arr, bxc, bxm, bxf, brr);
                    if (t + 1 < cnt) one_obs(t + 1, bxc, bxm, bxf, brr, axc, axm, axf, arr);
                }
            }
            emit_cam_run(); emit_frame_run();
        }
        __syncthreads();
        // the batch's frame-keyed blocks are complete: plain coalesced stores, and the windows are cleared for the next batch
        for (int i = tid; i < (sl1 - sl0) * 36; i += ACC2_THREADS) { W[(size_t)sl0 * 36 + i] = sW[i] * s2; sW[i] = 0.0; }
        for (int i = tid; i < (f1 - f0) * 27; i += ACC2_THREADS) { Hf[(size_t)f0 * HF_STRIDE + i] = sHf[i] * ((i % 27) < 21 ? s2 : s1); sHf[i] = 0.0; }
        __syncthreads();
    }
    // ---------------- camera / marker sums of the whole CTA: one flush (as k_jac_accumulate)
    for (int i = tid; i < (p.nrc + p.nrm) * 27; i += ACC2_THREADS) {
        const int b = i / 27, e = i % 27; const double v = sHcc[i];
        if (v == 0.0) continue;
        if (e < 21) {
            int r0 = 0, rem = e; while (rem >= 6 - r0) { rem -= 6 - r0; r0++; }
            const int c0 = r0 + rem;
            atomicAdd(Hrr + (size_t)(6 * b + r0) * n_r + 6 * b + c0, v * s2);
            if (c0 != r0) atomicAdd(Hrr + (size_t)(6 * b + c0) * n_r + 6 * b + r0, v * s2);
        } else atomicAdd(gr + 6 * b + (e - 21), v * s1);
    }
    for (int i = tid; i < pl.hcm_smem * 36; i += ACC2_THREADS) {
        const double v = sHcm[i];
        if (v == 0.0) continue;
        const int blk = i / 36, e = i % 36, cbb = blk / p.nrm, mbb = blk % p.nrm;
        atomicAdd(Hrr + (size_t)(6 * cbb + e / 6) * n_r + 6 * p.nrc + 6 * mbb + e % 6, v * s2);
    }
}

// ------------------------------------------------------------------------------------------------
// K2''  k_acc_frames + k_acc_reduced: the tensor-core accumulation split by KEY so that every sum runs over observations
// that are consecutive in the order it is visited in, accumulates inside the DMMA accumulators and leaves with one
// (shared-memory or global) add per RUN instead of one per observation.  k_jac_accumulate and k_jac_accumulate_mma visit the
// observations in row order only, where the marker-keyed blocks (99 of the 189 values of an observation) have no runs: ~2.5 G
// atomic adds per Jacobian evaluation at BASELINE cfg 4, which is what both kernels spend their time on
// (profiles/r1_notes.md).  Three passes over the observation rows ([N][144] numerators, [N][8] residuals):
//   k_acc_frames, phase A, row order (frame, camera, ...):       Hff + gf  (runs = frames), W_c (runs = (frame, camera))
//   k_acc_frames, phase B, (marker, camera) order inside a frame: W_m      (runs = (frame, marker))
//   k_acc_reduced, (camera, marker, frame) order over the shard:  Hcc + gc (runs = cameras), Hmm + gm, Hcm (runs = pairs)
// The orders are permutations built once by aar_problem_create; a row is 576 contiguous bytes, so gathering rows is cheap.
// k_acc_frames keeps the frame batches and shared-memory windows of k_jac_accumulate_mma (plain stores for Hf and W).
struct Acc3Plan { const int *perm_fm; /* [N] rows of each frame sorted by (marker, camera) */ const int *perm_cm; /* [N] rows sorted by (camera, marker, frame) */ };

template <typename JT>
__global__ void __launch_bounds__(ACC2_THREADS, 1) k_acc_frames(DevProblem p, Acc2Plan pl, Acc3Plan pm, const JT *__restrict__ Jn, const double *__restrict__ Rv,
                                                               double *__restrict__ Hf, double *__restrict__ W) {
    typedef typename Vec2<JT>::type V2;
    extern __shared__ __align__(16) double sAcc[];
    double *sW = sAcc;                                                        // [win_slots][36]  W blocks of the current batch
    double *sHf = sW + (size_t)pl.win_slots * 36;                             // [win_frames][27] frame blocks of the current batch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int n_all = pl.win_slots * 36 + pl.win_frames * 27;
    for (int i = tid; i < n_all; i += ACC2_THREADS) sAcc[i] = 0.0;
    __syncthreads();
    const double s1 = pl.s1, s2 = pl.s2;
    const unsigned aW = (unsigned)__cvta_generic_to_shared(sW), aHf = (unsigned)__cvta_generic_to_shared(sHf);
    unsigned eW[2], eHf[2]; int m36[2], m27[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int j = 2 * q + e;
        const int idx36 = (g < 6 && j < 6) ? g * 6 + j : -1;
        const int idxU = (g < 6 && j < 6 && j >= g) ? g * 6 - g * (g - 1) / 2 + (j - g) : -1;
        const int idx27 = idxU >= 0 ? idxU : ((g < 6 && j == 6) ? 21 + g : -1);
        m36[e] = idx36 >= 0; m27[e] = idx27 >= 0;
        eW[e] = aW + 8u * max(idx36, 0); eHf[e] = aHf + 8u * max(idx27, 0);
    }
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0, opt_f = p.opt_f != 0;
    auto add_W = [&](int slot_rel, const double (&T)[2]) {      // a finished 6x6 W block into the window
        unsigned ad[2]; double val[2]; int on[2];
#pragma unroll
        for (int e = 0; e < 2; e++) { ad[e] = eW[e] + 8u * 36u * max(slot_rel, 0); on[e] = m36[e] & (slot_rel >= 0); val[e] = T[e]; }
        smem_add_batch<2>(ad, val, on);
    };
    for (int b = blockIdx.x; b < pl.nbatch; b += gridDim.x) {
        const int f0 = pl.batch_f[b], f1 = pl.batch_f[b + 1];
        const int o0 = pl.frame_obs_ptr[f0], o1 = pl.frame_obs_ptr[f1];
        const int sl0 = p.frame_slot_ptr[f0], sl1 = p.frame_slot_ptr[f1];
        const int per = (o1 - o0 + ACC2_WARPS - 1) / ACC2_WARPS;
        const int wa = o0 + warp * per, wb = min(o1, wa + per);
        if (wa < wb && opt_f) {
            // ---------------- phase A: row order.  Hff + gf = Jf^T [Jf | r] over the frame run, W_c = Jc^T Jf over the (frame, camera) run
            {
                double Tff[2] = {0, 0}, Tcf[2] = {0, 0};
                int cur_f = -1, cur_c = -1, cur_slc = -1;
                auto emit_cam = [&]() { if (cur_slc >= 0) add_W(cur_slc, Tcf); Tcf[0] = Tcf[1] = 0.0; };
                auto emit_frame = [&]() {
                    if (cur_f >= 0) {
                        unsigned ad[2]; double val[2]; int on[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) { ad[e] = eHf[e] + 8u * 27u * (cur_f - f0); on[e] = m27[e]; val[e] = Tff[e]; }
                        smem_add_batch<2>(ad, val, on);
                    }
                    Tff[0] = Tff[1] = 0.0;
                };
                auto load_a = [&](int o, V2 &xc, V2 &xf, double2 &rr) {
                    if (g < 6) { const JT *row = Jn + (size_t)o * 144 + g * 8 + 2 * q; xc = *reinterpret_cast<const V2 *>(row); xf = *reinterpret_cast<const V2 *>(row + 96); }
                    else if (g == 6) rr = *reinterpret_cast<const double2 *>(Rv + (size_t)o * 8 + 2 * q);
                };
                V2 axc, axf, bxc, bxf; double2 arr, brr;
                axc.x = axc.y = axf.x = axf.y = bxc.x = bxc.y = bxf.x = bxf.y = (JT)0; arr.x = arr.y = brr.x = brr.y = 0.0;
                load_a(wa, axc, axf, arr);
                int cm_l = 0, f_l = 0, sl_l = 0, base = wa;
                auto one = [&](int t, V2 &xc, V2 &xf, double2 &rr, V2 &nxc, V2 &nxf, double2 &nrr) {
                    const int o = base + t;
                    const int cm = __shfl_sync(0xffffffffu, cm_l, t), f = __shfl_sync(0xffffffffu, f_l, t), slc = __shfl_sync(0xffffffffu, sl_l, t);
                    if (o + 1 < wb) load_a(o + 1, nxc, nxf, nrr);
                    const int c = obs_cam(cm);
                    if (f != cur_f || c != cur_c) {
                        emit_cam();
                        if (f != cur_f) { emit_frame(); cur_f = f; }
                        cur_c = c; cur_slc = slc;
                    }
                    const bool use = !obs_nojac(cm), uc = use && opt_c && c != p.root_cam;
                    if (!uc) { xc.x = (JT)0; xc.y = (JT)0; if (!use) { xf.x = (JT)0; xf.y = (JT)0; rr.x = 0.0; rr.y = 0.0; } }
                    const double ac0 = (double)xc.x, ac1 = (double)xc.y, af0 = (double)xf.x, af1 = (double)xf.y;
                    const double bfr0 = g == 6 ? rr.x : af0, bfr1 = g == 6 ? rr.y : af1;        // [Jf | r]
                    dmma884(Tff, af0, bfr0); dmma884(Tff, af1, bfr1);
                    if (opt_c) { dmma884(Tcf, ac0, af0); dmma884(Tcf, ac1, af1); }
                };
                for (; base < wb; base += 32) {
                    const int my = base + lane;
                    cm_l = (int)0x80000000u; f_l = 0; sl_l = -1;
                    if (my < wb) { cm_l = p.obs_cm[my]; f_l = p.obs_f[my]; const int a = p.obs_slot_c[my]; sl_l = a >= 0 ? a - sl0 : -1; }
                    const int cnt = min(32, wb - base);
                    for (int t = 0; t < cnt; t += 2) {
                        one(t, axc, axf, arr, bxc, bxf, brr);
                        if (t + 1 < cnt) one(t + 1, bxc, bxf, brr, axc, axf, arr);
                    }
                }
                emit_cam(); emit_frame();
            }
            // ---------------- phase B: the same rows in (marker, camera) order inside each frame.  W_m = Jm^T Jf over the (frame, marker) run
            if (opt_m) {
                double Tmf[2] = {0, 0};
                int cur_slm = -1;
                auto load_b = [&](int o, V2 &xm, V2 &xf) {
                    if (g < 6) { const JT *row = Jn + (size_t)o * 144 + g * 8 + 2 * q; xm = *reinterpret_cast<const V2 *>(row + 48); xf = *reinterpret_cast<const V2 *>(row + 96); }
                };
                V2 axm, axf, bxm, bxf;
                axm.x = axm.y = axf.x = axf.y = bxm.x = bxm.y = bxf.x = bxf.y = (JT)0;
                int o_l = 0, cm_l = 0, sl_l = 0, base = wa;
                auto one = [&](int t, int cnt, V2 &xm, V2 &xf, V2 &nxm, V2 &nxf) {
                    const int cm = __shfl_sync(0xffffffffu, cm_l, t), slm = __shfl_sync(0xffffffffu, sl_l, t);
                    const int on = __shfl_sync(0xffffffffu, o_l, min(t + 1, 31));
                    if (t + 1 < cnt) load_b(on, nxm, nxf);           // the next row of this group of 32
                    if (slm != cur_slm) { if (cur_slm >= 0) add_W(cur_slm, Tmf); Tmf[0] = Tmf[1] = 0.0; cur_slm = slm; }
                    const bool um = !obs_nojac(cm) && obs_marker(cm) != p.root_marker;
                    if (!um) { xm.x = (JT)0; xm.y = (JT)0; if (obs_nojac(cm)) { xf.x = (JT)0; xf.y = (JT)0; } }      // rows of erased duplicates are never written
                    dmma884(Tmf, (double)xm.x, (double)xf.x); dmma884(Tmf, (double)xm.y, (double)xf.y);
                };
                for (; base < wb; base += 32) {
                    const int my = base + lane;
                    o_l = wa; cm_l = (int)0x80000000u; sl_l = -1;
                    if (my < wb) { o_l = pm.perm_fm[my]; cm_l = p.obs_cm[o_l]; const int a = p.obs_slot_m[o_l]; sl_l = a >= 0 ? a - sl0 : -1; }
                    const int cnt = min(32, wb - base);
                    load_b(__shfl_sync(0xffffffffu, o_l, 0), axm, axf);
                    for (int t = 0; t < cnt; t += 2) {
                        one(t, cnt, axm, axf, bxm, bxf);
                        if (t + 1 < cnt) one(t + 1, cnt, bxm, bxf, axm, axf);
                    }
                }
                if (cur_slm >= 0) add_W(cur_slm, Tmf);
            }
        }
        __syncthreads();
        // the batch's frame-keyed blocks are complete: plain coalesced stores, and the windows are cleared for the next batch
        for (int i = tid; i < (sl1 - sl0) * 36; i += ACC2_THREADS) { W[(size_t)sl0 * 36 + i] = sW[i] * s2; sW[i] = 0.0; }
        for (int i = tid; i < (f1 - f0) * 27; i += ACC2_THREADS) { Hf[(size_t)f0 * HF_STRIDE + i] = sHf[i] * ((i % 27) < 21 ? s2 : s1); sHf[i] = 0.0; }
        __syncthreads();
    }
}

// Reduced-system blocks over the rows in (camera, marker, frame) order: every warp takes one contiguous piece of the sorted
// list, the three products accumulate in DMMA accumulators over the runs and leave with global atomic adds when the camera
// or the (camera, marker) pair changes — a few per warp, against 63 shared / global atomics per observation before.
constexpr int ACC3_THREADS = 256;
template <typename JT>
__global__ void __launch_bounds__(ACC3_THREADS, 3) k_acc_reduced(DevProblem p, Acc3Plan pm, double s1, double s2, const JT *__restrict__ Jn, const double *__restrict__ Rv,
                                                                double *__restrict__ Hrr, double *__restrict__ gr) {
    typedef typename Vec2<JT>::type V2;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const long long nwarps = (long long)gridDim.x * (ACC3_THREADS / 32), wid = (long long)blockIdx.x * (ACC3_THREADS / 32) + (threadIdx.x >> 5);
    const long long per = ((p.N + nwarps - 1) / nwarps + 31) / 32 * 32;       // whole groups of 32 rows
    const long long wa = wid * per, wb = min(p.N, wa + per);
    if (wa >= wb) return;
    const int n_r = p.n_r;
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0;
    double Tcc[2] = {0, 0}, Tmm[2] = {0, 0}, Tcm[2] = {0, 0};
    int cur_c = -1, cur_m = -1;
    auto emit_pair = [&]() {          // Hmm + gm and Hcm of the (camera, marker) run that just ended
        if (cur_m >= 0 && opt_m && cur_m != p.root_marker) {
            const int mb = cur_m - (cur_m > p.root_marker ? 1 : 0), bm = 6 * (p.nrc + mb);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = 2 * q + e; const double v = Tmm[e];
                if (g < 6 && v != 0.0) {
                    if (j == 6) atomicAdd(gr + bm + g, v * s1);
                    else if (j < 6 && j >= g) {                                  // both triangles of the diagonal block, as the flush of k_jac_accumulate
                        atomicAdd(Hrr + (size_t)(bm + g) * n_r + bm + j, v * s2);
                        if (j != g) atomicAdd(Hrr + (size_t)(bm + j) * n_r + bm + g, v * s2);
                    }
                }
            }
            if (cur_c >= 0 && opt_c && cur_c != p.root_cam) {
                const int cb = cur_c - (cur_c > p.root_cam ? 1 : 0);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = 2 * q + e; const double v = Tcm[e];
                    if (g < 6 && j < 6 && v != 0.0) atomicAdd(Hrr + (size_t)(6 * cb + g) * n_r + 6 * p.nrc + 6 * mb + j, v * s2);
                }
            }
        }
        Tmm[0] = Tmm[1] = Tcm[0] = Tcm[1] = 0.0;
    };
    auto emit_cam = [&]() {           // Hcc + gc of the camera run that just ended
        if (cur_c >= 0 && opt_c && cur_c != p.root_cam) {
            const int bc = 6 * (cur_c - (cur_c > p.root_cam ? 1 : 0));
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = 2 * q + e; const double v = Tcc[e];
                if (g < 6 && v != 0.0) {
                    if (j == 6) atomicAdd(gr + bc + g, v * s1);
                    else if (j < 6 && j >= g) {
                        atomicAdd(Hrr + (size_t)(bc + g) * n_r + bc + j, v * s2);
                        if (j != g) atomicAdd(Hrr + (size_t)(bc + j) * n_r + bc + g, v * s2);
                    }
                }
            }
        }
        Tcc[0] = Tcc[1] = 0.0;
    };
    auto load_c = [&](int o, V2 &xc, V2 &xm, double2 &rr) {
        if (g < 6) { const JT *row = Jn + (size_t)o * 144 + g * 8 + 2 * q; xc = *reinterpret_cast<const V2 *>(row); xm = *reinterpret_cast<const V2 *>(row + 48); }
        else if (g == 6) rr = *reinterpret_cast<const double2 *>(Rv + (size_t)o * 8 + 2 * q);
    };
    auto prefetch_rows = [&](int o) {    // camera + marker columns (384 contiguous bytes) and the residual of one row towards L2
        const char *row = reinterpret_cast<const char *>(Jn + (size_t)o * 144);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row)); asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128 * sizeof(JT) / 4));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 256 * sizeof(JT) / 4)); asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 96 * sizeof(JT) - 1));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Rv + (size_t)o * 8));
    };
    V2 axc, axm, bxc, bxm; double2 arr, brr;
    axc.x = axc.y = axm.x = axm.y = bxc.x = bxc.y = bxm.x = bxm.y = (JT)0; arr.x = arr.y = brr.x = brr.y = 0.0;
    int o_l = 0, cm_l = 0, on_l = 0;
    long long base = wa;
    {   // rows of the first group towards L2, indices of the first group
        const long long my = base + lane;
        on_l = my < wb ? pm.perm_cm[my] : 0;
        if (my < wb) prefetch_rows(on_l);
    }
    auto one = [&](int t, int cnt, V2 &xc, V2 &xm, double2 &rr, V2 &nxc, V2 &nxm, double2 &nrr) {
        const int cm = __shfl_sync(0xffffffffu, cm_l, t);
        const int on = __shfl_sync(0xffffffffu, o_l, min(t + 1, 31));
        if (t + 1 < cnt) load_c(on, nxc, nxm, nrr);
        const int c = obs_cam(cm), m = obs_marker(cm);
        if (c != cur_c || m != cur_m) {
            emit_pair();
            if (c != cur_c) { emit_cam(); cur_c = c; }
            cur_m = m;
        }
        const bool use = !obs_nojac(cm), uc = use && opt_c && c != p.root_cam, um = use && opt_m && m != p.root_marker;
        if (!(uc && um)) {
            if (!uc) { xc.x = (JT)0; xc.y = (JT)0; }
            if (!um) { xm.x = (JT)0; xm.y = (JT)0; }
            if (!use) { rr.x = 0.0; rr.y = 0.0; }
        }
        const double ac0 = (double)xc.x, ac1 = (double)xc.y, am0 = (double)xm.x, am1 = (double)xm.y;
        const double bcr0 = g == 6 ? rr.x : ac0, bcr1 = g == 6 ? rr.y : ac1, bmr0 = g == 6 ? rr.x : am0, bmr1 = g == 6 ? rr.y : am1;   // [Jc | r], [Jm | r]
        if (opt_c) { dmma884(Tcc, ac0, bcr0); dmma884(Tcc, ac1, bcr1); }
        if (opt_m) { dmma884(Tmm, am0, bmr0); dmma884(Tmm, am1, bmr1); }
        if (opt_c && opt_m) { dmma884(Tcm, ac0, am0); dmma884(Tcm, ac1, am1); }
    };
    for (; base < wb; base += 32) {
        o_l = on_l; cm_l = (int)0x80000000u;
        if (base + lane < wb) cm_l = p.obs_cm[o_l];
        {   // indices of the NEXT group, and its rows towards L2 while this group is multiplied
            const long long my = base + 32 + lane;
            on_l = my < wb ? pm.perm_cm[my] : 0;
            if (my < wb) prefetch_rows(on_l);
        }
        const int cnt = (int)min((long long)32, wb - base);
        load_c(__shfl_sync(0xffffffffu, o_l, 0), axc, axm, arr);
        for (int t = 0; t < cnt; t += 2) {
            one(t, cnt, axc, axm, arr, bxc, bxm, brr);
            if (t + 1 < cnt) one(t + 1, cnt, bxc, bxm, brr, axc, axm, arr);
        }
    }
    emit_pair(); emit_cam();
}

// k_acc_reduced with the rows of a warp staged through shared memory: k_acc_reduced keeps ONE row in flight per warp and a
// warp cannot start a row before its data has arrived (600-800 cycles from L2), although the work on a row is ~200 cycles
// (profiles/r1_notes.md).  Here a warp copies the camera / marker columns and the residual of the next ACC3_ST rows of its list
// with cp.async (two lanes per row, 16 bytes per copy) while it multiplies the previous ACC3_ST from the other buffer.
constexpr int ACC3_ST = 16;                 // rows per stage
#ifndef AAR_DMMA_TWO_SETS
// 1: the two k-steps of a row go to two accumulator sets that are added when a run ends, so that no DMMA waits for the one
// issued just before it (SASS of the kernel below: a row is two DEPENDENT DMMAs per product followed by register moves that
// wait for them).  Written at the very end of round 1 without GPU time left to measure it: off; in k_acc_reduced_staged and k_acc_frames_staged.
#define AAR_DMMA_TWO_SETS 0
#endif
template <typename JT> struct Acc3Row { static constexpr int BYTES = 96 * (int)sizeof(JT) + 64; };   // [Jc | Jm] + r
template <typename JT> constexpr size_t acc3_smem_bytes() { return (size_t)(ACC3_THREADS / 32) * 2 * ACC3_ST * Acc3Row<JT>::BYTES; }

template <typename JT>
__global__ void __launch_bounds__(ACC3_THREADS, 2) k_acc_reduced_staged(DevProblem p, Acc3Plan pm, double s1, double s2, const JT *__restrict__ Jn, const double *__restrict__ Rv,
                                                                       double *__restrict__ Hrr, double *__restrict__ gr) {
    typedef typename Vec2<JT>::type V2;
    extern __shared__ __align__(16) unsigned char sStage[];
    constexpr int ROWB = Acc3Row<JT>::BYTES, NCH = ROWB / 16, JCH = 96 * (int)sizeof(JT) / 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    unsigned char *wbuf = sStage + (size_t)warp * 2 * ACC3_ST * ROWB;
    const long long nwarps = (long long)gridDim.x * (ACC3_THREADS / 32), wid = (long long)blockIdx.x * (ACC3_THREADS / 32) + warp;
    const long long per = ((p.N + nwarps - 1) / nwarps + ACC3_ST - 1) / ACC3_ST * ACC3_ST;       // whole stages
    const long long wa = wid * per, wb = min(p.N, wa + per);
    if (wa >= wb) return;
    const int n_r = p.n_r;
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0;
    double Tcc[2] = {0, 0}, Tmm[2] = {0, 0}, Tcm[2] = {0, 0};
#if AAR_DMMA_TWO_SETS
    double Ucc[2] = {0, 0}, Umm[2] = {0, 0}, Ucm[2] = {0, 0};       // second k-step of every row: no DMMA depends on the one before it
#endif
    int cur_c = -1, cur_m = -1;
    auto emit_pair = [&]() {          // Hmm + gm and Hcm of the (camera, marker) run that just ended
#if AAR_DMMA_TWO_SETS
        Tmm[0] += Umm[0]; Tmm[1] += Umm[1]; Tcm[0] += Ucm[0]; Tcm[1] += Ucm[1]; Umm[0] = Umm[1] = Ucm[0] = Ucm[1] = 0.0;
#endif
        if (cur_m >= 0 && opt_m && cur_m != p.root_marker) {
            const int mb = cur_m - (cur_m > p.root_marker ? 1 : 0), bm = 6 * (p.nrc + mb);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = 2 * q + e; const double v = Tmm[e];
                if (g < 6 && v != 0.0) {
                    if (j == 6) atomicAdd(gr + bm + g, v * s1);
                    else if (j < 6 && j >= g) {
                        atomicAdd(Hrr + (size_t)(bm + g) * n_r + bm + j, v * s2);
                        if (j != g) atomicAdd(Hrr + (size_t)(bm + j) * n_r + bm + g, v * s2);
                    }
                }
            }
            if (cur_c >= 0 && opt_c && cur_c != p.root_cam) {
                const int cb = cur_c - (cur_c > p.root_cam ? 1 : 0);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = 2 * q + e; const double v = Tcm[e];
                    if (g < 6 && j < 6 && v != 0.0) atomicAdd(Hrr + (size_t)(6 * cb + g) * n_r + 6 * p.nrc + 6 * mb + j, v * s2);
                }
            }
        }
        Tmm[0] = Tmm[1] = Tcm[0] = Tcm[1] = 0.0;
    };
    auto emit_cam = [&]() {           // Hcc + gc of the camera run that just ended
#if AAR_DMMA_TWO_SETS
        Tcc[0] += Ucc[0]; Tcc[1] += Ucc[1]; Ucc[0] = Ucc[1] = 0.0;
#endif
        if (cur_c >= 0 && opt_c && cur_c != p.root_cam) {
            const int bc = 6 * (cur_c - (cur_c > p.root_cam ? 1 : 0));
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = 2 * q + e; const double v = Tcc[e];
                if (g < 6 && v != 0.0) {
                    if (j == 6) atomicAdd(gr + bc + g, v * s1);
                    else if (j < 6 && j >= g) {
                        atomicAdd(Hrr + (size_t)(bc + g) * n_r + bc + j, v * s2);
                        if (j != g) atomicAdd(Hrr + (size_t)(bc + j) * n_r + bc + g, v * s2);
                    }
                }
            }
        }
        Tcc[0] = Tcc[1] = 0.0;
    };
    // two lanes per row of a stage: lanes 2r and 2r + 1 hold the index of row r and copy one half of its chunks each
    const int myrow = lane >> 1, half = lane & 1;
    auto row_index = [&](long long pos) -> int { const long long my = pos + myrow; return my < wb ? pm.perm_cm[my] : -1; };
    auto issue = [&](int b, int idx) {
        if (idx >= 0) {
            const char *srcJ = reinterpret_cast<const char *>(Jn + (size_t)idx * 144), *srcR = reinterpret_cast<const char *>(Rv + (size_t)idx * 8);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(wbuf + ((size_t)b * ACC3_ST + myrow) * ROWB);
#pragma unroll
            for (int k = 0; k < NCH / 2; k++) {
                const int c = half * (NCH / 2) + k;
                const char *src = c < JCH ? srcJ + 16 * c : srcR + 16 * (c - JCH);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int idx_cur = row_index(wa);
    int cm_cur = idx_cur >= 0 ? p.obs_cm[idx_cur] : (int)0x80000000u;
    issue(0, idx_cur);
    int idx_nxt = row_index(wa + ACC3_ST);
    int stage = 0;
    for (long long pos = wa; pos < wb; pos += ACC3_ST, stage ^= 1) {
        const int cm_nxt = idx_nxt >= 0 ? p.obs_cm[idx_nxt] : (int)0x80000000u;
        issue(stage ^ 1, idx_nxt);                                     // the next stage's rows in flight while this stage is multiplied
        const int idx_nn = row_index(pos + 2 * ACC3_ST);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const int cnt = (int)min((long long)ACC3_ST, wb - pos);
        const unsigned char *sb = wbuf + (size_t)stage * ACC3_ST * ROWB;
        for (int t = 0; t < cnt; t++) {
            const int cm = __shfl_sync(0xffffffffu, cm_cur, 2 * t);
            const unsigned char *row = sb + (size_t)t * ROWB;
            V2 xc, xm; double2 rr;
            xc.x = xc.y = xm.x = xm.y = (JT)0; rr.x = rr.y = 0.0;
            if (g < 6) { xc = *reinterpret_cast<const V2 *>(row + (g * 8 + 2 * q) * sizeof(JT)); xm = *reinterpret_cast<const V2 *>(row + (48 + g * 8 + 2 * q) * sizeof(JT)); }
            else if (g == 6) rr = *reinterpret_cast<const double2 *>(row + 96 * sizeof(JT) + 16 * q);
            const int c = obs_cam(cm), m = obs_marker(cm);
            if (c != cur_c || m != cur_m) {
                emit_pair();
                if (c != cur_c) { emit_cam(); cur_c = c; }
                cur_m = m;
            }
            const bool use = !obs_nojac(cm), uc = use && opt_c && c != p.root_cam, um = use && opt_m && m != p.root_marker;
            if (!(uc && um)) {                       // stale columns: the projection kernel does not write what has no Jacobian
                if (!uc) { xc.x = (JT)0; xc.y = (JT)0; }
                if (!um) { xm.x = (JT)0; xm.y = (JT)0; }
                if (!use) { rr.x = 0.0; rr.y = 0.0; }
            }
            const double ac0 = (double)xc.x, ac1 = (double)xc.y, am0 = (double)xm.x, am1 = (double)xm.y;
            const double bcr0 = g == 6 ? rr.x : ac0, bcr1 = g == 6 ? rr.y : ac1, bmr0 = g == 6 ? rr.x : am0, bmr1 = g == 6 ? rr.y : am1;   // [Jc | r], [Jm | r]
#if AAR_DMMA_TWO_SETS
            if (opt_c) { dmma884(Tcc, ac0, bcr0); dmma884(Ucc, ac1, bcr1); }
            if (opt_m) { dmma884(Tmm, am0, bmr0); dmma884(Umm, am1, bmr1); }
            if (opt_c && opt_m) { dmma884(Tcm, ac0, am0); dmma884(Ucm, ac1, am1); }
#else
            if (opt_c) { dmma884(Tcc, ac0, bcr0); dmma884(Tcc, ac1, bcr1); }
            if (opt_m) { dmma884(Tmm, am0, bmr0); dmma884(Tmm, am1, bmr1); }
            if (opt_c && opt_m) { dmma884(Tcm, ac0, am0); dmma884(Tcm, ac1, am1); }
#endif
        }
        __syncwarp();                                                  // everyone is done with this buffer before it is refilled
        cm_cur = cm_nxt; idx_nxt = idx_nn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    emit_pair(); emit_cam();
}

// k_acc_frames with the same staging (float32 numerators only; the FP64-staging fallback keeps k_acc_frames): per warp two
// buffers of ACCF_ST rows — phase A copies [Jc | Jf | r] of consecutive rows, phase B [Jm | Jf] of the rows its permutation names.
constexpr int ACCF_ST = 8, ACCF_ROWB = 448;       // rows per stage (4 lanes copy one row), bytes per staged row
constexpr size_t accf_stage_bytes() { return (size_t)ACC2_WARPS * 2 * ACCF_ST * ACCF_ROWB; }

__global__ void __launch_bounds__(ACC2_THREADS, 1) k_acc_frames_staged(DevProblem p, Acc2Plan pl, Acc3Plan pm, const float *__restrict__ Jn, const double *__restrict__ Rv,
                                                                      double *__restrict__ Hf, double *__restrict__ W) {
    extern __shared__ __align__(16) double sAcc[];
    double *sW = sAcc;                                                        // [win_slots][36]  W blocks of the current batch
    double *sHf = sW + (size_t)pl.win_slots * 36;                             // [win_frames][27] frame blocks of the current batch
    const int n_all = pl.win_slots * 36 + pl.win_frames * 27;
    unsigned char *sStage = reinterpret_cast<unsigned char *>(sAcc + ((n_all + 1) & ~1));          // 16-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    unsigned char *wbuf = sStage + (size_t)warp * 2 * ACCF_ST * ACCF_ROWB;
    for (int i = tid; i < n_all; i += ACC2_THREADS) sAcc[i] = 0.0;
    __syncthreads();
    const double s1 = pl.s1, s2 = pl.s2;
    const unsigned aW = (unsigned)__cvta_generic_to_shared(sW), aHf = (unsigned)__cvta_generic_to_shared(sHf);
    unsigned eW[2], eHf[2]; int m36[2], m27[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int j = 2 * q + e;
        const int idx36 = (g < 6 && j < 6) ? g * 6 + j : -1;
        const int idxU = (g < 6 && j < 6 && j >= g) ? g * 6 - g * (g - 1) / 2 + (j - g) : -1;
        const int idx27 = idxU >= 0 ? idxU : ((g < 6 && j == 6) ? 21 + g : -1);
        m36[e] = idx36 >= 0; m27[e] = idx27 >= 0;
        eW[e] = aW + 8u * max(idx36, 0); eHf[e] = aHf + 8u * max(idx27, 0);
    }
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0, opt_f = p.opt_f != 0;
    auto add_W = [&](int slot_rel, const double (&T)[2]) {      // a finished 6x6 W block into the window
        unsigned ad[2]; double val[2]; int on[2];
#pragma unroll
        for (int e = 0; e < 2; e++) { ad[e] = eW[e] + 8u * 36u * max(slot_rel, 0); on[e] = m36[e] & (slot_rel >= 0); val[e] = T[e]; }
        smem_add_batch<2>(ad, val, on);
    };
    const int myrow = lane >> 2, part = lane & 3;            // four lanes per staged row
    for (int b = blockIdx.x; b < pl.nbatch; b += gridDim.x) {
        const int f0 = pl.batch_f[b], f1 = pl.batch_f[b + 1];
        const int o0 = pl.frame_obs_ptr[f0], o1 = pl.frame_obs_ptr[f1];
        const int sl0 = p.frame_slot_ptr[f0], sl1 = p.frame_slot_ptr[f1];
        const int per = ((o1 - o0 + ACC2_WARPS - 1) / ACC2_WARPS + ACCF_ST - 1) / ACCF_ST * ACCF_ST;      // whole stages
        const int wa = o0 + warp * per, wb = min(o1, wa + per);
        if (wa < wb && opt_f) {
            // ---------------- phase A: row order.  Hff + gf = Jf^T [Jf | r] over the frame run, W_c = Jc^T Jf over the (frame, camera) run
            {
                double Tff[2] = {0, 0}, Tcf[2] = {0, 0};
#if AAR_DMMA_TWO_SETS
                double Uff[2] = {0, 0}, Ucf[2] = {0, 0};
#endif
                int cur_f = -1, cur_c = -1, cur_slc = -1;
                auto emit_cam = [&]() {
#if AAR_DMMA_TWO_SETS
                    Tcf[0] += Ucf[0]; Tcf[1] += Ucf[1]; Ucf[0] = Ucf[1] = 0.0;
#endif
                    if (cur_slc >= 0) add_W(cur_slc, Tcf);
                    Tcf[0] = Tcf[1] = 0.0;
                };
                auto emit_frame = [&]() {
#if AAR_DMMA_TWO_SETS
                    Tff[0] += Uff[0]; Tff[1] += Uff[1]; Uff[0] = Uff[1] = 0.0;
#endif
                    if (cur_f >= 0) {
                        unsigned ad[2]; double val[2]; int on[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) { ad[e] = eHf[e] + 8u * 27u * (cur_f - f0); on[e] = m27[e]; val[e] = Tff[e]; }
                        smem_add_batch<2>(ad, val, on);
                    }
                    Tff[0] = Tff[1] = 0.0;
                };
                auto issue = [&](int bsel, int pos) {          // [Jc | Jf | r] of rows pos .. pos + ACCF_ST - 1: 28 chunks of 16 bytes per row, 7 per lane
                    const int o = pos + myrow;
                    if (o < wb) {
                        const char *srcJ = reinterpret_cast<const char *>(Jn + (size_t)o * 144), *srcR = reinterpret_cast<const char *>(Rv + (size_t)o * 8);
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(wbuf + ((size_t)bsel * ACCF_ST + myrow) * ACCF_ROWB);
#pragma unroll
                        for (int k = 0; k < 7; k++) {
                            const int c = part * 7 + k;
                            const char *src = c < 12 ? srcJ + 16 * c : (c < 24 ? srcJ + 384 + 16 * (c - 12) : srcR + 16 * (c - 24));
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(src) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                auto meta = [&](int pos, int &cm, int &f, int &slc) {          // held by the four lanes of row pos + myrow
                    const int o = pos + myrow;
                    cm = (int)0x80000000u; f = 0; slc = -1;
                    if (o < wb) { cm = p.obs_cm[o]; f = p.obs_f[o]; const int a = p.obs_slot_c[o]; slc = a >= 0 ? a - sl0 : -1; }
                };
                int cm_cur, f_cur, sl_cur, cm_nxt, f_nxt, sl_nxt;
                meta(wa, cm_cur, f_cur, sl_cur);
                issue(0, wa);
                int stage = 0;
                for (int pos = wa; pos < wb; pos += ACCF_ST, stage ^= 1) {
                    meta(pos + ACCF_ST, cm_nxt, f_nxt, sl_nxt);
                    issue(stage ^ 1, pos + ACCF_ST);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                    __syncwarp();
                    const int cnt = min(ACCF_ST, wb - pos);
                    const unsigned char *sb = wbuf + (size_t)stage * ACCF_ST * ACCF_ROWB;
                    for (int t = 0; t < cnt; t++) {
                        const int cm = __shfl_sync(0xffffffffu, cm_cur, 4 * t), f = __shfl_sync(0xffffffffu, f_cur, 4 * t), slc = __shfl_sync(0xffffffffu, sl_cur, 4 * t);
                        const unsigned char *row = sb + (size_t)t * ACCF_ROWB;
                        float2 xc = make_float2(0.f, 0.f), xf = make_float2(0.f, 0.f); double2 rr = make_double2(0.0, 0.0);
                        if (g < 6) { xc = *reinterpret_cast<const float2 *>(row + (g * 8 + 2 * q) * 4); xf = *reinterpret_cast<const float2 *>(row + 192 + (g * 8 + 2 * q) * 4); }
                        else if (g == 6) rr = *reinterpret_cast<const double2 *>(row + 384 + 16 * q);
                        const int c = obs_cam(cm);
                        if (f != cur_f || c != cur_c) {
                            emit_cam();
                            if (f != cur_f) { emit_frame(); cur_f = f; }
                            cur_c = c; cur_slc = slc;
                        }
                        const bool use = !obs_nojac(cm), uc = use && opt_c && c != p.root_cam;
                        if (!uc) { xc.x = 0.f; xc.y = 0.f; if (!use) { xf.x = 0.f; xf.y = 0.f; rr.x = 0.0; rr.y = 0.0; } }
                        const double ac0 = (double)xc.x, ac1 = (double)xc.y, af0 = (double)xf.x, af1 = (double)xf.y;
                        const double bfr0 = g == 6 ? rr.x : af0, bfr1 = g == 6 ? rr.y : af1;        // [Jf | r]
#if AAR_DMMA_TWO_SETS
                        dmma884(Tff, af0, bfr0); dmma884(Uff, af1, bfr1);
                        if (opt_c) { dmma884(Tcf, ac0, af0); dmma884(Ucf, ac1, af1); }
#else
                        dmma884(Tff, af0, bfr0); dmma884(Tff, af1, bfr1);
                        if (opt_c) { dmma884(Tcf, ac0, af0); dmma884(Tcf, ac1, af1); }
#endif
                    }
                    __syncwarp();
                    cm_cur = cm_nxt; f_cur = f_nxt; sl_cur = sl_nxt;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                emit_cam(); emit_frame();
            }
            // ---------------- phase B: the same rows in (marker, camera) order inside each frame.  W_m = Jm^T Jf over the (frame, marker) run
            if (opt_m) {
                double Tmf[2] = {0, 0};
#if AAR_DMMA_TWO_SETS
                double Umf[2] = {0, 0};
#endif
                int cur_slm = -1;
                auto meta = [&](int pos, int &o, int &cm, int &slm) {
                    const int i = pos + myrow;
                    o = -1; cm = (int)0x80000000u; slm = -1;
                    if (i < wb) { o = pm.perm_fm[i]; cm = p.obs_cm[o]; const int a = p.obs_slot_m[o]; slm = a >= 0 ? a - sl0 : -1; }
                };
                auto issue = [&](int bsel, int o) {            // [Jm | Jf] = bytes 192 .. 575 of the row: 24 chunks, 6 per lane
                    if (o >= 0) {
                        const char *srcJ = reinterpret_cast<const char *>(Jn + (size_t)o * 144) + 192;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(wbuf + ((size_t)bsel * ACCF_ST + myrow) * ACCF_ROWB);
#pragma unroll
                        for (int k = 0; k < 6; k++) {
                            const int c = part * 6 + k;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(srcJ + 16 * c) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                __syncwarp();                                  // phase A's last reads of the buffers are over
                int o_cur, cm_cur, sl_cur, o_nxt, cm_nxt, sl_nxt;
                meta(wa, o_cur, cm_cur, sl_cur);
                issue(0, o_cur);
                int stage = 0;
                for (int pos = wa; pos < wb; pos += ACCF_ST, stage ^= 1) {
                    meta(pos + ACCF_ST, o_nxt, cm_nxt, sl_nxt);
                    issue(stage ^ 1, o_nxt);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                    __syncwarp();
                    const int cnt = min(ACCF_ST, wb - pos);
                    const unsigned char *sb = wbuf + (size_t)stage * ACCF_ST * ACCF_ROWB;
                    for (int t = 0; t < cnt; t++) {
                        const int cm = __shfl_sync(0xffffffffu, cm_cur, 4 * t), slm = __shfl_sync(0xffffffffu, sl_cur, 4 * t);
                        const unsigned char *row = sb + (size_t)t * ACCF_ROWB;
                        float2 xm = make_float2(0.f, 0.f), xf = make_float2(0.f, 0.f);
                        if (g < 6) { xm = *reinterpret_cast<const float2 *>(row + (g * 8 + 2 * q) * 4); xf = *reinterpret_cast<const float2 *>(row + 192 + (g * 8 + 2 * q) * 4); }
                        if (slm != cur_slm) {
#if AAR_DMMA_TWO_SETS
                            Tmf[0] += Umf[0]; Tmf[1] += Umf[1]; Umf[0] = Umf[1] = 0.0;
#endif
                            if (cur_slm >= 0) add_W(cur_slm, Tmf);
                            Tmf[0] = Tmf[1] = 0.0; cur_slm = slm;
                        }
                        const bool um = !obs_nojac(cm) && obs_marker(cm) != p.root_marker;
                        if (!um) { xm.x = 0.f; xm.y = 0.f; if (obs_nojac(cm)) { xf.x = 0.f; xf.y = 0.f; } }      // rows of erased duplicates are never written
#if AAR_DMMA_TWO_SETS
                        dmma884(Tmf, (double)xm.x, (double)xf.x); dmma884(Umf, (double)xm.y, (double)xf.y);
#else
                        dmma884(Tmf, (double)xm.x, (double)xf.x); dmma884(Tmf, (double)xm.y, (double)xf.y);
#endif
                    }
                    __syncwarp();
                    o_cur = o_nxt; cm_cur = cm_nxt; sl_cur = sl_nxt;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#if AAR_DMMA_TWO_SETS
                Tmf[0] += Umf[0]; Tmf[1] += Umf[1];
#endif
                if (cur_slm >= 0) add_W(cur_slm, Tmf);
            }
        }
        __syncthreads();
        // the batch's frame-keyed blocks are complete: plain coalesced stores, and the windows are cleared for the next batch
        for (int i = tid; i < (sl1 - sl0) * 36; i += ACC2_THREADS) { W[(size_t)sl0 * 36 + i] = sW[i] * s2; sW[i] = 0.0; }
        for (int i = tid; i < (f1 - f0) * 27; i += ACC2_THREADS) { Hf[(size_t)f0 * HF_STRIDE + i] = sHf[i] * ((i % 27) < 21 ? s2 : s1); sHf[i] = 0.0; }
        __syncthreads();
    }
}

} // namespace aar
