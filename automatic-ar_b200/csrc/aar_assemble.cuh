// aar_assemble.cuh — J^T J blocks and J^T r of SparseLevMarq::step (/root/reference/libs/sparselevmarq.h:353-367: J^T J by
// `mult` :264-325, B = -J^T x :367) from the staged central-difference numerators of k_jac_project, on the FP64 tensor cores
// (mma.sync.m8n8k4.f64).  Replaces the lane-per-observation kernel of round 1 (15.4 ms at BASELINE cfg 4, 190 warp-instructions
// per observation, most of them moving 189 sums per observation through shared-memory transpositions and atomics).
//
// Formulation.  A marker observation owns three 8 x 6 column groups of J: camera (c), marker (m), frame (f), and the residual r
// (8).  The six block products it contributes to — ff, cf, cc, cm, mm, mf — are sums over observations that share a KEY:
//      Hff + gf   frame            W_c = Jc^T Jf   (frame, camera)      Hcc + gc   camera
//      Hmm + gm   marker           W_m = Jm^T Jf   (frame, marker)      Hcm        (camera, marker)
// A block product X^T Y over the 8 residual rows is two k-steps of the 8x8x4 FP64 mma with the SAME register serving as A
// fragment of X and B fragment of X (element (g = lane / 4, q = lane % 4, step s) = X[row 2q + s][dof g]); fragment column 6 of a
// B operand carries the residual, so J^T r is column 6 of the same product.  A sum over a RUN of observations is then nothing
// but more k-steps on the same accumulator: the reduction costs no instruction at all, and the result fragment (row g, columns
// 2q, 2q + 1) leaves with two stores per lane.  So every sum is visited in an order in which its key has runs, by the warp that
// owns the run:
//   k_asm_pairs   row order; run = the observations of one (frame, camera) pair (consecutive rows).  cf -> W_c, plain store (the
//                 pair owns its W slot); ff -> Hff / gf, one RED per value and pair; cc / gc -> CTA-lifetime shared accumulators,
//                 flushed once; cm has no run anywhere inside a frame -> one 6x6 block of REDs per observation.
//   k_asm_mruns   the rows of each frame in (marker, camera) order (a permutation built once by aar_problem_create); run = the
//                 observations of one (frame, marker).  mf -> W_m, plain store; mm / gm -> shared accumulators.
// k_jac_project writes one row of JROW = 160 elements per observation: [Jc (48) | Jm (48) | Jf (48) | e (8) | marker index | pad],
// zero where a block has no columns (root camera / root marker / a group that is not optimised / an erased duplicate), so the
// inner loops carry no per-observation conditions.  e is the residual before the Huber weight: exactly a float (float - float,
// mcm.cpp:1012-1013); with Huber the four weights of the observation come from a separate array.
// No shared-memory windows, no frame batches, no CTA barriers inside the loops; 12 DMMA + ~10 loads / conversions per observation.
#pragma once

namespace aar {

struct AsmPlan {
    const int4 *pair_info;      // [npairs]  first row, rows, local frame, camera
    const int4 *mrun_info;      // [nmruns]  first entry of perm_fm, entries, W slot (-1: none), reduced marker block
    const int *perm_fm;         // [N]       rows of each frame by (marker, camera); root-marker rows are not listed
    int npairs, nmruns;
    int smem_acc;               // 1: Hcc / Hmm accumulators in shared memory; 0: straight to global (rigs too large for 227 KB)
    double s1, s2;              // 1 / (2 delta), 1 / (2 delta)^2: the numerators are divided here, once per sum
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <typename JT> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

constexpr int ASM_THREADS = 256, ASM_WARPS = ASM_THREADS / 32;
constexpr int ACC_LD = 28;      // 27 values of a packed symmetric block + gradient, padded to an even stride
#ifndef AAR_ASM_MINBLOCKS
#define AAR_ASM_MINBLOCKS 3
#endif

// index of fragment element (row g, column j) in a packed block [upper triangle by rows (21) | gradient (6)], or -1
__device__ __forceinline__ int packed27(int g, int j) {
    if (g >= 6) return -1;
    if (j == 6) return 21 + g;
    if (j < 6 && j >= g) return g * 6 - g * (g - 1) / 2 + (j - g);
    return -1;
}

// flush of a CTA's camera or marker accumulators ([nblk][ACC_LD]) into the reduced system: both triangles of the diagonal block
__device__ __forceinline__ void flush_diag_blocks(const double *__restrict__ acc, int nblk, int block0, int n_r, double s1, double s2,
                                                  double *__restrict__ Hrr, double *__restrict__ gr) {
    for (int i = threadIdx.x; i < nblk * 27; i += blockDim.x) {
        const int b = i / 27, e = i - 27 * b; const double v = acc[b * ACC_LD + e];
        if (v == 0.0) continue;
        const int d0 = 6 * (block0 + b);
        if (e < 21) {
            int r0 = 0, rem = e; while (rem >= 6 - r0) { rem -= 6 - r0; r0++; }
            const int c0 = r0 + rem;
            atomicAdd(Hrr + (size_t)(d0 + r0) * n_r + d0 + c0, v * s2);
            if (c0 != r0) atomicAdd(Hrr + (size_t)(d0 + c0) * n_r + d0 + r0, v * s2);
        } else atomicAdd(gr + d0 + (e - 21), v * s1);
    }
}

// One lane's share of one staged row (JROW elements per observation, aar_jacobian.cuh): element pair (2q, 2q + 1) of fragment
// row g of each group, three 8-byte (float staging) loads through three per-lane pointers:
//   frame group   elements 96 + 8 g + 2q of [Jf (48) | e (8) | marker index, pad (8)] for every g: row 6 is the residual
//   marker group  48 + 8 g + 2q for g < 6; the residual (144 + 2q) for g == 6, so that the fragment is [Jm | r] as it stands
//   camera group  8 g + 2q for g < 6
// Whatever the lanes of fragment rows 6 / 7 hold beyond that only reaches rows / columns 6 and 7 of a product (element (i, j) of
// A^T B is the dot product of fragment row i of A with fragment row j of B), which are never read, except column 6 of a product
// with [X | r] — so nothing is masked.
template <typename JT> struct RowFrag { typename Vec2<JT>::type c, m, f; };
template <typename JT> struct LanePtrs {
    const JT *c, *m, *f;
    __device__ __forceinline__ LanePtrs(const JT *Jn, int g, int q) {
        c = Jn + (g < 6 ? 8 * g + 2 * q : 144 + 2 * q);
        m = Jn + (g < 6 ? 48 + 8 * g + 2 * q : 144 + 2 * q);
        f = Jn + 96 + 8 * g + 2 * q;
    }
};
template <typename JT, bool WITH_C, bool WITH_M, bool WITH_F>
__device__ __forceinline__ void load_frag(const LanePtrs<JT> &lp, int o, RowFrag<JT> &x) {
    typedef typename Vec2<JT>::type V2;
    const size_t off = (size_t)o * JROW;
    if (WITH_C) x.c = *reinterpret_cast<const V2 *>(lp.c + off);
    if (WITH_M) x.m = *reinterpret_cast<const V2 *>(lp.m + off);
    if (WITH_F) x.f = *reinterpret_cast<const V2 *>(lp.f + off);
}
// the marker index of the row, stored as an integer in element 152 (fragment row 7, q = 0): broadcast from lane 28
template <typename JT> __device__ __forceinline__ int frag_marker(const RowFrag<JT> &x);
template <> __device__ __forceinline__ int frag_marker<float>(const RowFrag<float> &x) { return __shfl_sync(0xffffffffu, __float_as_int(x.f.x), 28); }
template <> __device__ __forceinline__ int frag_marker<double>(const RowFrag<double> &x) { return __shfl_sync(0xffffffffu, __double2loint(x.f.x), 28); }

// ------------------------------------------------------------------------------------------------
// Row order: one warp per (frame, camera) pair.
template <typename JT>
__global__ void __launch_bounds__(ASM_THREADS, AAR_ASM_MINBLOCKS) k_asm_pairs(DevProblem p, AsmPlan pl, const JT *__restrict__ Jn /* [N][JROW] */, const double *__restrict__ Hw /* [N][4] Huber weights or null */,
                                                                             double *__restrict__ Hf, double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    extern __shared__ __align__(16) double sAcc[];                 // [nrc][ACC_LD]
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const bool opt_c = p.opt_c != 0, opt_m = p.opt_m != 0, opt_f = p.opt_f != 0;
    double *accC = pl.smem_acc ? sAcc : nullptr;
    if (accC) { for (int i = tid; i < p.nrc * ACC_LD; i += ASM_THREADS) sAcc[i] = 0.0; __syncthreads(); }
    const int n_r = p.n_r, nrm1 = max(p.nrm - 1, 0);
    const double s1 = pl.s1, s2 = pl.s2;
    // this lane's two result elements (row g, columns 2q and 2q + 1)
    const int j0 = 2 * q, j1 = 2 * q + 1;
    const int i27_0 = packed27(g, j0), i27_1 = packed27(g, j1);
    const bool in36 = g < 6 && j1 < 6;
    const bool row_lt6 = g < 6, row_is6 = g == 6;
    const LanePtrs<JT> lp(Jn, g, q);
    const long long gw = (long long)blockIdx.x * ASM_WARPS + (tid >> 5), nw = (long long)gridDim.x * ASM_WARPS;
    for (long long pr = gw; pr < pl.npairs; pr += nw) {
        const int4 pi = pl.pair_info[pr];
        const int o0 = pi.x, n = pi.y, f = pi.z, c = pi.w;
        const bool act_c = opt_c && c != p.root_cam, act_cm = act_c && opt_m;
        const int cb = c - (c > p.root_cam ? 1 : 0);
        if (pr + nw < pl.npairs) {        // the next pair of this warp towards L2: its rows are consecutive
            const int4 pn = pl.pair_info[pr + nw];
            const char *nx = reinterpret_cast<const char *>(Jn + (size_t)pn.x * JROW);
            const int bytes = pn.y * JROW * (int)sizeof(JT);
            for (int b = lane * 128; b < bytes; b += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + b));
        }
        double Tff[2] = {0, 0}, Tcf[2] = {0, 0}, Tcc[2] = {0, 0};
        double *hcm_row = Hrr + (size_t)(6 * cb + min(g, 5)) * n_r + 6 * p.nrc + j0;      // this lane's element of marker block 0 in the camera's block row
        RowFrag<JT> xa, xb;
        xa.c.x = xa.c.y = xa.m.x = xa.m.y = xb.c.x = xb.c.y = xb.m.x = xb.m.y = (JT)0;
        auto load = [&](int o, RowFrag<JT> &x) {
            if (act_cm) load_frag<JT, true, true, true>(lp, o, x); else if (act_c) load_frag<JT, true, false, true>(lp, o, x); else load_frag<JT, false, false, true>(lp, o, x);
        };
        auto step = [&](int t, RowFrag<JT> &x, RowFrag<JT> &nx) {
            if (t + 1 < n) load(o0 + t + 1, nx);                                 // the next row in flight during this one
            double f0 = (double)x.f.x, f1 = (double)x.f.y;                       // [Jf | r]: rows 0..5 Jf, row 6 the residual
            if (Hw && row_is6) { const double w = Hw[(size_t)(o0 + t) * 4 + q]; f0 = w * f0; f1 = w * f1; }   // Huber: r = w * e (mcm.cpp:1014-1019)
            dmma884(Tff, f0, f0); dmma884(Tff, f1, f1);                         // Hff and, in column 6, gf
            if (act_c) {
                const double c0 = (double)x.c.x, c1 = (double)x.c.y;
                dmma884(Tcf, c0, f0); dmma884(Tcf, c1, f1);                     // W_c and, in column 6, gc
                dmma884(Tcc, c0, c0); dmma884(Tcc, c1, c1);
                if (act_cm) {
                    // camera x marker block of this observation: no other observation of the frame shares it.  A root-marker
                    // row has zero marker columns: zeros are added to a valid block.
                    double Tcm[2] = {0, 0};
                    dmma884(Tcm, c0, (double)x.m.x); dmma884(Tcm, c1, (double)x.m.y);
                    const int mk = frag_marker<JT>(x);
                    const int mb = min(mk - (mk > p.root_marker ? 1 : 0), nrm1);
                    if (in36) { double *dst = hcm_row + 6 * mb; atomicAdd(dst, Tcm[0] * s2); atomicAdd(dst + 1, Tcm[1] * s2); }
                }
            }
        };
        load(o0, xa);
        for (int t = 0; t < n; t += 2) {
            step(t, xa, xb);
            if (t + 1 < n) step(t + 1, xb, xa);
        }
        // ---- the pair's sums
        if (opt_f) {        // Hff + gf: every pair of the frame adds its share
            double *dst = Hf + (size_t)f * HF_STRIDE;
            if (i27_0 >= 0) atomicAdd(dst + i27_0, Tff[0] * (j0 == 6 ? s1 : s2));
            if (i27_1 >= 0) atomicAdd(dst + i27_1, Tff[1] * s2);
        }
        if (act_c) {
            if (opt_f && in36)      // W_c: this pair owns the slot
                *reinterpret_cast<double2 *>(W + (size_t)p.obs_slot_c[o0] * 36 + g * 6 + j0) = make_double2(Tcf[0] * s2, Tcf[1] * s2);
            // gc (column 6 of Jc^T [Jf | r]) and the upper triangle of Hcc
            if (accC) {
                double *dst = accC + cb * ACC_LD;
                if (row_lt6 && j0 == 6) atomicAdd(dst + 21 + g, Tcf[0]);
                if (i27_0 >= 0 && j0 < 6) atomicAdd(dst + i27_0, Tcc[0]);
                if (i27_1 >= 0 && j1 < 6) atomicAdd(dst + i27_1, Tcc[1]);
            } else {
                const int d0 = 6 * cb;
                if (row_lt6 && j0 == 6) atomicAdd(gr + d0 + g, Tcf[0] * s1);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = 2 * q + e; const double v = e ? Tcc[1] : Tcc[0];
                    if (row_lt6 && j < 6 && j >= g) { atomicAdd(Hrr + (size_t)(d0 + g) * n_r + d0 + j, v * s2); if (j != g) atomicAdd(Hrr + (size_t)(d0 + j) * n_r + d0 + g, v * s2); }
                }
            }
        }
    }
    if (accC) { __syncthreads(); flush_diag_blocks(accC, p.nrc, 0, n_r, s1, s2, Hrr, gr); }
}

// ------------------------------------------------------------------------------------------------
// (marker, camera) order inside each frame: one warp per (frame, marker) run.
template <typename JT>
__global__ void __launch_bounds__(ASM_THREADS, AAR_ASM_MINBLOCKS) k_asm_mruns(DevProblem p, AsmPlan pl, const JT *__restrict__ Jn, const double *__restrict__ Hw,
                                                                             double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    extern __shared__ __align__(16) double sAcc[];                 // [nrm][ACC_LD]
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const bool opt_f = p.opt_f != 0;
    double *accM = pl.smem_acc ? sAcc : nullptr;
    if (accM) { for (int i = tid; i < p.nrm * ACC_LD; i += ASM_THREADS) sAcc[i] = 0.0; __syncthreads(); }
    const int n_r = p.n_r;
    const double s1 = pl.s1, s2 = pl.s2;
    const int j0 = 2 * q, j1 = 2 * q + 1;
    const int i27_0 = packed27(g, j0), i27_1 = packed27(g, j1);
    const bool in36 = g < 6 && j1 < 6, row_lt6 = g < 6, row_is6 = g == 6;
    const LanePtrs<JT> lp(Jn, g, q);
    const long long gw = (long long)blockIdx.x * ASM_WARPS + (tid >> 5), nw = (long long)gridDim.x * ASM_WARPS;
    // run descriptors and row indices are fetched one run ahead, and the rows of the next run pulled towards L2
    int4 ri = gw < pl.nmruns ? pl.mrun_info[gw] : make_int4(0, 0, -1, 0);
    int my = lane < ri.y ? pl.perm_fm[ri.x + lane] : 0;
    for (long long r = gw; r < pl.nmruns; r += nw) {
        const int i0 = ri.x, n = ri.y, slot = ri.z, mb = ri.w, my_cur = my;
        if (r + nw < pl.nmruns) {
            ri = pl.mrun_info[r + nw];
            my = lane < ri.y ? pl.perm_fm[ri.x + lane] : 0;
            if (lane < ri.y) {
                const char *nx = reinterpret_cast<const char *>(Jn + (size_t)my * JROW + 48);      // marker + frame groups: 112 elements
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx)); asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128 * sizeof(JT) / 4));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 256 * sizeof(JT) / 4)); asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 112 * sizeof(JT) - 1));
            }
        }
        double Tmm[2] = {0, 0}, Tmf[2] = {0, 0};
        RowFrag<JT> xa, xb;
        xa.f.x = xa.f.y = xb.f.x = xb.f.y = (JT)0;
        auto row_of = [&](int t) { return t < 32 ? __shfl_sync(0xffffffffu, my_cur, t & 31) : pl.perm_fm[i0 + t]; };
        int o_cur = row_of(0);
        if (opt_f) load_frag<JT, false, true, true>(lp, o_cur, xa); else load_frag<JT, false, true, false>(lp, o_cur, xa);
        auto step = [&](int t, RowFrag<JT> &x, RowFrag<JT> &nx) {
            const int o_this = o_cur;
            if (t + 1 < n) { o_cur = row_of(t + 1); if (opt_f) load_frag<JT, false, true, true>(lp, o_cur, nx); else load_frag<JT, false, true, false>(lp, o_cur, nx); }
            double m0 = (double)x.m.x, m1 = (double)x.m.y;                       // [Jm | r]
            if (Hw && row_is6) { const double w = Hw[(size_t)o_this * 4 + q]; m0 = w * m0; m1 = w * m1; }
            dmma884(Tmm, m0, m0); dmma884(Tmm, m1, m1);                         // Hmm and, in column 6, gm
            if (opt_f) { dmma884(Tmf, m0, (double)x.f.x); dmma884(Tmf, m1, (double)x.f.y); }      // W_m
        };
        for (int t = 0; t < n; t += 2) {
            step(t, xa, xb);
            if (t + 1 < n) step(t + 1, xb, xa);
        }
        if (opt_f && slot >= 0 && in36)         // W_m: this run owns the slot
            *reinterpret_cast<double2 *>(W + (size_t)slot * 36 + g * 6 + j0) = make_double2(Tmf[0] * s2, Tmf[1] * s2);
        if (accM) {
            double *dst = accM + mb * ACC_LD;
            if (i27_0 >= 0) atomicAdd(dst + i27_0, Tmm[0]);
            if (i27_1 >= 0) atomicAdd(dst + i27_1, Tmm[1]);
        } else {
            const int d0 = 6 * (p.nrc + mb);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = 2 * q + e; const double v = e ? Tmm[1] : Tmm[0];
                if (row_lt6 && j == 6) atomicAdd(gr + d0 + g, v * s1);
                else if (row_lt6 && j < 6 && j >= g) { atomicAdd(Hrr + (size_t)(d0 + g) * n_r + d0 + j, v * s2); if (j != g) atomicAdd(Hrr + (size_t)(d0 + j) * n_r + d0 + g, v * s2); }
            }
        }
    }
    if (accM) { __syncthreads(); flush_diag_blocks(accM, p.nrm, p.nrc, n_r, s1, s2, Hrr, gr); }
}

} // namespace aar
