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
#include <type_traits>

namespace aar {

struct AsmPlan {
    const int4 *pair_info;      // [npairs]  (frame, camera) pairs in row order: first row, rows, local frame, W slot (-1: none)
    const int *pair_cam;        // [npairs]  camera
    int npairs, ring_pairs;     // stages per warp of k_asm_pairs
    const int4 *mrun_info;      // [nmruns]  first entry of perm_fm, entries, W slot (-1: none), reduced marker block
    const int *perm_fm;         // [nperm]   rows of each frame by (marker, camera); root-marker rows are not listed
    int nmruns, nperm;
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
// Rows reach the warps through per-warp rings of shared-memory stages filled by the bulk-copy engine (cp.async.bulk + mbarrier
// complete_tx): ASM_NST stages of ASM_SROWS rows are in flight per warp at no cost in registers or issue slots.  ncu of the
// version that loaded fragments straight from global memory (one row in flight per warp): 8.4 cycles of long-scoreboard stall
// per issued instruction, 40 % of the DRAM bandwidth; a register ring of four rows was slower still (profiles/r2_notes.md).
constexpr int ASM_NST = 4, ASM_SROWS = 4;
constexpr int ASM_MROW = 112;   // elements of a staged row the marker pass needs: [Jm (48) | Jf (48) | e (8) | marker index, pad (8)]
#ifndef AAR_ASM_MINBLOCKS
#define AAR_ASM_MINBLOCKS 3
#endif
template <typename JT> __host__ __device__ constexpr int asm_stage_bytes(int row_elems) { return ASM_SROWS * (row_elems * (int)sizeof(JT) + 32); }   // rows | Huber weights (4 doubles per row)
template <typename JT> __host__ __device__ constexpr size_t asm_ring_bytes(int row_elems) { return (size_t)ASM_WARPS * ASM_NST * asm_stage_bytes<JT>(row_elems); }
__host__ __device__ inline size_t asm_acc_bytes(int nblk) { return ((size_t)nblk * ACC_LD * sizeof(double) + 127) & ~(size_t)127; }
constexpr size_t ASM_BAR_BYTES = 256;   // ASM_WARPS x ASM_NST mbarriers

// ---- mbarrier / bulk copy (PTX ISA: mbarrier, cp.async.bulk)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

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
// a diagonal block's fragment straight to global memory (rigs whose accumulators do not fit shared memory)
__device__ __forceinline__ void red_diag_block(const double (&T)[2], int g, int q, int d0, int n_r, double s1, double s2, bool with_grad, double *__restrict__ Hrr, double *__restrict__ gr) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int j = 2 * q + e; const double v = T[e];
        if (g >= 6) continue;
        if (j == 6) { if (with_grad) atomicAdd(gr + d0 + g, v * s1); }
        else if (j < 6 && j >= g) { atomicAdd(Hrr + (size_t)(d0 + g) * n_r + d0 + j, v * s2); if (j != g) atomicAdd(Hrr + (size_t)(d0 + j) * n_r + d0 + g, v * s2); }
    }
}

// first index i in [0, n) with info[i].x >= key (info[].x ascending): where a warp's share of the row stream starts
__device__ __forceinline__ int lower_bound_x(const int4 *__restrict__ info, int n, long long key) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)info[mid].x < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// One lane's share of one staged row (JROW elements per observation, aar_jacobian.cuh): element pair (2q, 2q + 1) of fragment
// row g of each group, 8-byte (float staging) shared-memory loads at three per-lane offsets:
//   frame group   elements 96 + 8 g + 2q of [Jf (48) | e (8) | marker index, pad (8)] for every g: row 6 is the residual
//   marker group  48 + 8 g + 2q for g < 6; the residual (144 + 2q) for g == 6, so that the fragment is [Jm | r] as it stands
//   camera group  8 g + 2q for g < 6
// Whatever the lanes of fragment rows 6 / 7 hold beyond that only reaches rows / columns 6 and 7 of a product (element (i, j) of
// A^T B is the dot product of fragment row i of A with fragment row j of B), which are never read, except column 6 of a product
// with [X | r] — so nothing is masked.  Half-warps read 128 consecutive bytes (or a broadcast): no bank conflicts.

// The rows of one stage of k_asm_pairs as straight-line code per (row count, column groups present, Huber): ALL loads first — the
// warp-private camera x marker table lives in the same shared memory as the ring, so a load placed after a table store could not
// be moved above it by the compiler (ncu of the first version: rows strictly one after the other, 31 % of the cycles waiting on
// the scoreboard of the shared-memory loads) — then the conversions and products of all rows, then the table updates.
struct PairRowsCtx {
    const unsigned char *sb; int off_f, off_c, off_m, q, g, j0; bool row_is6, in36; double *cm_lane /* this lane's element of block (camera, 0) in the warp's table copy, or null */, *hcm_row; double s2; int root_marker, nrm1;
};
template <typename JT, int NR, int MODE /* 0: frame group only, 1: + camera, 2: + camera x marker */, bool HUBER>
__device__ __forceinline__ void pair_rows(const PairRowsCtx &cx, double (&Tff)[2], double (&Tcf)[2], double (&Tcc)[2]) {
    typedef typename Vec2<JT>::type V2;
    constexpr int ROWB = JROW * (int)sizeof(JT);
    V2 xf[NR], xc[NR], xm[NR]; int mk[NR]; double hw[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const unsigned char *row = cx.sb + r * ROWB;
        xf[r] = *reinterpret_cast<const V2 *>(row + cx.off_f);
        if (MODE >= 1) xc[r] = *reinterpret_cast<const V2 *>(row + cx.off_c);
        if (MODE >= 2) { xm[r] = *reinterpret_cast<const V2 *>(row + cx.off_m); mk[r] = *reinterpret_cast<const int *>(row + 152 * (int)sizeof(JT)); }
        if (HUBER) hw[r] = reinterpret_cast<const double *>(cx.sb + ASM_SROWS * ROWB)[r * 4 + cx.q];
    }
    double Tcm[NR][2];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        double f0 = (double)xf[r].x, f1 = (double)xf[r].y;                   // [Jf | r]: rows 0..5 Jf, row 6 the residual
        if (HUBER && cx.row_is6) { f0 = hw[r] * f0; f1 = hw[r] * f1; }       // Huber: r = w * e (mcm.cpp:1014-1019)
        dmma884(Tff, f0, f0); dmma884(Tff, f1, f1);                         // Hff and, in column 6, gf
        if (MODE >= 1) {
            const double c0 = (double)xc[r].x, c1 = (double)xc[r].y;
            dmma884(Tcf, c0, f0); dmma884(Tcf, c1, f1);                     // W_c and, in column 6, gc
            dmma884(Tcc, c0, c0); dmma884(Tcc, c1, c1);
            if (MODE >= 2) { Tcm[r][0] = Tcm[r][1] = 0.0; dmma884(Tcm[r], c0, (double)xm[r].x); dmma884(Tcm[r], c1, (double)xm[r].y); }
        }
    }
    if (MODE >= 2 && cx.in36) {
        // camera x marker block of each observation.  A root-marker row has zero marker columns: zeros are added to a valid block.
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int mb = min(mk[r] - (mk[r] > cx.root_marker ? 1 : 0), cx.nrm1);
            if (cx.cm_lane) { double *dst = cx.cm_lane + mb * 36; atomicAdd(dst, Tcm[r][0]); atomicAdd(dst + 1, Tcm[r][1]); }      // this warp's copy of the table, scaled in k_cm_reduce
            else { double *dst = cx.hcm_row + 6 * mb; atomicAdd(dst, Tcm[r][0] * cx.s2); atomicAdd(dst + 1, Tcm[r][1] * cx.s2); }
        }
    }
}
template <typename JT, int MODE, bool HUBER>
__device__ __forceinline__ void pair_rows_n(int nr, const PairRowsCtx &cx, double (&Tff)[2], double (&Tcf)[2], double (&Tcc)[2]) {
    switch (nr) {
        case 4: pair_rows<JT, 4, MODE, HUBER>(cx, Tff, Tcf, Tcc); break;
        case 3: pair_rows<JT, 3, MODE, HUBER>(cx, Tff, Tcf, Tcc); break;
        case 2: pair_rows<JT, 2, MODE, HUBER>(cx, Tff, Tcf, Tcc); break;
        default: pair_rows<JT, 1, MODE, HUBER>(cx, Tff, Tcf, Tcc);
    }
}
template <typename JT, bool HUBER>
__device__ __forceinline__ void pair_rows_dispatch(int nr, int mode, const PairRowsCtx &cx, double (&Tff)[2], double (&Tcf)[2], double (&Tcc)[2]) {
    if (mode == 2) pair_rows_n<JT, 2, HUBER>(nr, cx, Tff, Tcf, Tcc);
    else if (mode == 1) pair_rows_n<JT, 1, HUBER>(nr, cx, Tff, Tcf, Tcc);
    else pair_rows_n<JT, 0, HUBER>(nr, cx, Tff, Tcf, Tcc);
}

// the rows of one stage of k_asm_mruns, same scheme
template <typename JT, int NR, bool WITH_F, bool HUBER>
__device__ __forceinline__ void mrun_rows(const unsigned char *sb, int off_m, int off_f, int q, bool row_is6, double (&Tmm)[2], double (&Tmf)[2]) {
    typedef typename Vec2<JT>::type V2;
    constexpr int ROWB = ASM_MROW * (int)sizeof(JT);
    V2 xm[NR], xf[NR]; double hw[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const unsigned char *row = sb + r * ROWB;
        xm[r] = *reinterpret_cast<const V2 *>(row + off_m);
        if (WITH_F) xf[r] = *reinterpret_cast<const V2 *>(row + off_f);
        if (HUBER) hw[r] = reinterpret_cast<const double *>(sb + ASM_SROWS * ROWB)[r * 4 + q];
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
        double m0 = (double)xm[r].x, m1 = (double)xm[r].y;                   // [Jm | r]
        if (HUBER && row_is6) { m0 = hw[r] * m0; m1 = hw[r] * m1; }
        dmma884(Tmm, m0, m0); dmma884(Tmm, m1, m1);                         // Hmm and, in column 6, gm
        if (WITH_F) { dmma884(Tmf, m0, (double)xf[r].x); dmma884(Tmf, m1, (double)xf[r].y); }      // W_m
    }
}
template <typename JT, bool WITH_F, bool HUBER>
__device__ __forceinline__ void mrun_rows_n(int nr, const unsigned char *sb, int off_m, int off_f, int q, bool row_is6, double (&Tmm)[2], double (&Tmf)[2]) {
    switch (nr) {
        case 4: mrun_rows<JT, 4, WITH_F, HUBER>(sb, off_m, off_f, q, row_is6, Tmm, Tmf); break;
        case 3: mrun_rows<JT, 3, WITH_F, HUBER>(sb, off_m, off_f, q, row_is6, Tmm, Tmf); break;
        case 2: mrun_rows<JT, 2, WITH_F, HUBER>(sb, off_m, off_f, q, row_is6, Tmm, Tmf); break;
        default: mrun_rows<JT, 1, WITH_F, HUBER>(sb, off_m, off_f, q, row_is6, Tmm, Tmf);
    }
}

// ------------------------------------------------------------------------------------------------
// Row order: run = the rows of one (frame, camera) pair; a stage holds up to ASM_SROWS rows of ONE pair.  Every warp owns a
// contiguous share of the pairs (balanced by rows), so its ring never drains.
//   * W_c leaves with a plain store per pair (the pair owns its slot), Hff / gf with one RED per value and pair, Hcc / gc go to
//     CTA-lifetime shared accumulators (flushed once);
//   * the camera x marker blocks — one per observation, no run in any order inside a frame — leave as REDs into one of
//     ASM_CM_REPLICAS copies of a [camera][marker][36] table in FRAGMENT order (the 18 lanes of a block write 36 consecutive
//     doubles; k_cm_reduce adds the copies into the reduced matrix).  Measured on the way here (profiles/r2_notes.md): REDs
//     straight into the 272 KB camera x marker region of the reduced matrix ran at the device's single-copy RED rate (144 G
//     adds/s, 2.3 x the tensor-core time of the kernel; tools/red_bench.cu: 310 G/s coalesced over 64 copies); warp-private
//     shared-memory tables (camera-major jobs) removed the REDs but left room for 8 warps per SM only — slower still.
constexpr int ASM_CM_REPLICAS = 32;
template <typename JT>
__global__ void __launch_bounds__(ASM_THREADS, AAR_ASM_MINBLOCKS) k_asm_pairs(DevProblem p, AsmPlan pl, const JT *__restrict__ Jn /* [N][JROW] */, const double *__restrict__ Hw /* [N][4] Huber weights or null */,
                                                                             double *__restrict__ Hf, double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr, double *__restrict__ cm_rep) {
    constexpr int ROWB = JROW * (int)sizeof(JT), STAGE = asm_stage_bytes<JT>(JROW);
    extern __shared__ __align__(128) unsigned char asm_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const bool opt_c = p.opt_c != 0, opt_f = p.opt_f != 0, opt_m = p.opt_m != 0;
    const int NST = pl.ring_pairs;
    const size_t acc_bytes = pl.smem_acc ? asm_acc_bytes(p.nrc) : 0;
    double *accC = pl.smem_acc ? reinterpret_cast<double *>(asm_smem) : nullptr;
    unsigned char *ring = asm_smem + acc_bytes + ASM_BAR_BYTES + (size_t)warp * NST * STAGE;
    const unsigned bar0 = smem_u32(asm_smem + acc_bytes) + warp * ASM_NST * 8, ring0 = smem_u32(ring);
    if (accC) for (int i = tid; i < p.nrc * ACC_LD; i += ASM_THREADS) accC[i] = 0.0;
    if (lane < ASM_NST) mbar_init(bar0 + 8 * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_r = p.n_r, nrm1 = max(p.nrm - 1, 0), nrc1 = max(p.nrc - 1, 0);
    const double s1 = pl.s1, s2 = pl.s2;
    // this lane's two result elements (row g, columns 2q and 2q + 1)
    const int j0 = 2 * q, j1 = 2 * q + 1;
    const int i27_0 = packed27(g, j0), i27_1 = packed27(g, j1);
    const bool in36 = g < 6 && j1 < 6, row_lt6 = g < 6, row_is6 = g == 6;
    const int off_c = (g < 6 ? 8 * g + 2 * q : 144 + 2 * q) * (int)sizeof(JT), off_m = (g < 6 ? 48 + 8 * g + 2 * q : 144 + 2 * q) * (int)sizeof(JT), off_f = (96 + 8 * g + 2 * q) * (int)sizeof(JT);
    // this warp's share: pairs [pA, pB)
    const long long gw = (long long)blockIdx.x * ASM_WARPS + warp, nw = (long long)gridDim.x * ASM_WARPS;
    const int pA = lower_bound_x(pl.pair_info, pl.npairs, (long long)p.N * gw / nw), pB = lower_bound_x(pl.pair_info, pl.npairs, (long long)p.N * (gw + 1) / nw);
    double *cm_mine = cm_rep ? cm_rep + (size_t)(gw % ASM_CM_REPLICAS) * p.nrc * p.nrm * 36 + min(g, 5) * 6 + j0 : nullptr;      // this lane's element of block (0, 0) of this warp's copy
    if (pA < pB) {
        const int npr = pB - pA;
        // pair descriptors (first row, rows, frame, W slot) and cameras: lane j holds pair 32 * block + j of the share; two blocks resident
        int4 dcur = pl.pair_info[pA + min(lane, npr - 1)], dnxt = pl.pair_info[pA + min(32 + lane, npr - 1)];
        int ccur = pl.pair_cam[pA + min(lane, npr - 1)], cnxt = pl.pair_cam[pA + min(32 + lane, npr - 1)];
        int cblk = 0;
        // ---- producer: one stage = up to ASM_SROWS rows of one pair
        int pp = 0, prow = 0, pslot = 0, pfirst = __shfl_sync(0xffffffffu, dcur.x, 0), pn = __shfl_sync(0xffffffffu, dcur.y, 0);
        auto issue = [&]() {
            if (pp >= npr) return;
            const int nr = min(ASM_SROWS, pn - prow), r = pfirst + prow;
            if (lane == 0) {
                const unsigned bar = bar0 + 8 * pslot, dst = ring0 + pslot * STAGE;
                mbar_expect_tx(bar, nr * (ROWB + (Hw ? 32 : 0)));
                bulk_g2s(dst, Jn + (size_t)r * JROW, nr * ROWB, bar);
                if (Hw) bulk_g2s(dst + ASM_SROWS * ROWB, Hw + (size_t)r * 4, nr * 32, bar);
            }
            pslot = pslot + 1 == NST ? 0 : pslot + 1;
            prow += nr;
            if (prow == pn) {
                prow = 0;
                if (++pp < npr) {
                    const bool in_cur = (pp >> 5) == cblk;
                    pfirst = __shfl_sync(0xffffffffu, in_cur ? dcur.x : dnxt.x, pp & 31); pn = __shfl_sync(0xffffffffu, in_cur ? dcur.y : dnxt.y, pp & 31);
                }
            }
        };
        for (int k = 0; k < NST; k++) issue();
        // ---- consumer
        int cslot = 0; unsigned cphase = 0;
        for (int cp = 0; cp < npr; cp++) {
            if ((cp >> 5) != cblk) { cblk++; dcur = dnxt; ccur = cnxt; dnxt = pl.pair_info[pA + min(32 * (cblk + 1) + lane, npr - 1)]; cnxt = pl.pair_cam[pA + min(32 * (cblk + 1) + lane, npr - 1)]; }
            const int n = __shfl_sync(0xffffffffu, dcur.y, cp & 31), cam = __shfl_sync(0xffffffffu, ccur, cp & 31);
            const bool act_c = opt_c && cam != p.root_cam, with_cm = act_c && opt_m && p.nrm > 0;
            const int cb = min(cam - (cam > p.root_cam ? 1 : 0), nrc1);
            double Tff[2] = {0, 0}, Tcf[2] = {0, 0}, Tcc[2] = {0, 0};
            const PairRowsCtx cx0{nullptr, off_f, off_c, off_m, q, g, j0, row_is6, in36, cm_mine ? cm_mine + (size_t)cb * p.nrm * 36 : nullptr,
                                  Hrr + (size_t)(6 * cb + min(g, 5)) * n_r + 6 * p.nrc + j0, s2, p.root_marker, nrm1};
            const int mode = with_cm ? 2 : (act_c ? 1 : 0);
            for (int r0 = 0; r0 < n; r0 += ASM_SROWS) {
                const int nr = min(ASM_SROWS, n - r0);
                mbar_wait(bar0 + 8 * cslot, (cphase >> cslot) & 1);
                PairRowsCtx cx = cx0; cx.sb = ring + cslot * STAGE;
                if (Hw) pair_rows_dispatch<JT, true>(nr, mode, cx, Tff, Tcf, Tcc); else pair_rows_dispatch<JT, false>(nr, mode, cx, Tff, Tcf, Tcc);
                __syncwarp();                     // every lane is done reading the slot
                cphase ^= 1u << cslot; cslot = cslot + 1 == NST ? 0 : cslot + 1;
                issue();
            }
            // ---- the pair's sums
            if (opt_f) {        // Hff + gf: every pair of the frame adds its share
                double *dst = Hf + (size_t)__shfl_sync(0xffffffffu, dcur.z, cp & 31) * HF_STRIDE;
                if (i27_0 >= 0) atomicAdd(dst + i27_0, Tff[0] * (j0 == 6 ? s1 : s2));
                if (i27_1 >= 0) atomicAdd(dst + i27_1, Tff[1] * s2);
            }
            const int slot_c = __shfl_sync(0xffffffffu, dcur.w, cp & 31);
            if (act_c) {
                if (opt_f && in36 && slot_c >= 0)      // W_c: this pair owns the slot
                    *reinterpret_cast<double2 *>(W + (size_t)slot_c * 36 + g * 6 + j0) = make_double2(Tcf[0] * s2, Tcf[1] * s2);
                // gc (column 6 of Jc^T [Jf | r]) and the upper triangle of Hcc
                if (accC) {
                    double *dst = accC + cb * ACC_LD;
                    if (row_lt6 && j0 == 6) atomicAdd(dst + 21 + g, Tcf[0]);
                    if (i27_0 >= 0 && j0 < 6) atomicAdd(dst + i27_0, Tcc[0]);
                    if (i27_1 >= 0 && j1 < 6) atomicAdd(dst + i27_1, Tcc[1]);
                } else {
                    if (row_lt6 && j0 == 6) atomicAdd(gr + 6 * cb + g, Tcf[0] * s1);
                    red_diag_block(Tcc, g, q, 6 * cb, n_r, s1, s2, false, Hrr, gr);
                }
            }
        }
    }
    if (accC) { __syncthreads(); flush_diag_blocks(accC, p.nrc, 0, n_r, s1, s2, Hrr, gr); }
}
// adds the ASM_CM_REPLICAS copies of the [camera][marker][36] table into the camera x marker region of the reduced matrix
__global__ void k_cm_reduce(int nrc, int nrm, int n_r, double s2, const double *__restrict__ cm_rep, double *__restrict__ Hrr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = nrc * nrm * 36;
    if (i >= n) return;
    double v = 0;
#pragma unroll 4
    for (int k = 0; k < ASM_CM_REPLICAS; k++) v += cm_rep[(size_t)k * n + i];
    if (v == 0.0) return;
    const int blk = i / 36, e = i - 36 * blk, cb = blk / nrm, mb = blk - cb * nrm;
    atomicAdd(Hrr + (size_t)(6 * cb + e / 6) * n_r + 6 * nrc + 6 * mb + e % 6, v * s2);
}

// ------------------------------------------------------------------------------------------------
// (marker, camera) order inside each frame: run = the rows of one (frame, marker), gathered through perm_fm — one bulk copy of
// the row's [Jm | Jf | e | marker] part (and one of its Huber weights) per row; a stage holds up to ASM_SROWS rows of ONE run, so
// the rows of a stage are multiplied in one unrolled block and the run's sums leave at a stage boundary.  Every warp owns a
// contiguous share of the runs (balanced by rows).
template <typename JT>
__global__ void __launch_bounds__(ASM_THREADS, AAR_ASM_MINBLOCKS) k_asm_mruns(DevProblem p, AsmPlan pl, const JT *__restrict__ Jn, const double *__restrict__ Hw,
                                                                             double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    typedef typename Vec2<JT>::type V2;
    constexpr int ROWB = ASM_MROW * (int)sizeof(JT), STAGE = asm_stage_bytes<JT>(ASM_MROW);
    extern __shared__ __align__(128) unsigned char asm_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const bool opt_f = p.opt_f != 0;
    const size_t acc_bytes = pl.smem_acc ? asm_acc_bytes(p.nrm) : 0;
    double *accM = pl.smem_acc ? reinterpret_cast<double *>(asm_smem) : nullptr;
    unsigned char *ring = asm_smem + acc_bytes + ASM_BAR_BYTES + (size_t)warp * ASM_NST * STAGE;
    const unsigned bar0 = smem_u32(asm_smem + acc_bytes) + warp * ASM_NST * 8, ring0 = smem_u32(ring);
    if (accM) for (int i = tid; i < p.nrm * ACC_LD; i += ASM_THREADS) accM[i] = 0.0;
    if (lane < ASM_NST) mbar_init(bar0 + 8 * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int n_r = p.n_r;
    const double s1 = pl.s1, s2 = pl.s2;
    const int j0 = 2 * q, j1 = 2 * q + 1;
    const int i27_0 = packed27(g, j0), i27_1 = packed27(g, j1);
    const bool in36 = g < 6 && j1 < 6, row_is6 = g == 6;
    const int off_m = (g < 6 ? 8 * g + 2 * q : 96 + 2 * q) * (int)sizeof(JT), off_f = (48 + 8 * g + 2 * q) * (int)sizeof(JT);      // relative to element 48 of the row
    // this warp's share: runs [rA, rB), entries [E0, E1) of perm_fm
    const long long gw = (long long)blockIdx.x * ASM_WARPS + warp, nw = (long long)gridDim.x * ASM_WARPS;
    const int rA = lower_bound_x(pl.mrun_info, pl.nmruns, (long long)pl.nperm * gw / nw), rB = lower_bound_x(pl.mrun_info, pl.nmruns, (long long)pl.nperm * (gw + 1) / nw);
    if (rA < rB) {
        const int nrun = rB - rA;
        // run descriptors (first entry, rows, W slot, marker block): lane j holds run 32 * block + j of the share; two blocks resident
        int4 dcur = pl.mrun_info[rA + min(lane, nrun - 1)], dnxt = pl.mrun_info[rA + min(32 + lane, nrun - 1)];
        int cblk = 0;
        const int E0 = __shfl_sync(0xffffffffu, dcur.x, 0), E1 = rB < pl.nmruns ? pl.mrun_info[rB].x : pl.nperm;
        // row indices of the producer: lane l holds entry E0 + 32 * pblk + l, the next block is already on its way
        int pblk = 0;
        int pv = E0 + lane < E1 ? pl.perm_fm[E0 + lane] : 0, pvn = E0 + 32 + lane < E1 ? pl.perm_fm[E0 + 32 + lane] : 0;
        // ---- producer: one stage = up to ASM_SROWS rows of one run, one copy per row issued by lanes 0 .. nr - 1
        int pp = 0, prow = 0, pslot = 0, pfirst = E0, pn = __shfl_sync(0xffffffffu, dcur.y, 0);
        auto issue = [&]() {
            if (pp >= nrun) return;
            const int nr = min(ASM_SROWS, pn - prow), e = pfirst + prow - E0;      // entries e .. e + nr - 1 of the share
            if ((e >> 5) > pblk) { pblk++; pv = pvn; const int nx = E0 + 32 * (pblk + 1) + lane; pvn = nx < E1 ? pl.perm_fm[nx] : 0; }
            const int x = e + (lane & (ASM_SROWS - 1));
            const int oa = __shfl_sync(0xffffffffu, pv, x & 31), ob = __shfl_sync(0xffffffffu, pvn, x & 31);
            const int o = (x >> 5) == pblk ? oa : ob;
            const unsigned bar = bar0 + 8 * pslot, dst = ring0 + pslot * STAGE;
            if (lane == 0) mbar_expect_tx(bar, nr * (ROWB + (Hw ? 32 : 0)));
            __syncwarp();
            if (lane < nr) {
                bulk_g2s(dst + lane * ROWB, Jn + (size_t)o * JROW + 48, ROWB, bar);
                if (Hw) bulk_g2s(dst + ASM_SROWS * ROWB + lane * 32, Hw + (size_t)o * 4, 32, bar);
            }
            pslot = pslot + 1 == ASM_NST ? 0 : pslot + 1;
            prow += nr;
            if (prow == pn) {
                prow = 0;
                if (++pp < nrun) {
                    const bool in_cur = (pp >> 5) == cblk;
                    pfirst = __shfl_sync(0xffffffffu, in_cur ? dcur.x : dnxt.x, pp & 31); pn = __shfl_sync(0xffffffffu, in_cur ? dcur.y : dnxt.y, pp & 31);
                }
            }
        };
        for (int k = 0; k < ASM_NST; k++) issue();
        // ---- consumer
        int cslot = 0; unsigned cphase = 0;
        for (int cr = 0; cr < nrun; cr++) {
            if ((cr >> 5) != cblk) { cblk++; dcur = dnxt; dnxt = pl.mrun_info[rA + min(32 * (cblk + 1) + lane, nrun - 1)]; }
            const int n = __shfl_sync(0xffffffffu, dcur.y, cr & 31);
            double Tmm[2] = {0, 0}, Tmf[2] = {0, 0};
            for (int r0 = 0; r0 < n; r0 += ASM_SROWS) {
                const int nr = min(ASM_SROWS, n - r0);
                mbar_wait(bar0 + 8 * cslot, (cphase >> cslot) & 1);
                const unsigned char *sb = ring + cslot * STAGE;
                if (opt_f) { if (Hw) mrun_rows_n<JT, true, true>(nr, sb, off_m, off_f, q, row_is6, Tmm, Tmf); else mrun_rows_n<JT, true, false>(nr, sb, off_m, off_f, q, row_is6, Tmm, Tmf); }
                else { if (Hw) mrun_rows_n<JT, false, true>(nr, sb, off_m, off_f, q, row_is6, Tmm, Tmf); else mrun_rows_n<JT, false, false>(nr, sb, off_m, off_f, q, row_is6, Tmm, Tmf); }
                __syncwarp();                     // every lane is done reading the slot
                cphase ^= 1u << cslot; cslot = cslot + 1 == ASM_NST ? 0 : cslot + 1;
                issue();
            }
            // ---- the run's sums
            const int slot_m = __shfl_sync(0xffffffffu, dcur.z, cr & 31), mb = __shfl_sync(0xffffffffu, dcur.w, cr & 31);
            if (opt_f && slot_m >= 0 && in36)         // W_m: this run owns the slot
                *reinterpret_cast<double2 *>(W + (size_t)slot_m * 36 + g * 6 + j0) = make_double2(Tmf[0] * s2, Tmf[1] * s2);
            if (accM) {
                double *dst = accM + mb * ACC_LD;
                if (i27_0 >= 0) atomicAdd(dst + i27_0, Tmm[0]);
                if (i27_1 >= 0) atomicAdd(dst + i27_1, Tmm[1]);
            } else red_diag_block(Tmm, g, q, 6 * (p.nrc + mb), n_r, s1, s2, true, Hrr, gr);
        }
    }
    if (accM) { __syncthreads(); flush_diag_blocks(accM, p.nrm, p.nrc, n_r, s1, s2, Hrr, gr); }
}

} // namespace aar
