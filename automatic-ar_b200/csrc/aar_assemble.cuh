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
    const int *pair_slot;       // [npairs]  W slot of the pair's (frame, camera) block, -1: none
    const int4 *mrun_info;      // [nmruns]  first entry of perm_fm, entries, W slot (-1: none), reduced marker block
    const int *perm_fm;         // [nperm]   rows of each frame by (marker, camera); root-marker rows are not listed
    int npairs, nmruns, nperm;
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
#define AAR_ASM_MINBLOCKS 2
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

// ------------------------------------------------------------------------------------------------
// Row order: run = the rows of one (frame, camera) pair.  Every warp owns a contiguous share of the row stream (cut at pair
// boundaries, balanced by rows), so its ring never drains and the Hff / gf sums of a frame leave once per frame, not per pair.
template <typename JT>
__global__ void __launch_bounds__(ASM_THREADS, AAR_ASM_MINBLOCKS) k_asm_pairs(DevProblem p, AsmPlan pl, const JT *__restrict__ Jn /* [N][JROW] */, const double *__restrict__ Hw /* [N][4] Huber weights or null */,
                                                                             double *__restrict__ Hf, double *__restrict__ W, double *__restrict__ Hrr, double *__restrict__ gr) {
    typedef typename Vec2<JT>::type V2;
    constexpr int ROWB = JROW * (int)sizeof(JT), STAGE = asm_stage_bytes<JT>(JROW);
    extern __shared__ __align__(128) unsigned char asm_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const bool opt_c = p.opt_c != 0, opt_f = p.opt_f != 0, with_cm = opt_c && p.opt_m != 0;
    const size_t acc_bytes = pl.smem_acc ? asm_acc_bytes(p.nrc) : 0;
    double *accC = pl.smem_acc ? reinterpret_cast<double *>(asm_smem) : nullptr;
    unsigned char *ring = asm_smem + acc_bytes + ASM_BAR_BYTES + (size_t)warp * ASM_NST * STAGE;
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
    // this warp's share: pairs [pA, pB), rows [R0, R1)
    const long long gw = (long long)blockIdx.x * ASM_WARPS + warp, nw = (long long)gridDim.x * ASM_WARPS;
    const int pA = lower_bound_x(pl.pair_info, pl.npairs, (long long)p.N * gw / nw), pB = lower_bound_x(pl.pair_info, pl.npairs, (long long)p.N * (gw + 1) / nw);
    if (pA < pB) {
        const int R0 = pl.pair_info[pA].x, R1 = pB < pl.npairs ? pl.pair_info[pB].x : (int)p.N;
        const int nst = (R1 - R0 + ASM_SROWS - 1) / ASM_SROWS;
        auto issue = [&](int k) {            // stage k of the stream into slot k % ASM_NST
            if (lane == 0) {
                const int r = R0 + k * ASM_SROWS, nr = min(ASM_SROWS, R1 - r), slot = k % ASM_NST;
                const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
                mbar_expect_tx(bar, nr * (ROWB + (Hw ? 32 : 0)));
                bulk_g2s(dst, Jn + (size_t)r * JROW, nr * ROWB, bar);
                if (Hw) bulk_g2s(dst + ASM_SROWS * ROWB, Hw + (size_t)r * 4, nr * 32, bar);
            }
        };
        for (int k = 0; k < min(ASM_NST, nst); k++) issue(k);
        // run descriptors: lane j holds pair pbase + j
        int pbase = pA;
        int4 mine = pl.pair_info[min(pbase + lane, pB - 1)];
        int mslot = pl.pair_slot[min(pbase + lane, pB - 1)];
        int jp = 0, left = __shfl_sync(0xffffffffu, mine.y, 0), cam = __shfl_sync(0xffffffffu, mine.w, 0), frame = __shfl_sync(0xffffffffu, mine.z, 0);
        bool act_c = opt_c && cam != p.root_cam;
        int cb = min(cam - (cam > p.root_cam ? 1 : 0), nrc1);
        double *hcm_row = Hrr + (size_t)(6 * cb + min(g, 5)) * n_r + 6 * p.nrc + j0;      // this lane's element of marker block 0 in the camera's block row
        double Tff[2] = {0, 0}, Tcf[2] = {0, 0}, Tcc[2] = {0, 0};
        for (int k = 0; k < nst; k++) {
            const int slot = k % ASM_NST, nr = min(ASM_SROWS, R1 - (R0 + k * ASM_SROWS));
            mbar_wait(bar0 + 8 * slot, (k / ASM_NST) & 1);
            const unsigned char *sb = ring + slot * STAGE;
#pragma unroll 1
            for (int r = 0; r < nr; r++) {
                const unsigned char *row = sb + r * ROWB;
                const V2 xf = *reinterpret_cast<const V2 *>(row + off_f);
                double f0 = (double)xf.x, f1 = (double)xf.y;                   // [Jf | r]: rows 0..5 Jf, row 6 the residual
                if (Hw && row_is6) { const double w = reinterpret_cast<const double *>(sb + ASM_SROWS * ROWB)[r * 4 + q]; f0 = w * f0; f1 = w * f1; }   // Huber: r = w * e (mcm.cpp:1014-1019)
                dmma884(Tff, f0, f0); dmma884(Tff, f1, f1);                     // Hff and, in column 6, gf
                if (opt_c) {            // a root-camera row has zero camera columns: its products are zeros that nobody stores
                    const V2 xc = *reinterpret_cast<const V2 *>(row + off_c);
                    const double c0 = (double)xc.x, c1 = (double)xc.y;
                    dmma884(Tcf, c0, f0); dmma884(Tcf, c1, f1);                 // W_c and, in column 6, gc
                    dmma884(Tcc, c0, c0); dmma884(Tcc, c1, c1);
                    if (with_cm) {
                        // camera x marker block of this observation: no other observation of the frame shares it.  A root-marker
                        // row has zero marker columns: zeros are added to a valid block.
                        const V2 xm = *reinterpret_cast<const V2 *>(row + off_m);
                        double Tcm[2] = {0, 0};
                        dmma884(Tcm, c0, (double)xm.x); dmma884(Tcm, c1, (double)xm.y);
                        const int mk = *reinterpret_cast<const int *>(row + 152 * (int)sizeof(JT));
                        const int mb = min(mk - (mk > p.root_marker ? 1 : 0), nrm1);
                        if (in36 && act_c) { double *dst = hcm_row + 6 * mb; atomicAdd(dst, Tcm[0] * s2); atomicAdd(dst + 1, Tcm[1] * s2); }
                    }
                }
                if (--left == 0) {
                    // ---- the pair's sums
                    if (act_c) {
                        const int slot_c = __shfl_sync(0xffffffffu, mslot, jp);
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
                    Tcf[0] = Tcf[1] = Tcc[0] = Tcc[1] = 0.0;
                    // next pair of the stream
                    int nframe = -1;
                    if (++jp == 32 && pbase + 32 < pB) { pbase += 32; jp = 0; mine = pl.pair_info[min(pbase + lane, pB - 1)]; mslot = pl.pair_slot[min(pbase + lane, pB - 1)]; }
                    if (pbase + jp < pB) {
                        left = __shfl_sync(0xffffffffu, mine.y, jp); cam = __shfl_sync(0xffffffffu, mine.w, jp); nframe = __shfl_sync(0xffffffffu, mine.z, jp);
                        act_c = opt_c && cam != p.root_cam; cb = min(cam - (cam > p.root_cam ? 1 : 0), nrc1);
                        hcm_row = Hrr + (size_t)(6 * cb + min(g, 5)) * n_r + 6 * p.nrc + j0;
                    }
                    if (nframe != frame) {      // Hff + gf: the warp's pairs of this frame are done (other warps may add their share)
                        if (opt_f) {
                            double *dst = Hf + (size_t)frame * HF_STRIDE;
                            if (i27_0 >= 0) atomicAdd(dst + i27_0, Tff[0] * (j0 == 6 ? s1 : s2));
                            if (i27_1 >= 0) atomicAdd(dst + i27_1, Tff[1] * s2);
                        }
                        Tff[0] = Tff[1] = 0.0; frame = nframe;
                    }
                }
            }
            __syncwarp();                     // every lane is done reading the slot
            if (k + ASM_NST < nst) issue(k + ASM_NST);
        }
    }
    if (accC) { __syncthreads(); flush_diag_blocks(accC, p.nrc, 0, n_r, s1, s2, Hrr, gr); }
}

// ------------------------------------------------------------------------------------------------
// (marker, camera) order inside each frame: run = the rows of one (frame, marker), gathered through perm_fm — one bulk copy of
// the row's [Jm | Jf | e | marker] part (and one of its Huber weights) per row.
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
        const int E0 = pl.mrun_info[rA].x, E1 = rB < pl.nmruns ? pl.mrun_info[rB].x : pl.nperm;
        const int nst = (E1 - E0 + ASM_SROWS - 1) / ASM_SROWS;
        // row indices of the producer: lane l holds entry E0 + 32 * pblk + l, the next block is already on its way
        int pblk = 0;
        int pv = E0 + lane < E1 ? pl.perm_fm[E0 + lane] : 0, pvn = E0 + 32 + lane < E1 ? pl.perm_fm[E0 + 32 + lane] : 0;
        auto issue = [&](int k) {            // stage k of the stream into slot k % ASM_NST: one copy per row, issued by lanes 0 .. nr - 1
            const int e = k * ASM_SROWS, nr = min(ASM_SROWS, E1 - E0 - e), slot = k % ASM_NST;      // a stage never straddles a block of 32 entries
            if ((e >> 5) != pblk) { pblk++; pv = pvn; const int nx = E0 + 32 * (pblk + 1) + lane; pvn = nx < E1 ? pl.perm_fm[nx] : 0; }
            const int o = __shfl_sync(0xffffffffu, pv, (e & 31) + (lane & (ASM_SROWS - 1)));
            const unsigned bar = bar0 + 8 * slot, dst = ring0 + slot * STAGE;
            if (lane == 0) mbar_expect_tx(bar, nr * (ROWB + (Hw ? 32 : 0)));
            __syncwarp();
            if (lane < nr) {
                bulk_g2s(dst + lane * ROWB, Jn + (size_t)o * JROW + 48, ROWB, bar);
                if (Hw) bulk_g2s(dst + ASM_SROWS * ROWB + lane * 32, Hw + (size_t)o * 4, 32, bar);
            }
        };
        for (int k = 0; k < min(ASM_NST, nst); k++) issue(k);
        int rbase = rA;
        int4 mine = pl.mrun_info[min(rbase + lane, rB - 1)];
        int jr = 0, left = __shfl_sync(0xffffffffu, mine.y, 0);
        double Tmm[2] = {0, 0}, Tmf[2] = {0, 0};
        for (int k = 0; k < nst; k++) {
            const int slot = k % ASM_NST, nr = min(ASM_SROWS, E1 - E0 - k * ASM_SROWS);
            mbar_wait(bar0 + 8 * slot, (k / ASM_NST) & 1);
            const unsigned char *sb = ring + slot * STAGE;
#pragma unroll 1
            for (int r = 0; r < nr; r++) {
                const unsigned char *row = sb + r * ROWB;
                const V2 xm = *reinterpret_cast<const V2 *>(row + off_m);
                double m0 = (double)xm.x, m1 = (double)xm.y;                   // [Jm | r]
                if (Hw && row_is6) { const double w = reinterpret_cast<const double *>(sb + ASM_SROWS * ROWB)[r * 4 + q]; m0 = w * m0; m1 = w * m1; }
                dmma884(Tmm, m0, m0); dmma884(Tmm, m1, m1);                     // Hmm and, in column 6, gm
                if (opt_f) { const V2 xf = *reinterpret_cast<const V2 *>(row + off_f); dmma884(Tmf, m0, (double)xf.x); dmma884(Tmf, m1, (double)xf.y); }      // W_m
                if (--left == 0) {
                    const int slot_m = __shfl_sync(0xffffffffu, mine.z, jr), mb = __shfl_sync(0xffffffffu, mine.w, jr);
                    if (opt_f && slot_m >= 0 && in36)         // W_m: this run owns the slot
                        *reinterpret_cast<double2 *>(W + (size_t)slot_m * 36 + g * 6 + j0) = make_double2(Tmf[0] * s2, Tmf[1] * s2);
                    if (accM) {
                        double *dst = accM + mb * ACC_LD;
                        if (i27_0 >= 0) atomicAdd(dst + i27_0, Tmm[0]);
                        if (i27_1 >= 0) atomicAdd(dst + i27_1, Tmm[1]);
                    } else red_diag_block(Tmm, g, q, 6 * (p.nrc + mb), n_r, s1, s2, true, Hrr, gr);
                    Tmm[0] = Tmm[1] = Tmf[0] = Tmf[1] = 0.0;
                    if (++jr == 32 && rbase + 32 < rB) { rbase += 32; jr = 0; mine = pl.mrun_info[min(rbase + lane, rB - 1)]; }
                    if (rbase + jr < rB) left = __shfl_sync(0xffffffffu, mine.y, jr);
                }
            }
            __syncwarp();
            if (k + ASM_NST < nst) issue(k + ASM_NST);
        }
    }
    if (accM) { __syncthreads(); flush_diag_blocks(accM, p.nrm, p.nrc, n_r, s1, s2, Hrr, gr); }
}

} // namespace aar
