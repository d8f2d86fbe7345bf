// aar_schur.cuh — elimination of the per-frame 6x6 blocks (replaces the sparse LDLT of
// /root/reference/libs/sparselevmarq.h:394-400 on the arrow-shaped JtJ + mu I):
//
//   D_f = Hff_f + mu I = L_f L_f^T          y_f = L_f^-1 B_f            (k_frame_chol, one thread per frame)
//   E_s = W_s L_f^-T   for every W slot s    b[blk(s)] -= E_s y_f         (k_schur_prepare, one thread per slot row)
//   S[blk(s), blk(t)] -= E_s E_t^T  over all slot pairs of every frame    (k_schur_syrk, output-stationary tiles)
//
// The last step is a block-sparse SYRK of an n_r x 6F matrix.  S does not fit shared memory (468^2 doubles at
// BASELINE cfg 4), so the output is tiled: a CTA owns a 96x96 tile of S (16x16 blocks of 6x6, one block pair —
// 36 FP64 accumulators — per thread) and streams over a chunk of frames, staging the E blocks of its 16 row
// blocks and 16 column blocks through shared memory; the partial tiles of the frame chunks leave with one RED
// per entry.  DESIGN.md ("k_schur_syrk") has the roofline.
#pragma once

namespace aar {

constexpr int FC_STRIDE = 28;      // per frame: L (lower, packed by rows, 21) | y (6) | pad
constexpr int SY_TB = 16;          // 6x6 blocks per tile side
constexpr int SY_LD = 37;          // padded block stride in shared memory (conflict-free 64-bit loads across blocks)
constexpr int SY_FB = 6;           // frames per staged batch
constexpr int SY_THREADS = 256;
constexpr size_t SY_SMEM = sizeof(double) * 2 * SY_FB * SY_TB * SY_LD + sizeof(int) * 2 * SY_FB * SY_TB;

// D = Hff + mu I = L L^T, y = L^-1 (-gf).  Non-positive pivot -> chol_fail (the reference does not check its LDLT,
// sparselevmarq.h:394-400; the host treats it as a rejected step).
__global__ void k_frame_chol(DevProblem p, const LmState *__restrict__ st, const double *__restrict__ Hf, double *__restrict__ fc, int *__restrict__ chol_fail) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.F) return;
    double L[36], y[6];
    if (!chol6(Hf + (size_t)f * HF_STRIDE, st->mu, L)) atomicExch(chol_fail, 1);
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = -Hf[(size_t)f * HF_STRIDE + 21 + i];
    fwd6(L, y);
    double *dst = fc + (size_t)f * FC_STRIDE;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) dst[idx++] = L[i * 6 + j];
#pragma unroll
    for (int i = 0; i < 6; i++) dst[21 + i] = y[i];
}

// One thread per (slot, row i): E_s[i][:] = L^-1 W_s[i][:]^T, and the row's share of b[blk(s)] -= E_s y.
__global__ void __launch_bounds__(256) k_schur_prepare(DevProblem p, long long nslots, const int *__restrict__ slot_frame, const double *__restrict__ fc,
                                                       const double *__restrict__ W, double *__restrict__ E, double *__restrict__ b) {
    extern __shared__ double sb[];   // [n_r] partial b of this CTA
    for (int i = threadIdx.x; i < p.n_r; i += blockDim.x) sb[i] = 0.0;
    __syncthreads();
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nslots * 6; g += (long long)gridDim.x * blockDim.x) {
        const long long s = g / 6; const int i = (int)(g % 6);
        const double *l = fc + (size_t)slot_frame[s] * FC_STRIDE;
        double x[6];
#pragma unroll
        for (int k = 0; k < 6; k++) x[k] = W[(size_t)g * 6 + k];
        // forward substitution with the packed lower factor
        int idx = 0; double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double v = x[r];
#pragma unroll
            for (int k = 0; k < r; k++) v = fma(-l[idx + k], x[k], v);
            x[r] = v / l[idx + r];
            idx += r + 1;
            acc = fma(x[r], l[21 + r], acc);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) E[(size_t)g * 6 + k] = x[k];
        atomicAdd(sb + 6 * p.slot_block[s] + i, -acc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.n_r; i += blockDim.x) if (sb[i] != 0.0) atomicAdd(b + i, sb[i]);
}

// S[tile] -= sum over the frames of this chunk of E_I E_J^T.  frame_block_slot[f*nb + blk] is the W slot of
// reduced block blk in frame f, or -1.  gridDim.x = ntiles * nchunks (tile fastest).
__global__ void __launch_bounds__(SY_THREADS, 1) k_schur_syrk(DevProblem p, int nb, int tiles_side, int nchunks, const int *__restrict__ frame_block_slot,
                                                              const double *__restrict__ E, double *__restrict__ S) {
    extern __shared__ __align__(16) unsigned char sy_raw[];
    double (*sE)[SY_FB][SY_TB * SY_LD] = reinterpret_cast<double (*)[SY_FB][SY_TB * SY_LD]>(sy_raw);            // [I|J][frame in batch][block][36 (+1)]
    int (*sPresent)[SY_FB][SY_TB] = reinterpret_cast<int (*)[SY_FB][SY_TB]>(sy_raw + sizeof(double) * 2 * SY_FB * SY_TB * SY_LD);
    const int ntiles = tiles_side * (tiles_side + 1) / 2;
    int tile = blockIdx.x % ntiles; const int chunk = blockIdx.x / ntiles;
    int ti = 0;
    while (tile >= tiles_side - ti) { tile -= tiles_side - ti; ti++; }
    const int tj = ti + tile;
    const int tid = threadIdx.x, bi = tid / SY_TB, bj = tid % SY_TB;
    const int f0 = (int)((long long)p.F * chunk / nchunks), f1 = (int)((long long)p.F * (chunk + 1) / nchunks);
    const int gi = ti * SY_TB + bi, gj = tj * SY_TB + bj;                    // global block indices of this thread's pair
    const bool mine = gi < nb && gj < nb && (ti != tj || bi <= bj);          // upper block triangle only
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = 0.0;
    for (int fb0 = f0; fb0 < f1; fb0 += SY_FB) {
        const int nf = min(SY_FB, f1 - fb0);
        __syncthreads();
        // stage: 2 sides x nf frames x 16 blocks x 36 doubles
        for (int e = tid; e < 2 * nf * SY_TB; e += SY_THREADS) {
            const int side = e / (nf * SY_TB), rem = e % (nf * SY_TB), ff = rem / SY_TB, blk = rem % SY_TB;
            const int g = (side ? tj : ti) * SY_TB + blk;
            sPresent[side][ff][blk] = g < nb ? frame_block_slot[(size_t)(fb0 + ff) * nb + g] : -1;
        }
        __syncthreads();
        for (int e = tid; e < 2 * nf * SY_TB * 36; e += SY_THREADS) {
            const int k = e % 36, rem = e / 36, side = rem / (nf * SY_TB), rem2 = rem % (nf * SY_TB), ff = rem2 / SY_TB, blk = rem2 % SY_TB;
            const int slot = sPresent[side][ff][blk];
            sE[side][ff][blk * SY_LD + k] = slot >= 0 ? E[(size_t)slot * 36 + k] : 0.0;
        }
        __syncthreads();
        if (mine) {
            for (int ff = 0; ff < nf; ff++) {
                if (sPresent[0][ff][bi] < 0 || sPresent[1][ff][bj] < 0) continue;
                const double *ei = &sE[0][ff][bi * SY_LD], *ej = &sE[1][ff][bj * SY_LD];
                double a[36];
#pragma unroll
                for (int i = 0; i < 36; i++) a[i] = ei[i];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double bcol[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) bcol[k] = ej[c * 6 + k];
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double s = acc[r * 6 + c];
#pragma unroll
                        for (int k = 0; k < 6; k++) s = fma(a[r * 6 + k], bcol[k], s);
                        acc[r * 6 + c] = s;
                    }
                }
            }
        }
    }
    if (mine) {
        double *dst = S + (size_t)(6 * gi) * p.n_r + 6 * gj;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) if (acc[r * 6 + c] != 0.0) atomicAdd(dst + (size_t)r * p.n_r + c, -acc[r * 6 + c]);
    }
}

} // namespace aar
