// aar_dense.cuh — dense Cholesky solve of the reduced camera+marker system (S + mu I) x = b, the part of
// sparselevmarq.h:394-400 that is left after the frames are eliminated.  n_r = 6(C-1) + 6(M-1) is a few hundred at
// most (468 at BASELINE cfg 4): far too small for tensor cores to matter, far too large for one thread block's
// shared memory, and on the critical path of every LM try.  Left-looking blocked factorisation, one CTA per block
// row of 32, two grid-wide synchronisations per block column (cooperative launch):
//   step k:  every CTA i >= k:  A_ik -= sum_{j<k} L_ij L_kj^T          (32x32x32 tile products from L2)
//            CTA k:             L_kk = chol(A_kk)  in shared memory      -> grid sync
//            every CTA i > k:   L_ik = A_ik L_kk^-T                      -> grid sync
// then forward / backward substitution by CTA 0.  S: row-major n x n, UPPER triangle valid on entry (the Schur
// kernels only write the upper block triangle); the lower triangle holds L on exit.
#pragma once
#include <cooperative_groups.h>

namespace aar {

constexpr int CH_NB = 32;
constexpr int CH_THREADS = 256;
constexpr int CH_LD = CH_NB + 1;

__global__ void __launch_bounds__(CH_THREADS) k_reduced_solve(int n, double *__restrict__ S, const double *__restrict__ b, double *__restrict__ x, const LmState *__restrict__ st,
                                                              int *__restrict__ chol_fail) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sA[CH_NB * CH_LD], sB[CH_NB * CH_LD], sC[CH_NB * CH_LD];
    __shared__ double sv[CH_NB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nblk = (n + CH_NB - 1) / CH_NB;
    const double mu = st->mu;
    // lower <- upper, diagonal += mu
    for (long long e = (long long)blockIdx.x * CH_THREADS + tid; e < (long long)n * n; e += (long long)gridDim.x * CH_THREADS) {
        const int i = (int)(e / n), j = (int)(e % n);
        if (i > j) S[e] = S[(size_t)j * n + i];
        else if (i == j) S[e] += mu;
    }
    grid.sync();
    auto load_tile = [&](double *dst, int bi, int bj) {       // rows of block bi, columns of block bj, zero padded
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int r = e / CH_NB, c = e % CH_NB, gr = bi * CH_NB + r, gc = bj * CH_NB + c;
            dst[r * CH_LD + c] = (gr < n && gc < n) ? S[(size_t)gr * n + gc] : 0.0;
        }
    };
    for (int k = 0; k < nblk; k++) {
        for (int i = k + blockIdx.x; i < nblk; i += gridDim.x) {
            // ---- A_ik -= sum_{j<k} L_ij L_kj^T ; each thread owns 2x2 entries of the tile
            double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
            for (int j = 0; j < k; j++) {
                __syncthreads();
                load_tile(sA, i, j); load_tile(sB, k, j);
                __syncthreads();
#pragma unroll 8
                for (int q = 0; q < CH_NB; q++) {
                    const double a0 = sA[ty * CH_LD + q], a1 = sA[(ty + 16) * CH_LD + q], b0 = sB[tx * CH_LD + q], b1 = sB[(tx + 16) * CH_LD + q];
                    c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
                }
            }
            __syncthreads();
            load_tile(sC, i, k);
            __syncthreads();
            sC[ty * CH_LD + tx] -= c00; sC[ty * CH_LD + tx + 16] -= c01; sC[(ty + 16) * CH_LD + tx] -= c10; sC[(ty + 16) * CH_LD + tx + 16] -= c11;
            __syncthreads();
            if (i == k) {
                // ---- L_kk = chol(A_kk) in shared memory (padding rows/cols beyond n get a unit diagonal)
                for (int e = tid; e < CH_NB; e += CH_THREADS) if (k * CH_NB + e >= n) sC[e * CH_LD + e] = 1.0;
                __syncthreads();
                for (int j = 0; j < CH_NB; j++) {
                    if (tid == 0) { double d = sC[j * CH_LD + j]; if (!(d > 0)) { atomicExch(chol_fail, 1); d = 1; } sv[0] = sqrt(d); }
                    __syncthreads();
                    const double dj = sv[0];
                    if (tid > j && tid < CH_NB) sC[tid * CH_LD + j] /= dj;
                    if (tid == j) sC[j * CH_LD + j] = dj;
                    __syncthreads();
                    // trailing update of the lower triangle: rows r > j, columns j < c <= r
                    for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                        const int r = e / CH_NB, c = e % CH_NB;
                        if (c > j && r >= c) sC[r * CH_LD + c] = fma(-sC[r * CH_LD + j], sC[c * CH_LD + j], sC[r * CH_LD + c]);
                    }
                    __syncthreads();
                }
            }
            // write the tile back (A_ik updated, or L_kk: lower triangle only)
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                const int r = e / CH_NB, c = e % CH_NB, gr = i * CH_NB + r, gc = k * CH_NB + c;
                if (gr < n && gc < n && (i != k || c <= r)) S[(size_t)gr * n + gc] = sC[r * CH_LD + c];
            }
        }
        grid.sync();
        // ---- L_ik = A_ik L_kk^-T for i > k: one thread per row, forward substitution against L_kk
        {
            bool loaded = false;
            for (int i = k + 1 + blockIdx.x; i < nblk; i += gridDim.x) {
                if (!loaded) { __syncthreads(); load_tile(sB, k, k); loaded = true; }
                __syncthreads();
                load_tile(sC, i, k);
                __syncthreads();
                if (tid < CH_NB) {
                    double *row = sC + tid * CH_LD;
                    for (int c = 0; c < CH_NB; c++) {
                        double v = row[c];
                        for (int q = 0; q < c; q++) v = fma(-row[q], sB[c * CH_LD + q], v);
                        const double d = sB[c * CH_LD + c];
                        row[c] = d != 0.0 ? v / d : 0.0;       // padding columns of the last block have no diagonal
                    }
                }
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                    const int r = e / CH_NB, c = e % CH_NB, gr = i * CH_NB + r, gc = k * CH_NB + c;
                    if (gr < n && gc < n) S[(size_t)gr * n + gc] = sC[r * CH_LD + c];
                }
            }
        }
        grid.sync();
    }
    // ---- L y = b, L^T x = y (CTA 0; dot products spread over the block, one warp per 4 rows)
    if (blockIdx.x != 0) return;
    for (int i = tid; i < n; i += CH_THREADS) x[i] = b[i];
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int kb = 0; kb < nblk; kb++) {
        const int r0 = kb * CH_NB, nr = min(CH_NB, n - r0);
        // rows of this block minus the contribution of the solved part
        for (int r = warp; r < nr; r += CH_THREADS / 32) {
            double s = 0;
            for (int c = lane; c < r0; c += 32) s = fma(S[(size_t)(r0 + r) * n + c], x[c], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sv[r] = x[r0 + r] - s;
        }
        __syncthreads();
        if (warp == 0) {   // 32x32 triangular solve, lane r owns row r
            double v = lane < nr ? sv[lane] : 0.0;
            for (int c = 0; c < nr; c++) {
                const double d = S[(size_t)(r0 + c) * n + r0 + c];
                const double xc = __shfl_sync(0xffffffffu, v, c) / d;
                if (lane == c) v = xc;
                else if (lane > c && lane < nr) v = fma(-S[(size_t)(r0 + lane) * n + r0 + c], xc, v);
            }
            if (lane < nr) x[r0 + lane] = v;
        }
        __syncthreads();
    }
    for (int kb = nblk - 1; kb >= 0; kb--) {
        const int r0 = kb * CH_NB, nr = min(CH_NB, n - r0), c0 = r0 + nr;
        for (int r = warp; r < nr; r += CH_THREADS / 32) {
            double s = 0;
            for (int c = c0 + lane; c < n; c += 32) s = fma(S[(size_t)c * n + r0 + r], x[c], s);     // L^T: column r0+r below the block
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sv[r] = x[r0 + r] - s;
        }
        __syncthreads();
        if (warp == 0) {
            double v = lane < nr ? sv[lane] : 0.0;
            for (int c = nr - 1; c >= 0; c--) {
                const double d = S[(size_t)(r0 + c) * n + r0 + c];
                const double xc = __shfl_sync(0xffffffffu, v, c) / d;
                if (lane == c) v = xc;
                else if (lane < c) v = fma(-S[(size_t)(r0 + c) * n + r0 + lane], xc, v);
            }
            if (lane < nr) x[r0 + lane] = v;
        }
        __syncthreads();
    }
}

} // namespace aar
