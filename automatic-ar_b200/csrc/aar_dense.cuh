// aar_dense.cuh — dense Cholesky solve of the reduced camera+marker system (S + mu I) x = b, the part of
// sparselevmarq.h:394-400 that is left after the frames are eliminated.  n_r = 6(C-1) + 6(M-1) is a few hundred at
// most (468 at BASELINE cfg 4): far too small for tensor cores to matter, far too large for one thread block's
// shared memory, and on the critical path of every LM try.  Left-looking blocked factorisation, one CTA per block
// row of 32, two grid-wide synchronisations per block column (cooperative launch):
//   step k:  every CTA i >= k:  A_ik -= sum_{j<k} L_ij L_kj^T          (32x32x32 tile products from L2)
//            CTA k:             L_kk = chol(A_kk), X_k = L_kk^-1 in shared memory -> grid sync
//            every CTA i > k:   L_ik = A_ik X_k^T   (a tile product, not a row-serial solve) -> grid sync
// then forward / backward substitution by CTA 0.  S: row-major n x n, UPPER triangle valid on entry (the Schur
// kernels only write the upper block triangle); the lower triangle holds L on exit.
#pragma once
#include <cooperative_groups.h>

namespace aar {

constexpr int CH_NB = 32;
constexpr int CH_THREADS = 256;
constexpr int CH_LD = CH_NB + 1;

__global__ void __launch_bounds__(CH_THREADS) k_reduced_solve(int n, double *__restrict__ S, const double *__restrict__ b, double *__restrict__ x, const LmState *__restrict__ st,
                                                              int *__restrict__ chol_fail, double *__restrict__ Xinv /* [nblk][32][32] inverses of the diagonal factors */) {
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
                // X = L_kk^-1 (lower triangular): thread c < 32 solves L X[:, c] = e_c by forward substitution
                if (tid < CH_NB) {
                    const int c = tid;
                    for (int r = 0; r < CH_NB; r++) {
                        double v = r == c ? 1.0 : 0.0;
                        for (int q = c; q < r; q++) v = fma(-sC[r * CH_LD + q], sB[q * CH_LD + c], v);
                        sB[r * CH_LD + c] = r < c ? 0.0 : v / sC[r * CH_LD + r];
                    }
                }
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) Xinv[(size_t)k * CH_NB * CH_NB + e] = sB[(e / CH_NB) * CH_LD + e % CH_NB];
            }
            // write the tile back (A_ik updated, or L_kk: lower triangle only)
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                const int r = e / CH_NB, c = e % CH_NB, gr = i * CH_NB + r, gc = k * CH_NB + c;
                if (gr < n && gc < n && (i != k || c <= r)) S[(size_t)gr * n + gc] = sC[r * CH_LD + c];
            }
        }
        grid.sync();
        // ---- L_ik = A_ik X_k^T for i > k (X_k = L_kk^-1): L_ik[r][c] = sum_q A_ik[r][q] X_k[c][q]
        {
            bool loaded = false;
            for (int i = k + 1 + blockIdx.x; i < nblk; i += gridDim.x) {
                __syncthreads();
                if (!loaded) { for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sB[(e / CH_NB) * CH_LD + e % CH_NB] = Xinv[(size_t)k * CH_NB * CH_NB + e]; loaded = true; }
                load_tile(sC, i, k);
                __syncthreads();
                double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll 8
                for (int q = 0; q < CH_NB; q++) {
                    const double a0 = sC[ty * CH_LD + q], a1 = sC[(ty + 16) * CH_LD + q], b0 = sB[tx * CH_LD + q], b1 = sB[(tx + 16) * CH_LD + q];
                    c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
                }
                const int gr0 = i * CH_NB + ty, gr1 = gr0 + 16, gc0 = k * CH_NB + tx, gc1 = gc0 + 16;
                if (gr0 < n && gc0 < n) S[(size_t)gr0 * n + gc0] = c00;
                if (gr0 < n && gc1 < n) S[(size_t)gr0 * n + gc1] = c01;
                if (gr1 < n && gc0 < n) S[(size_t)gr1 * n + gc0] = c10;
                if (gr1 < n && gc1 < n) S[(size_t)gr1 * n + gc1] = c11;
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
        load_tile(sC, kb, kb);          // the diagonal factor, staged once (the solve below is latency bound)
        __syncthreads();
        if (warp == 0) {   // 32x32 triangular solve, lane r owns row r
            double v = lane < nr ? sv[lane] : 0.0;
            for (int c = 0; c < nr; c++) {
                const double xc = __shfl_sync(0xffffffffu, v, c) / sC[c * CH_LD + c];
                if (lane == c) v = xc;
                else if (lane > c && lane < nr) v = fma(-sC[lane * CH_LD + c], xc, v);
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
        load_tile(sC, kb, kb);
        __syncthreads();
        if (warp == 0) {
            double v = lane < nr ? sv[lane] : 0.0;
            for (int c = nr - 1; c >= 0; c--) {
                const double xc = __shfl_sync(0xffffffffu, v, c) / sC[c * CH_LD + c];
                if (lane == c) v = xc;
                else if (lane < c) v = fma(-sC[c * CH_LD + lane], xc, v);
            }
            if (lane < nr) x[r0 + lane] = v;
        }
        __syncthreads();
    }
}

} // namespace aar
