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


// ------------------------------------------------------------------------------------------------
// Cluster variant for n <= 512 (every BASELINE configuration: n_r = 468 at cfg 4): ONE thread-block cluster of
// nblk <= 16 CTAs, CTA i owns block row i of the matrix in its shared memory for the whole solve, tiles of other
// block rows are read through distributed shared memory and the 2 synchronisations per block column are hardware
// cluster barriers instead of grid-wide barriers through global memory.  Nothing but the initial load and the
// final x touches global memory.  Same left-looking algorithm and same inverse-diagonal-block trick as above.
namespace cg = cooperative_groups;

__device__ __forceinline__ void tile_product_2x2(const double *A, int lda, const double *B, int ldb, int ty, int tx, double &c00, double &c01, double &c10, double &c11) {
#pragma unroll 8
    for (int q = 0; q < CH_NB; q++) {
        const double a0 = A[ty * lda + q], a1 = A[(ty + 16) * lda + q], b0 = B[tx * ldb + q], b1 = B[(tx + 16) * ldb + q];
        c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
    }
}

__global__ void __launch_bounds__(CH_THREADS) k_reduced_solve_cluster(int n, const double *__restrict__ S, const double *__restrict__ b, double *__restrict__ x, const LmState *__restrict__ st,
                                                                      int *__restrict__ chol_fail) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    const int nblk = (n + CH_NB - 1) / CH_NB, LD = nblk * CH_NB + 2;
    double *sRow = sm;                                   // [32][LD]   block row `me` of the matrix / of L
    double *sT = sRow + CH_NB * LD;                      // [32][33]   staging of a remote tile
    double *sX = sT + CH_NB * CH_LD;                     // [32][33]   inverse of this CTA's diagonal factor
    double *sXr = sX + CH_NB * CH_LD;                    // [32][33]   staging of a remote inverse
    double *sy = sXr + CH_NB * CH_LD;                    // [nblk*32]  right-hand side / y / x, replicated in every CTA
    double *ss = sy + nblk * CH_NB;                      // [32]       partial sums of the back substitution
    __shared__ double sv[CH_NB + 1];
    const int me = (int)cluster.block_rank(), tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const double mu = st->mu;
    // ---- load block row `me` (lower part from the UPPER triangle of S), mu on the diagonal, identity padding
    for (int e = tid; e < CH_NB * nblk * CH_NB; e += CH_THREADS) {
        const int r = e / (nblk * CH_NB), c = e % (nblk * CH_NB), gr = me * CH_NB + r;
        double v = 0.0;
        if (gr < n && c <= gr) v = S[(size_t)c * n + gr] + (c == gr ? mu : 0.0);
        else if (gr >= n && c == gr) v = 1.0;
        sRow[r * LD + c] = v;
    }
    for (int e = tid; e < nblk * CH_NB; e += CH_THREADS) sy[e] = e < n ? b[e] : 0.0;
    if (tid < CH_NB) ss[tid] = 0.0;
    cluster.sync();
    for (int k = 0; k <= me; k++) {
        // ---- own tile (me, k) -= sum_{j<k} L(me, j) L(k, j)^T ; L(k, j) from CTA k through distributed shared memory
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        const double *rowk = cluster.map_shared_rank(sRow, k);
        for (int j = 0; j < k; j++) {
            const double *B;
            int ldb;
            if (k == me) { B = sRow + j * CH_NB; ldb = LD; }
            else {
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sT[(e / CH_NB) * CH_LD + e % CH_NB] = rowk[(e / CH_NB) * LD + j * CH_NB + e % CH_NB];
                __syncthreads();
                B = sT; ldb = CH_LD;
            }
            tile_product_2x2(sRow + j * CH_NB, LD, B, ldb, ty, tx, c00, c01, c10, c11);
        }
        __syncthreads();
        double *C = sRow + k * CH_NB;
        C[ty * LD + tx] -= c00; C[ty * LD + tx + 16] -= c01; C[(ty + 16) * LD + tx] -= c10; C[(ty + 16) * LD + tx + 16] -= c11;
        __syncthreads();
        if (k == me) {
            // ---- diagonal block: L_kk = chol(C) in place, then X = L_kk^-1
            for (int j = 0; j < CH_NB; j++) {
                if (tid == 0) { double d = C[j * LD + j]; if (!(d > 0)) { atomicExch(chol_fail, 1); d = 1; } sv[0] = sqrt(d); }
                __syncthreads();
                const double dj = sv[0];
                if (tid > j && tid < CH_NB) C[tid * LD + j] /= dj;
                if (tid == j) C[j * LD + j] = dj;
                __syncthreads();
                for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
                    const int r = e / CH_NB, c = e % CH_NB;
                    if (c > j && r >= c) C[r * LD + c] = fma(-C[r * LD + j], C[c * LD + j], C[r * LD + c]);
                }
                __syncthreads();
            }
            if (warp == 0) {      // lane c: column c of X by forward substitution, X column kept in registers
                double xc[CH_NB];
#pragma unroll
                for (int r = 0; r < CH_NB; r++) {
                    double v = r == lane ? 1.0 : 0.0;
#pragma unroll
                    for (int q = 0; q < CH_NB; q++) if (q < r) v = fma(-C[r * LD + q], xc[q], v);
                    xc[r] = r < lane ? 0.0 : v / C[r * LD + r];
                }
#pragma unroll
                for (int r = 0; r < CH_NB; r++) sX[r * CH_LD + lane] = xc[r];
            }
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) { const int r = e / CH_NB, c = e % CH_NB; if (c > r) C[r * LD + c] = 0.0; }
        }
        cluster.sync();                                  // L_kk and X_k are visible to the cluster
        if (k < me) {
            // ---- L(me, k) = C X_k^T
            const double *Xk = cluster.map_shared_rank(sX, k);
            for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) sXr[(e / CH_NB) * CH_LD + e % CH_NB] = Xk[(e / CH_NB) * CH_LD + e % CH_NB];
            __syncthreads();
            double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
            tile_product_2x2(C, LD, sXr, CH_LD, ty, tx, d00, d01, d10, d11);
            __syncthreads();
            C[ty * LD + tx] = d00; C[ty * LD + tx + 16] = d01; C[(ty + 16) * LD + tx] = d10; C[(ty + 16) * LD + tx + 16] = d11;
        }
        cluster.sync();                                  // block column k of L is final
    }
    // CTAs with me < k idle through the remaining steps but must take part in the barriers
    for (int k = me + 1; k < nblk; k++) { cluster.sync(); cluster.sync(); }
    // ---- forward substitution  L y = b : CTA k owns block row k, y_k is broadcast into every CTA's sy
    for (int k = 0; k < nblk; k++) {
        if (k == me) {
            // t = b_k - sum_{c < 32k} L(k, c) y_c   (rows by warps, dot products by lanes)
            for (int r = warp; r < CH_NB; r += CH_THREADS / 32) {
                double s = 0;
                for (int c = lane; c < k * CH_NB; c += 32) s = fma(sRow[r * LD + c], sy[c], s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) sv[r] = sy[k * CH_NB + r] - s;
            }
            __syncthreads();
            if (tid < CH_NB) {    // y_k = X_k t
                double s = 0;
                for (int q = 0; q <= tid; q++) s = fma(sX[tid * CH_LD + q], sv[q], s);
                for (int rk = 0; rk < nblk; rk++) cluster.map_shared_rank(sy, rk)[k * CH_NB + tid] = s;
            }
        }
        cluster.sync();
    }
    // ---- back substitution  L^T x = y : x_i = X_i^T (y_i - s_i), then every CTA k < i adds L(i, k)^T x_i to its s_k
    for (int i = nblk - 1; i >= 0; i--) {
        if (i == me) {
            if (tid < CH_NB) sv[tid] = sy[i * CH_NB + tid] - ss[tid];
            __syncthreads();
            if (tid < CH_NB) {
                double s = 0;
                for (int q = tid; q < CH_NB; q++) s = fma(sX[q * CH_LD + tid], sv[q], s);
                for (int rk = 0; rk < nblk; rk++) cluster.map_shared_rank(sy, rk)[i * CH_NB + tid] = s;      // x_i overwrites y_i
                if (i * CH_NB + tid < n) x[i * CH_NB + tid] = s;
            }
        }
        cluster.sync();
        if (me < i) {
            const double *rowi = cluster.map_shared_rank(sRow, i);       // tile (i, me) lives in CTA i
            if (tid < CH_NB) {
                double s = 0;
                for (int q = 0; q < CH_NB; q++) s = fma(rowi[q * LD + me * CH_NB + tid], sy[i * CH_NB + q], s);
                ss[tid] += s;
            }
        }
        __syncthreads();
    }
    cluster.sync();                                      // no CTA may exit while its shared memory can still be read remotely
}

} // namespace aar
