// Development aid (round 2): what does it cost to scatter-add 36-double blocks into a small hot table?
// Decides how the camera x marker blocks of J^T J (945 blocks x 36 doubles = 272 KB at BASELINE cfg 4, one block per
// marker observation, 25.6 M observations per Jacobian evaluation) leave the assembly kernel:
//   red_lane     every lane adds its own block: 36 RED instructions per warp, 32 different blocks each (uncoalesced)
//   red_coal     the C fragment of an FP64 mma (lane (g, q) holds row g, columns 2q, 2q+1 of ONE block): 2 RED instructions per
//                block, consecutive addresses inside a row
//   replicas     the table replicated R times (blockIdx % R) to spread same-address traffic over more L2 sectors
//   smem_cas     the same per-lane adds into a shared-memory table (CAS loop), for comparison
// nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o red_bench tools/red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NBLK = 945, BLK = 36;

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(256) k_red_lane(double *tab, int replicas, int per_thread) {
    double *t = tab + (size_t)(blockIdx.x % replicas) * NBLK * BLK;
    unsigned s = hash32(blockIdx.x * 256 + threadIdx.x);
    for (int i = 0; i < per_thread; i++) {
        s = hash32(s + i);
        double *d = t + (size_t)(s % NBLK) * BLK;
#pragma unroll
        for (int v = 0; v < BLK; v++) atomicAdd(d + v, 1.0);
    }
}
// one block per WARP and step, emitted from the mma C-fragment layout
__global__ void __launch_bounds__(256) k_red_coal(double *tab, int replicas, int per_warp) {
    double *t = tab + (size_t)(blockIdx.x % replicas) * NBLK * BLK;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    unsigned s = hash32(blockIdx.x * 8 + (threadIdx.x >> 5));
    for (int i = 0; i < per_warp; i++) {
        s = hash32(s + i);
        double *d = t + (size_t)(s % NBLK) * BLK;
        if (g < 6 && q < 3) { atomicAdd(d + g * 6 + 2 * q, 1.0); atomicAdd(d + g * 6 + 2 * q + 1, 1.0); }
    }
}
__global__ void __launch_bounds__(256) k_smem_cas(double *out, int per_thread) {
    extern __shared__ double st[];
    for (int i = threadIdx.x; i < 600 * 37; i += 256) st[i] = 0;
    __syncthreads();
    unsigned s = hash32(blockIdx.x * 256 + threadIdx.x);
    for (int i = 0; i < per_thread; i++) {
        s = hash32(s + i);
        double *d = st + (size_t)(s % 600) * 37;
#pragma unroll
        for (int v = 0; v < BLK; v++) atomicAdd(d + v, 1.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = st[5];
}

template <typename F> float best_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *tab; cudaMalloc(&tab, sizeof(double) * NBLK * BLK * 64); cudaMemset(tab, 0, sizeof(double) * NBLK * BLK * 64);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8);
    const int grid = sms * 4;
    for (int rep : {1, 4, 16, 64}) {
        const int per_thread = 64;
        float t = best_ms([&] { k_red_lane<<<grid, 256>>>(tab, rep, per_thread); });
        const double blocks = (double)grid * 256 * per_thread;
        printf("{\"kernel\": \"red_lane\", \"replicas\": %d, \"ms\": %.3f, \"Gblocks_per_s\": %.3f, \"Gadds_per_s\": %.1f}\n", rep, t, blocks / t * 1e-6, blocks * BLK / t * 1e-6);
    }
    for (int rep : {1, 4, 16, 64}) {
        const int per_warp = 2048;
        float t = best_ms([&] { k_red_coal<<<grid, 256>>>(tab, rep, per_warp); });
        const double blocks = (double)grid * 8 * per_warp;
        printf("{\"kernel\": \"red_coal\", \"replicas\": %d, \"ms\": %.3f, \"Gblocks_per_s\": %.3f, \"Gadds_per_s\": %.1f}\n", rep, t, blocks / t * 1e-6, blocks * BLK / t * 1e-6);
    }
    {
        cudaFuncSetAttribute(k_smem_cas, cudaFuncAttributeMaxDynamicSharedMemorySize, 600 * 37 * 8);
        const int per_thread = 64;
        float t = best_ms([&] { k_smem_cas<<<sms, 256, 600 * 37 * 8>>>(out, per_thread); });
        const double blocks = (double)sms * 256 * per_thread;
        printf("{\"kernel\": \"smem_cas\", \"ms\": %.3f, \"Gblocks_per_s\": %.3f, \"Gadds_per_s\": %.1f}\n", t, blocks / t * 1e-6, blocks * BLK / t * 1e-6);
    }
    return 0;
}
