// fp64_peak.cu — microbenchmarks that size the roofline of the hot path on the B200 in front of us.
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 figures only; the Jacobian kernel is bound by the
// FP64 pipe, so its denominator is measured here: DFMA issue rate, DADD/DMUL (the -fmad=false share of the
// kernel), FP64 mma.sync (DMMA), conversions and divisions, FP64 reductions to L2 and to shared memory.
// Prints one JSON object.  nvcc -gencode arch=compute_100a,code=sm_100a fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double *out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_dmuladd(double *out, double a, double b) { // alternating DMUL / DADD, no contraction
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS / 2; i++) {
        x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a);
        x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a);
        x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);
        x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double *out, double a, double b) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    for (int i = 0; i < ITERS; i++) { dmma(c0, c1, a, b); dmma(c2, c3, a, b); dmma(c4, c5, a, b); dmma(c6, c7, a, b); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
__global__ void k_dmma_dfma(double *out, double a, double b) { // do the two pipes overlap?
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; i++) {
        dmma(c0, c1, a, b); dmma(c2, c3, a, b); dmma(c4, c5, a, b); dmma(c6, c7, a, b);
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7 + x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_ddiv(double *out, double a) {
    double x0 = threadIdx.x + 1.5, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < ITERS / 8; i++) { x0 = a / x0 + 1.25; x1 = a / x1 + 1.25; x2 = a / x2 + 1.25; x3 = a / x3 + 1.25; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void k_cvt(double *out, double a) { // f64 -> f32 -> f64 round trips
    double x0 = threadIdx.x + 1.5, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < ITERS / 4; i++) {
        x0 = (double)(float)x0 + a; x1 = (double)(float)x1 + a; x2 = (double)(float)x2 + a; x3 = (double)(float)x3 + a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void k_red(double *target, int spread, double v) { // fire-and-forget FP64 reductions to L2
    size_t idx = spread ? ((size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 37) % (size_t)spread : 0;
    for (int i = 0; i < 256; i++) { atomicAdd(target + idx, v); if (spread) idx = (idx + 4099) % (size_t)spread; }
}
__global__ void k_atoms(double *out, double v) { // shared-memory FP64 atomic adds, spread addresses
    __shared__ double s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = 0;
    __syncthreads();
    int idx = (threadIdx.x * 7) & 2047;
    for (int i = 0; i < 256; i++) { atomicAdd(&s[idx], v); idx = (idx + 33) & 2047; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}

template <typename F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256;
    double *out; CK(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
    double *tgt; const int SPREAD = 1 << 20; CK(cudaMalloc(&tgt, sizeof(double) * SPREAD)); CK(cudaMemset(tgt, 0, sizeof(double) * SPREAD));
    const double nthreads = (double)blocks * threads;
    float t;
    t = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double dfma_tflops = nthreads * ITERS * 8 * 2 / (t * 1e-3) / 1e12;
    t = time_ms([&] { k_dmuladd<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double dmuladd_tops = nthreads * ITERS * 8 / (t * 1e-3) / 1e12;    // one flop per instruction
    t = time_ms([&] { k_dmma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double dmma_tflops = (nthreads / 32) * ITERS * 4 * (8 * 8 * 4 * 2) / (t * 1e-3) / 1e12;
    t = time_ms([&] { k_dmma_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double both_tflops = ((nthreads / 32) * ITERS * 4 * (8 * 8 * 4 * 2) + nthreads * ITERS * 8 * 2) / (t * 1e-3) / 1e12;
    t = time_ms([&] { k_ddiv<<<blocks, threads>>>(out, 3.3); });
    const double ddiv_g = nthreads * (ITERS / 8) * 4 / (t * 1e-3) / 1e9;
    t = time_ms([&] { k_cvt<<<blocks, threads>>>(out, 0.125); });
    const double cvt_g = nthreads * (ITERS / 4) * 4 * 2 / (t * 1e-3) / 1e9;
    t = time_ms([&] { k_red<<<blocks, threads>>>(tgt, SPREAD, 1.0); });
    const double red_spread_g = nthreads * 256 / (t * 1e-3) / 1e9;
    t = time_ms([&] { k_red<<<blocks, threads>>>(tgt, 4096, 1.0); });
    const double red_4k_g = nthreads * 256 / (t * 1e-3) / 1e9;
    t = time_ms([&] { k_atoms<<<blocks, threads>>>(out, 1.0); });
    const double atoms_g = nthreads * 256 / (t * 1e-3) / 1e9;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, \"dfma_tflops\": %.2f, \"dmul_dadd_tops\": %.2f, \"dmma_tflops\": %.2f, "
           "\"dmma_plus_dfma_tflops\": %.2f, \"ddiv_gops\": %.1f, \"cvt_f64_f32_gops\": %.1f, \"red_f64_spread_gops\": %.1f, "
           "\"red_f64_4k_addr_gops\": %.1f, \"atoms_f64_gops\": %.1f}\n",
           prop.name, sms, clk / 1000.0, dfma_tflops, dmuladd_tops, dmma_tflops, both_tflops, ddiv_g, cvt_g, red_spread_g, red_4k_g, atoms_g);
    return 0;
}
