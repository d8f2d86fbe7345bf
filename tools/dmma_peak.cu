// Development aid: throughput of the FP64 tensor-core instruction (mma.sync.aligned.m8n8k4.f64) on this GPU, next to
// the DFMA figure of csrc/fp64_peak.cu — decides whether the Schur SYRK (a real 468 x 468 x 600k contraction at
// BASELINE cfg 4) should move from CUDA-core FMAs to DMMA.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> void run(int ctas_per_sm) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm, iters = 20000;
    double *d; cudaMalloc(&d, sizeof(double) * grid * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dmma<NACC><<<grid, 256>>>(d, 100, 1.0, 1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dmma<NACC><<<grid, 256>>>(d, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)grid * 8 * iters * NACC * 512.0;
    printf("{\"instr\": \"mma.sync.m8n8k4.f64\", \"independent_accumulators\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", NACC, 8 * ctas_per_sm, ms, flop / ms * 1e-9);
    cudaFree(d);
}
// 1 or 2 accumulators per warp at 8 warps/SM (2 per scheduler): the rate then measures the dependent-issue latency of the
// instruction — what a kernel pays when consecutive observations chain on the same accumulator (profiles/r1_notes.md)
int main() { run<1>(1); run<2>(1); run<4>(1); run<8>(1); run<8>(2); run<16>(2); run<16>(4); return 0; }
