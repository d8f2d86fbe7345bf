// Development aid: FP64 dependent-issue latency and throughput vs ILP / warps per SM on the GPU in front of us.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            else if (OP == 1) x[i] = __dadd_rn(x[i], b);
            else x[i] = __dmul_rn(x[i], a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP> void run(const char *name, int threads, double *out, long long *cyc) {
    const int iters = 4096;
    k<ILP, OP><<<148, threads>>>(out, cyc, 1.0000001, 1e-9, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_instr = (double)h / (iters * ILP);              // cycles between instruction issues of one warp
    const double sm_rate = (threads / 32.0) * iters * ILP / (double)h;      // warp-instructions per cycle per SM
    printf("%s ILP=%d warps/SM=%2d: %.2f cycles/instr/warp, %.3f warp-instr/cycle/SM (peak 2.0)\n", name, ILP, threads / 32, per_warp_instr, sm_rate);
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int threads : {32, 128, 256, 384, 512, 1024}) {
        if (threads == 32) { run<1, 0>("DFMA", threads, out, cyc); run<1, 1>("DADD", threads, out, cyc); run<1, 2>("DMUL", threads, out, cyc); run<2, 0>("DFMA", threads, out, cyc); run<4, 0>("DFMA", threads, out, cyc); run<8, 0>("DFMA", threads, out, cyc); }
        else { run<1, 0>("DFMA", threads, out, cyc); run<2, 0>("DFMA", threads, out, cyc); run<4, 0>("DFMA", threads, out, cyc); run<2, 2>("DMUL", threads, out, cyc); }
    }
    return 0;
}
