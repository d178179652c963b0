// FP64 dependent-chain latency and issue throughput on sm_100a (one warp / several warps per scheduler).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o dp_latency dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) x[i] = x[i] + b;
                else if (OP == 1) x[i] = x[i] * b;
                else x[i] = x[i] + x[i] * b;      // DMUL + DADD without contraction
            }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP> void run(const char* name, int threads) {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    chain<ILP, OP><<<1, threads>>>(out, cyc, 1.0, 1.0000001, iters);
    chain<ILP, OP><<<1, threads>>>(out, cyc, 1.0, 1.0000001, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 16 * ILP * (OP == 2 ? 2 : 1);
    printf("%-10s ILP %2d threads %4d (warps/scheduler %.1f): %.2f cycles per dependent step, %.2f cycles per warp-instruction per scheduler\n", name, ILP, threads,
           threads / 128.0, (double)c / (iters * 16), (double)c / (ops * (threads / 32.0) / 4.0 > 0 ? ops * (threads / 32.0 < 4 ? 1 : threads / 128.0) : 1));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>("DADD", 32); run<1, 1>("DMUL", 32); run<1, 2>("DMUL+DADD", 32);
    run<2, 0>("DADD", 32); run<4, 0>("DADD", 32); run<8, 0>("DADD", 32);
    run<1, 0>("DADD", 128); run<1, 0>("DADD", 512); run<4, 0>("DADD", 512); run<8, 2>("DMUL+DADD", 512); run<8, 0>("DADD", 512);
    return 0;
}
