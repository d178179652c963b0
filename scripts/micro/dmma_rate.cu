// FP64 tensor-core (mma.sync.m8n8k4.f64, SASS DMMA) latency and throughput on sm_100a, next to the DFMA / DMUL+DADD rate of the
// CUDA-core FP64 pipe.  Answers the north-star question "does FP64 MMA beat the CUDA-core path for the (N+1)-term contractions".
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ILP independent accumulator pairs per warp, each a dependent chain of DMMAs
template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, double a, double b, int iters) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, bool FMA>
__global__ void k_dfma(double* out, long long* cyc, double a, double b, int iters) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) c[i] = FMA ? __fma_rn(c[i], a, b) : __dadd_rn(__dmul_rn(c[i], a), b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F> double timeMs(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8 * 2); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int iters = 4000;
    long long c;
    printf("device %s, %d SMs\n", p.name, sms);
    // latency: one warp, one chain
    k_dmma<1><<<1, 32>>>(out, cyc, 1.0, 1e-9, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 dependent-chain latency: %.1f cycles\n", (double)c / (iters * 8));
    k_dfma<1, true><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent-chain latency: %.1f cycles\n", (double)c / (iters * 8));
#define RUN_DMMA(ILP, THREADS)                                                                                                    \
    {                                                                                                                             \
        k_dmma<ILP><<<1, THREADS>>>(out, cyc, 1.0, 1e-9, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                   \
        const double perSM = (double)iters * 8 * ILP * (THREADS / 32) / (double)c;                                               \
        const double ms = timeMs([&] { k_dmma<ILP><<<sms, THREADS>>>(out, cyc, 1.0, 1e-9, iters); });                             \
        const double tf = (double)iters * 8 * ILP * (THREADS / 32) * sms * 512.0 / (ms * 1e-3) / 1e12;                            \
        printf("DMMA ILP %d warps/SM %2d: %.3f DMMA/cycle/SM (%.1f FMA/cycle/SM), whole chip %.1f TFLOP/s\n", ILP, THREADS / 32, perSM, perSM * 256, tf); \
    }
    RUN_DMMA(1, 128) RUN_DMMA(2, 128) RUN_DMMA(4, 128) RUN_DMMA(8, 128) RUN_DMMA(1, 256) RUN_DMMA(4, 256) RUN_DMMA(1, 512) RUN_DMMA(2, 512) RUN_DMMA(4, 512) RUN_DMMA(8, 512)
#define RUN_DFMA(ILP, THREADS, FMA)                                                                                               \
    {                                                                                                                             \
        k_dfma<ILP, FMA><<<1, THREADS>>>(out, cyc, 1.0000001, 1e-9, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);       \
        const double perSM = (double)iters * 8 * ILP * THREADS / (double)c;                                                       \
        const double ms = timeMs([&] { k_dfma<ILP, FMA><<<sms, THREADS>>>(out, cyc, 1.0000001, 1e-9, iters); });                  \
        const double tf = (double)iters * 8 * ILP * THREADS * sms * 2.0 / (ms * 1e-3) / 1e12;                                     \
        printf("%s ILP %d warps/SM %2d: %.1f MAC/cycle/SM, whole chip %.1f TFLOP/s\n", FMA ? "DFMA     " : "DMUL+DADD", ILP, THREADS / 32, perSM, tf); \
    }
    RUN_DFMA(4, 512, true) RUN_DFMA(8, 512, true) RUN_DFMA(8, 1024, true) RUN_DFMA(4, 512, false) RUN_DFMA(8, 512, false) RUN_DFMA(8, 1024, false)
    return 0;
}
