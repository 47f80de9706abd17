// DFMA dependent-issue latency and LDS latency on one warp (sm_100a).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double *out, long long *cyc, double a, double b, int iters)
{
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_chain(int *out, long long *cyc, int iters)
{
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 32) & 1023;
    __syncthreads();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) j = nxt[j];
    }
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
#define RUN(C, T)                                                                                       \
    dfma_chain<C><<<1, T>>>(out, cyc, 1.0000001, 1e-9, iters);                                          \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                     \
    printf("DFMA chains/thread %d, threads %4d: %.2f cycles per DFMA per chain step (%.2f cycles per warp-DFMA)\n", C, T, \
           (double)h / (iters * 16.0), (double)h / (iters * 16.0 * C * ((T + 127) / 128)));
    RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(6, 32) RUN(8, 32) RUN(1, 128) RUN(1, 256) RUN(3, 256) RUN(1, 512) RUN(4, 512)
    lds_chain<<<1, 32>>>((int *)out, cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.32 dependent latency: %.2f cycles\n", (double)h / (iters * 16.0));
    return 0;
}
