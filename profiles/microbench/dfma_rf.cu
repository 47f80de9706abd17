// Micro-benchmark: does DFMA throughput on sm_100a depend on how many DISTINCT register operands
// each instruction reads?  (Decides how the Clenshaw inner step has to be scheduled.)
//   V0  a = fma(a, m, c)          1 per-accumulator operand, 2 warp-wide constants
//   V1  a = fma(b, c_i, a)        3 distinct register pairs per instruction, nothing shared
//   V2  a = fma(b, m, a)          2 distinct + 1 shared
//   V3  clenshaw-like step on 4 points (the real inner loop body, coefficients in registers)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_rf dfma_rf.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ACC = 16;

template <int V>
__global__ void __launch_bounds__(256) k(double *out, const double *in, int iters)
{
    double a[ACC], b[ACC], c[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        a[i] = in[i] + threadIdx.x * 1e-9;
        b[i] = in[ACC + i];
        c[i] = in[2 * ACC + i];
    }
    const double m = in[100], cc = in[101];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < ACC; ++i) {
                if (V == 0) a[i] = fma(a[i], m, cc);
                if (V == 1) a[i] = fma(b[i], c[i], a[i]);
                if (V == 2) a[i] = fma(b[i], m, a[i]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the real step: 4 points, coefficients (na, ne, sv, cr, ci) change every step (read from smem)
__global__ void __launch_bounds__(256) kstep(double *out, const double *in, int iters)
{
    __shared__ double4 tab[256];
    __shared__ double2 cs[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        tab[i] = make_double4(-in[i % 50] * 0.5, -in[(i + 1) % 50], in[(i + 2) % 50] * 1e-3, 0);
        cs[i] = make_double2(in[(i + 3) % 50], in[(i + 4) % 50]);
    }
    __syncthreads();
    double B[4], y0r[4], y0i[4], y1r[4], y1i[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        B[p] = in[p] + threadIdx.x * 1e-3;
        y0r[p] = y0i[p] = y1r[p] = y1i[p] = 0.0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int r = 0; r < 256; ++r) {
            const double4 t = tab[r];
            const double2 c = cs[r];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const double tn = fma(B[p], t.z, t.y);
                const double n0r = fma(t.x, y1r[p], c.x);
                const double n0i = fma(t.x, y1i[p], c.y);
                const double n1r = fma(tn, y1r[p], y0r[p]);
                const double n1i = fma(tn, y1i[p], y0i[p]);
                y0r[p] = n0r; y0i[p] = n0i; y1r[p] = n1r; y1i[p] = n1i;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) s += y0r[p] + y0i[p] + y1r[p] + y1i[p];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    double *in, *out;
    cudaMalloc(&in, 1024 * 8);
    cudaMalloc(&out, 148 * 16 * 256 * 8);
    double h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.5 + 1e-3 * i;
    h[100] = 0.999999;
    h[101] = 1e-7;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double clk = p.clockRate * 1e3;
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, clk / 1e6);
    for (int bps : {1, 2, 4}) {  // CTAs of 256 threads per SM -> 2, 4, 8 warps per SMSP
        const int grid = p.multiProcessorCount * bps;
        const int iters = 20000;
        const double dfma = (double)grid * 256 * iters * 4 * ACC;
        float t0 = timeit([&] { k<0><<<grid, 256>>>(out, in, iters); });
        float t1 = timeit([&] { k<1><<<grid, 256>>>(out, in, iters); });
        float t2 = timeit([&] { k<2><<<grid, 256>>>(out, in, iters); });
        const int it3 = 400;
        const double dfma3 = (double)grid * 256 * it3 * 256 * 20;
        float t3 = timeit([&] { kstep<<<grid, 256>>>(out, in, it3); });
        printf("CTAs/SM %d:  V0 %.2f TF  V1 %.2f TF  V2 %.2f TF  step %.2f TF   (2 flops per DFMA)\n", bps,
               2 * dfma / t0 / 1e9, 2 * dfma / t1 / 1e9, 2 * dfma / t2 / 1e9, 2 * dfma3 / t3 / 1e9);
    }
    return 0;
}
