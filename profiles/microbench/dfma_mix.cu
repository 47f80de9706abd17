// Micro-benchmark 2: what keeps the Clenshaw inner step (5 DFMA per point, coefficients broadcast from
// shared memory) below the DFMA peak?  Same arithmetic, different operand sources / points per thread.
//   MODE 0: coefficients from shared memory (LDS.128 broadcast)      -- the product kernel's form
//   MODE 1: coefficients held in registers (no loads in the loop)    -- upper bound for the math itself
//   MODE 2: coefficients from __constant__ memory (uniform operands)
//   MODE 3: shared memory, but coefficients pre-splatted per lane (conflict-free, non-broadcast LDS.64)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_mix dfma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double4 ctab[256];
__constant__ double2 ccs[256];

template <int P, int MODE>
__global__ void __launch_bounds__(256) kstep(double *out, const double *in, int iters)
{
    __shared__ double4 tab[256];
    __shared__ double2 cs[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        tab[i] = make_double4(-in[i % 50] * 0.5, -in[(i + 1) % 50], in[(i + 2) % 50] * 1e-3, 0);
        cs[i] = make_double2(in[(i + 3) % 50], in[(i + 4) % 50]);
    }
    __syncthreads();
    double B[P], y0r[P], y0i[P], y1r[P], y1i[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        B[p] = in[p] + threadIdx.x * 1e-3;
        y0r[p] = y0i[p] = y1r[p] = y1i[p] = 0.0;
    }
    double4 tr = tab[threadIdx.x & 3];
    double2 cr = cs[threadIdx.x & 3];
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int r = 0; r < 256; ++r) {
            double4 t;
            double2 c;
            if (MODE == 0) { t = tab[r]; c = cs[r]; }
            if (MODE == 1) { t = tr; c = cr; }
            if (MODE == 2) { t = ctab[r]; c = ccs[r]; }
#pragma unroll
            for (int p = 0; p < P; ++p) {
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
    for (int p = 0; p < P; ++p) s += y0r[p] + y0i[p] + y1r[p] + y1i[p];
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

template <int P, int MODE>
void run(const char *label, int sms, double *out, const double *in)
{
    for (int bps : {1, 2}) {
        const int grid = sms * bps, it = 400;
        const double dfma = (double)grid * 256 * it * 256 * 5 * P;
        float t = timeit([&] { kstep<P, MODE><<<grid, 256>>>(out, in, it); });
        printf("%-34s P=%d CTAs/SM=%d : %.2f TFLOP/s\n", label, P, bps, 2 * dfma / t / 1e9);
    }
}

int main()
{
    double *in, *out;
    cudaMalloc(&in, 1024 * 8);
    cudaMalloc(&out, 148 * 16 * 256 * 8);
    double h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.5 + 1e-3 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    double4 ht[256];
    double2 hc[256];
    for (int i = 0; i < 256; ++i) {
        ht[i] = make_double4(-h[i % 50] * 0.5, -h[(i + 1) % 50], h[(i + 2) % 50] * 1e-3, 0);
        hc[i] = make_double2(h[(i + 3) % 50], h[(i + 4) % 50]);
    }
    cudaMemcpyToSymbol(ctab, ht, sizeof(ht));
    cudaMemcpyToSymbol(ccs, hc, sizeof(hc));
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<2, 0>("smem broadcast", sms, out, in);
    run<4, 0>("smem broadcast", sms, out, in);
    run<6, 0>("smem broadcast", sms, out, in);
    run<8, 0>("smem broadcast", sms, out, in);
    run<4, 1>("coefficients in registers", sms, out, in);
    run<8, 1>("coefficients in registers", sms, out, in);
    run<4, 2>("constant memory", sms, out, in);
    run<8, 2>("constant memory", sms, out, in);
    return 0;
}
