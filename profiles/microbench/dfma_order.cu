// Micro-benchmark 3: the Clenshaw step written so that consecutive DFMAs share an operand in the same
// source slot (operand-reuse cache) and no DFMA reads more than two fresh register operands.
//   ORDER 0: natural order (per point: tn, n0r, n0i, n1r, n1i)          -- what the first kernel used
//   ORDER 1: C++ statements in reuse order (T block; per point n0r, n1r, n1i, n0i)
//   ORDER 2: same order, every fma an `asm volatile` (pins the PTX order and operand slots)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_order dfma_order.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double vfma(double a, double b, double c)
{
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}

template <int P, int ORDER>
__device__ __forceinline__ void step(const double4 t, const double2 c, const double (&B)[P], double (&y0r)[P],
                                     double (&y0i)[P], double (&y1r)[P], double (&y1i)[P])
{
    if (ORDER == 0) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double tn = fma(B[p], t.z, t.y);
            const double n0r = fma(t.x, y1r[p], c.x);
            const double n0i = fma(t.x, y1i[p], c.y);
            const double n1r = fma(tn, y1r[p], y0r[p]);
            const double n1i = fma(tn, y1i[p], y0i[p]);
            y0r[p] = n0r; y0i[p] = n0i; y1r[p] = n1r; y1i[p] = n1i;
        }
    } else if (ORDER == 1) {
        double tn[P];
#pragma unroll
        for (int p = 0; p < P; ++p) tn[p] = fma(B[p], t.z, t.y);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double n0r = fma(t.x, y1r[p], c.x);
            const double n1r = fma(tn[p], y1r[p], y0r[p]);
            const double n1i = fma(tn[p], y1i[p], y0i[p]);
            const double n0i = fma(t.x, y1i[p], c.y);
            y0r[p] = n0r; y0i[p] = n0i; y1r[p] = n1r; y1i[p] = n1i;
        }
    } else if (ORDER == 3) {
        // NOT the real recurrence: the third operand of the y1 update is a shared coefficient instead of y0,
        // so that no DFMA reads more than two per-thread registers.  Same instruction count and mix;
        // tells how fast the loop would run if the three-register DFMAs were gone.
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double tn = fma(B[p], t.z, t.y);
            const double n0r = fma(t.x, y0r[p], c.x);
            const double n0i = fma(t.x, y0i[p], c.y);
            const double n1r = fma(tn, y1r[p], t.w);
            const double n1i = fma(tn, y1i[p], t.w);
            y0r[p] = n0r; y0i[p] = n0i; y1r[p] = n1r; y1i[p] = n1i;
        }
    } else {
        double tn[P];
#pragma unroll
        for (int p = 0; p < P; ++p) tn[p] = vfma(B[p], t.z, t.y);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double n0r = vfma(t.x, y1r[p], c.x);
            const double n1r = vfma(tn[p], y1r[p], y0r[p]);
            const double n1i = vfma(tn[p], y1i[p], y0i[p]);
            const double n0i = vfma(t.x, y1i[p], c.y);
            y0r[p] = n0r; y0i[p] = n0i; y1r[p] = n1r; y1i[p] = n1i;
        }
    }
}

template <int P, int ORDER, int UNROLL>
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
    for (int it = 0; it < iters; ++it) {
#pragma unroll UNROLL
        for (int r = 0; r < 256; ++r) step<P, ORDER>(tab[r], cs[r], B, y0r, y0i, y1r, y1i);
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

template <int P, int ORDER, int UNROLL>
void run(int sms, double *out, const double *in)
{
    for (int bps : {1, 2}) {
        const int grid = sms * bps, it = 400;
        const double dfma = (double)grid * 256 * it * 256 * 5 * P;
        float t = timeit([&] { kstep<P, ORDER, UNROLL><<<grid, 256>>>(out, in, it); });
        printf("P=%d order=%d unroll=%d CTAs/SM=%d : %.2f TFLOP/s\n", P, ORDER, UNROLL, bps, 2 * dfma / t / 1e9);
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
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<4, 0, 4>(sms, out, in);
    run<4, 1, 4>(sms, out, in);
    run<4, 2, 4>(sms, out, in);
    run<4, 2, 2>(sms, out, in);
    run<8, 0, 2>(sms, out, in);
    run<8, 1, 2>(sms, out, in);
    run<8, 2, 2>(sms, out, in);
    run<6, 2, 2>(sms, out, in);
    run<4, 3, 4>(sms, out, in);
    run<8, 3, 2>(sms, out, in);
    return 0;
}
