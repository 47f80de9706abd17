// method="fft" of qutip.wigner (qutip 5.2.2 qutip/wigner.py: _wigner_fourier, _psi_wigner_fft, _wigner_fft,
// _osc_eigen), reached through src/bosonic_qiskit/wigner.py:190 with method="fft".
//
// QuTiP eigendecomposes rho and, per eigenvector, Fourier-transforms psi(x - s) conj(psi(x + s)) along s with an FFT
// of length n = 2 S over a Toeplitz product, then sums with the eigenvalues.  Every step is linear in
// |psi><psi|, so the sum over eigenvectors is taken FIRST, in closed form, without an eigensolver:
//
//   A[m][j]   = phi_m(x'_j), x' = xvec g / sqrt(2)                 (osc_eigen_kernel; _osc_eigen's recurrence)
//   R[b][a]   = sum_{m,n} A[m][b] rho[m][n] A[n][a]                 (two tiled fp64 products; = sum_k lambda_k psi_k(b) conj(psi_k(a)))
//   F[i][k]   = sum_{d=-m_i}^{m_i} R[i+d][i-d] exp(-2 pi i d k / n), m_i = min(i, S-1-i)      (the Toeplitz band of row i)
//   W[c][i]   = g^2/2 * Re F[i][k_c] / (p_1 - p_0) / n,  k_c = (3n/4 + c) mod n                (the half of the spectrum QuTiP keeps)
//
// The transform is evaluated as a banded DFT: twiddles advance by one complex multiplication per term and are
// re-read from an exact table (one cos/sin pair per residue of d k mod n) every 32 terms.  S^3 terms in all, no
// power-of-two restriction on S, and only the S frequencies QuTiP returns are computed.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "bq_internal.h"

namespace bq {

// A[k][j], k < N, j < S: A0 = exp(-x^2/2) / pi^(1/4), A1 = sqrt(2) x A0, Ak = sqrt(2/k) x A(k-1) - sqrt((k-1)/k) A(k-2)
__global__ void osc_eigen_kernel(double *__restrict__ A, int N, int S, const double *__restrict__ xvec, double g)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    const double x = xvec[j] * g / sqrt(2.0);
    double a0 = exp(-(x * x) / 2.0) / sqrt(sqrt(CUDART_PI));
    A[j] = a0;
    if (N == 1) return;
    double a1 = sqrt(2.0) * x * a0;
    A[(int64_t)S + j] = a1;
    for (int k = 2; k < N; ++k) {
        const double ak = sqrt(2.0 / k) * x * a1 - sqrt((k - 1.0) / k) * a0;
        A[(int64_t)k * S + j] = ak;
        a0 = a1;
        a1 = ak;
    }
}

// C[i][j] = sum_k X(i, k) Y(k, j), complex result, 64 x 64 tile per CTA, 4 x 4 per thread, K step 16.
//   MODE 0: X = rho[i][k] complex (ld = ldx),  Y = A[k][j] real     -> M = rho A          (I = N, K = N, J = S)
//   MODE 1: X = A[k][i] real (transposed use), Y = M[k][j] complex  -> R = A^T (rho A)    (I = S, K = N, J = S)
template <int MODE>
__global__ void __launch_bounds__(256) fft_product_kernel(const double *__restrict__ X, int64_t ldx,
                                                          const double *__restrict__ Y, int64_t ldy,
                                                          double2 *__restrict__ C, int64_t ldc, int I, int K, int J)
{
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ double2 Xs[BK][BM + 1];
    __shared__ double2 Ys[BK][BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    double2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = make_double2(0.0, 0.0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        for (int e = tid; e < BK * BM; e += 256) {
            int kk, ii;
            if (MODE == 0) {  // rho rows are contiguous in k
                kk = e % BK;
                ii = e / BK;
            } else {          // A rows are contiguous in i
                ii = e % BM;
                kk = e / BM;
            }
            double2 v = make_double2(0.0, 0.0);
            if (i0 + ii < I && k0 + kk < K) {
                if (MODE == 0)
                    v = reinterpret_cast<const double2 *>(X)[(int64_t)(i0 + ii) * ldx + k0 + kk];
                else
                    v.x = X[(int64_t)(k0 + kk) * ldx + i0 + ii];
            }
            Xs[kk][ii] = v;
        }
        for (int e = tid; e < BK * BN; e += 256) {
            const int jj = e % BN, kk = e / BN;
            double2 v = make_double2(0.0, 0.0);
            if (j0 + jj < J && k0 + kk < K) {
                if (MODE == 0)
                    v.x = Y[(int64_t)(k0 + kk) * ldy + j0 + jj];
                else
                    v = reinterpret_cast<const double2 *>(Y)[(int64_t)(k0 + kk) * ldy + j0 + jj];
            }
            Ys[kk][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double2 xv[4], yv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xv[a] = Xs[kk][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) yv[b] = Ys[kk][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (MODE == 0) {  // complex x real
                        acc[a][b].x = fma(xv[a].x, yv[b].x, acc[a][b].x);
                        acc[a][b].y = fma(xv[a].y, yv[b].x, acc[a][b].y);
                    } else {          // real x complex
                        acc[a][b].x = fma(xv[a].x, yv[b].x, acc[a][b].x);
                        acc[a][b].y = fma(xv[a].x, yv[b].y, acc[a][b].y);
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i < I && j < J) C[(int64_t)i * ldc + j] = acc[a][b];
        }
}

// T[t] = (cos, sin)(2 pi t / n), t < n
__global__ void fft_twiddle_kernel(double2 *__restrict__ T, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double s, c;
    sincospi(2.0 * (double)t / (double)n, &s, &c);
    T[t] = make_double2(c, s);
}

// One CTA per x index i: the row's band R[i+d][i-d] goes to shared memory, every thread owns output columns c.
__global__ void __launch_bounds__(256) fft_band_dft_kernel(const double2 *__restrict__ R, const double2 *__restrict__ T,
                                                           double *__restrict__ W, int S, double scale)
{
    extern __shared__ double2 band[];
    const int i = blockIdx.x, n = 2 * S;
    const int m = min(i, S - 1 - i);
    for (int e = threadIdx.x; e <= 2 * m; e += blockDim.x) {
        const int d = e - m;
        band[e] = R[(int64_t)(i + d) * S + (i - d)];
    }
    __syncthreads();
    const int k_first = (3 * n) / 4, n_hi = n - k_first;  // columns [0, n_hi): k = 3n/4 + c;  then k = c - n_hi
    for (int c = threadIdx.x; c < S; c += blockDim.x) {
        const int k = c < n_hi ? k_first + c : c - n_hi;
        // exp(-i d theta_k), theta_k = 2 pi k / n, is a geometric sequence in d: taken exactly from the table every
        // kResync terms (residue of d k mod n) and advanced by one complex multiplication in between -- a table
        // gather per term made this kernel wait on the load/store unit (0.7 TFLOP/s at S = 2048)
        constexpr int kResync = 32;
        const double2 wk = __ldg(T + k);
        int t = (int)((((int64_t)(-m) * k) % n + n) % n);  // residue at d = -m
        const int t_step = (int)(((int64_t)kResync * k) % n);
        double acc = 0.0;
        const int len = 2 * m + 1;
        for (int e0 = 0; e0 < len; e0 += kResync) {
            double2 z = __ldg(T + t);
            const int e1 = min(e0 + kResync, len);
            for (int e = e0; e < e1; ++e) {
                const double2 w = band[e];
                acc = fma(w.x, z.x, fma(w.y, z.y, acc));  // Re(w exp(-i d theta))
                const double zx = fma(z.x, wk.x, -(z.y * wk.y));
                z.y = fma(z.x, wk.y, z.y * wk.x);
                z.x = zx;
            }
            t += t_step;
            if (t >= n) t -= n;
        }
        W[(int64_t)c * S + i] = acc * scale;
    }
}

cudaError_t launch_wigner_fft(const double2 *rho, int N, int64_t ld, const double *d_xvec, int S, double g, double scale,
                              double *A, double2 *M, double2 *R, double2 *T, double *W, cudaStream_t st, int *launches)
{
    osc_eigen_kernel<<<(S + 127) / 128, 128, 0, st>>>(A, N, S, d_xvec, g);
    fft_twiddle_kernel<<<(2 * S + 255) / 256, 256, 0, st>>>(T, 2 * S);
    dim3 g1((S + 63) / 64, (N + 63) / 64), g2((S + 63) / 64, (S + 63) / 64);
    fft_product_kernel<0><<<g1, 256, 0, st>>>(reinterpret_cast<const double *>(rho), ld, A, S, M, S, N, N, S);
    fft_product_kernel<1><<<g2, 256, 0, st>>>(A, S, reinterpret_cast<const double *>(M), S, R, S, S, N, S);
    const size_t smem = (size_t)S * sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(fft_band_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fft_band_dft_kernel<<<S, 256, smem, st>>>(R, T, W, S, scale);
    if (launches) *launches += 5;
    return cudaGetLastError();
}

}  // namespace bq
