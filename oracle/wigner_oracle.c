/*
 * Scalar C restatement of the Wigner/partial-trace path -- ORACLE, test
 * infrastructure only (see oracle/__init__.py).  Nothing in the product
 * (bosonic_qiskit_b200/) links or loads this file.
 *
 * Follows, point by point instead of grid-at-a-time, the algorithm behind
 *   src/bosonic_qiskit/wigner.py:190   qutip.wigner(..., method="clenshaw")
 *     (qutip 5.2.2 qutip/wigner.py::_wigner_clenshaw / _wig_laguerre_val)
 *   src/bosonic_qiskit/util.py:580,596,619,693  qiskit partial_trace
 *     (qiskit 2.1.2 quantum_info/states/utils.py, Statevector branch)
 *
 * REAL selects the working precision: double (default) or long double
 * (-DORACLE_LONG_DOUBLE) -- the x87 80-bit build is the "more digits than the
 * reference" check used by tests/ for large cutoffs.  Operation order inside a
 * point is the reference's (no FMA contraction is requested; build with
 * -ffp-contract=off).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef ORACLE_LONG_DOUBLE
typedef long double REAL;
#define SQRT sqrtl
#define EXP expl
#define HYPOT hypotl
#define SUFFIX(n) n##_ld
#else
typedef double REAL;
#define SQRT sqrt
#define EXP exp
#define HYPOT hypot
#define SUFFIX(n) n##_f64
#endif

/* rho: N x N complex128 row-major, interleaved (re, im), leading dimension ld (in complex elements).
 * W:   ny x nx doubles, W[iy*nx + ix]. Rows [row0, row1) only are written. */
int SUFFIX(oracle_wigner_clenshaw)(const double *rho, int N, int64_t ld, const double *xvec, int nx,
                                   const double *yvec, int ny, double g_, int row0, int row1, double *W)
{
    if (N < 1 || nx < 0 || ny < 0 || row0 < 0 || row1 > ny) return -1;
    const REAL g = (REAL)g_;
    const REAL pi = (REAL)3.14159265358979323846264338327950288L;
    for (int iy = row0; iy < row1; ++iy) {
        for (int ix = 0; ix < nx; ++ix) {
            const REAL ar = g * (REAL)xvec[ix], ai = g * (REAL)yvec[iy];
            REAL B = HYPOT(ar, ai);
            B *= B;
            /* w0 = 2 rho[0, N-1] */
            REAL wr = 2 * (REAL)rho[2 * (size_t)(N - 1)], wi = 2 * (REAL)rho[2 * (size_t)(N - 1) + 1];
            for (int L = N - 2; L >= 0; --L) {
                const int K = N - L;
                /* c[n] = f * rho[n, n+L], f = 1 on the main diagonal, 2 off it */
                const REAL f = (L == 0) ? 1 : 2;
#define CRE(n) (f * (REAL)rho[2 * ((size_t)(n) * ld + (n) + L)])
#define CIM(n) (f * (REAL)rho[2 * ((size_t)(n) * ld + (n) + L) + 1])
                REAL y0r, y0i, y1r, y1i;
                if (K == 1) {
                    y0r = CRE(0); y0i = CIM(0); y1r = 0; y1i = 0;
                } else if (K == 2) {
                    y0r = CRE(0); y0i = CIM(0); y1r = CRE(1); y1i = CIM(1);
                } else {
                    int k = K;
                    y0r = CRE(K - 2); y0i = CIM(K - 2);
                    y1r = CRE(K - 1); y1i = CIM(K - 1);
                    for (int i = 3; i <= K; ++i) {
                        --k;
                        const REAL a = SQRT((REAL)((double)(k - 1) * (double)(L + k - 1)) /
                                            (REAL)((double)(L + k) * (double)k));
                        const REAL t = ((REAL)(L + 2 * k - 1) - B) / SQRT((REAL)((double)(L + k) * (double)k));
                        const REAL n0r = CRE(K - i) - y1r * a, n0i = CIM(K - i) - y1i * a;
                        const REAL n1r = y0r - y1r * t, n1i = y0i - y1i * t;
                        y0r = n0r; y0i = n0i; y1r = n1r; y1i = n1i;
                    }
                }
                const REAL s1 = 1 / SQRT((REAL)(L + 1));
                const REAL t1 = ((REAL)(L + 1) - B) * s1;
                const REAL Sr = y0r - y1r * t1, Si = y0i - y1i * t1;
                /* w0 = S + w0 * A2 * (L+1)^-1/2 */
                const REAL pr = (wr * ar - wi * ai) * s1, pim = (wr * ai + wi * ar) * s1;
                wr = Sr + pr;
                wi = Si + pim;
#undef CRE
#undef CIM
            }
            W[(size_t)iy * nx + ix] = (double)(wr * EXP(-B * (REAL)0.5) * (g * g * (REAL)0.5 / pi));
        }
    }
    return 0;
}

#ifndef ORACLE_LONG_DOUBLE
/* rho_keep = Psi Psi^dagger, traced = bit positions summed over (qiskit little-endian).
 * psi: 2^n complex128 interleaved; rho_out: D x D complex128 row-major interleaved, D = 2^(n-n_traced). */
static uint64_t deposit(uint64_t v, const int *pos, int npos)
{
    uint64_t r = 0;
    for (int b = 0; b < npos; ++b) r |= ((v >> b) & 1ull) << pos[b];
    return r;
}

int oracle_partial_trace_sv(const double *psi, int n_qubits, const int *traced, int n_traced, double *rho_out)
{
    if (n_qubits < 0 || n_qubits > 40 || n_traced < 0 || n_traced > n_qubits) return -1;
    int is_traced[64] = {0}, keep[64], tr[64], nk = 0, nt = 0;
    for (int i = 0; i < n_traced; ++i) {
        if (traced[i] < 0 || traced[i] >= n_qubits || is_traced[traced[i]]) return -2;
        is_traced[traced[i]] = 1;
    }
    for (int q = 0; q < n_qubits; ++q) {
        if (is_traced[q]) tr[nt++] = q; else keep[nk++] = q;
    }
    const uint64_t D = 1ull << nk, T = 1ull << nt;
    for (int64_t i = 0; i < (int64_t)D; ++i) {
        const uint64_t bi = deposit((uint64_t)i, keep, nk);
        for (uint64_t j = 0; j < D; ++j) {
            const uint64_t bj = deposit(j, keep, nk);
            double sr = 0, si = 0;
            for (uint64_t t = 0; t < T; ++t) {
                const uint64_t o = deposit(t, tr, nt);
                const double xr = psi[2 * (bi | o)], xi = psi[2 * (bi | o) + 1];
                const double yr = psi[2 * (bj | o)], yi = psi[2 * (bj | o) + 1];
                sr += xr * yr + xi * yi;  /* x * conj(y) */
                si += xi * yr - xr * yi;
            }
            rho_out[2 * ((size_t)i * D + j)] = sr;
            rho_out[2 * ((size_t)i * D + j) + 1] = si;
        }
    }
    return 0;
}
#endif
