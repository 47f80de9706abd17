/*
 * bq_b200.h -- C ABI of the B200-native statevector -> reduced rho -> Wigner grid path.
 *
 * The reference (C2QA/bosonic-qiskit) has no FFI for this path: the seam is two
 * Python call sites into third-party libraries,
 *     qutip.wigner(psi=Qobj(rho), xvec, yvec, g, method)      src/bosonic_qiskit/wigner.py:190
 *     qiskit.quantum_info.partial_trace(state, qargs)         src/bosonic_qiskit/util.py:580,596,619,693
 * plus the frame loops that call them once per frame
 *     src/bosonic_qiskit/wigner.py:92-96, src/bosonic_qiskit/animate.py:180-206,326-352,
 *     src/bosonic_qiskit/util.py:688-694.
 * Every entry point below names the seam it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - complex128 is passed as interleaved (re, im) doubles: `const double *` of 2*count.
 *   - All functions return 0 on success, a negative bq_status otherwise; the message
 *     for the calling thread's last failure is bq_last_error().
 *   - `_host` entry points take HOST pointers, run on the handle's device, and return
 *     when the outputs are complete in host memory.
 *   - `_dev` entry points take DEVICE pointers of the handle's device and enqueue on
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) without
 *     synchronising.
 *   - There is no CPU fallback: without a usable CUDA device bq_create fails.
 *   - A handle may be used from any host thread, one call at a time.
 */
#ifndef BQ_B200_H
#define BQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BQ_B200_VERSION 200 /* major*10000 + minor*100 + patch */

typedef enum bq_status {
    BQ_OK = 0,
    BQ_ERR_INVALID = -1,  /* bad argument (NULL, negative size, bad qubit index, non power of two ...) */
    BQ_ERR_CUDA = -2,     /* a CUDA runtime call failed; see bq_last_error() */
    BQ_ERR_NO_DEVICE = -3,/* no CUDA device / requested ordinal absent */
    BQ_ERR_UNSUPPORTED = -4 /* e.g. method="fft" handed to a grid entry point (it has its own: bq_wigner_fft_*) */
} bq_status;

/* qutip.wigner `method=` values (src/bosonic_qiskit/wigner.py:22 WignerMethods).  The three grid
 * methods are the same function evaluated by different recurrences; this library serves all three
 * names from the Clenshaw kernel.  "fft" evaluates W on a p axis it derives from the x grid and returns that
 * axis too, so it has its own entry points (bq_wigner_fft_host / _dev); the grid entry points answer
 * BQ_WIGNER_FFT with BQ_ERR_UNSUPPORTED. */
typedef enum bq_wigner_method {
    BQ_WIGNER_CLENSHAW = 0,
    BQ_WIGNER_ITERATIVE = 1,
    BQ_WIGNER_LAGUERRE = 2,
    BQ_WIGNER_FFT = 3
} bq_wigner_method;

typedef struct bq_handle_s *bq_handle_t;

/* ---- lifetime ------------------------------------------------------------------------------ */
int bq_version(void);
const char *bq_last_error(void);
int bq_device_count(int *count);
int bq_create(bq_handle_t *out, int device);
int bq_destroy(bq_handle_t h);
int bq_device(bq_handle_t h, int *device);
/* kernels this handle has launched since creation (bench.py's gpu_launches) */
int bq_launch_count(bq_handle_t h, uint64_t *count);
/* blocks until everything the handle enqueued on its own streams is done */
int bq_synchronize(bq_handle_t h);

/* ---- pinned host buffers (optional; outputs placed here are written by the GPU directly) ---- */
int bq_host_alloc(bq_handle_t h, size_t bytes, void **out);
int bq_host_free(bq_handle_t h, void *ptr);

/* ---- partial trace: replaces qiskit.quantum_info.partial_trace ------------------------------ */
/* Statevector branch (src/bosonic_qiskit/util.py:580,596,619,693).
 *   psi      batch statevectors, each 2^n_qubits complex128, `psi_stride` complex elements apart
 *   traced   n_traced distinct qubit positions (Qiskit little-endian bit index) to sum over
 *   rho_out  batch x D x D complex128 row-major, D = 2^(n_qubits-n_traced); kept qubits keep
 *            their significance order.  n_traced == n_qubits gives the 1x1 norm <psi|psi>.        */
int bq_partial_trace_sv_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                             int batch, int64_t psi_stride, double *rho_out);
int bq_partial_trace_sv_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                            int batch, int64_t psi_stride, double *d_rho_out, void *stream);

/* Several traces of the SAME statevector in one call: the per-qumode loop of
 * util.avg_photon_num (src/bosonic_qiskit/util.py:688-694).  `traced` holds n_sets lists of
 * n_traced positions each (all sets trace the same number of qubits); rho_out is n_sets x D x D. */
int bq_partial_trace_sv_multi_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced,
                                   int n_traced, int n_sets, double *rho_out);
int bq_partial_trace_sv_multi_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced,
                                  int n_traced, int n_sets, double *d_rho_out, void *stream);

/* Shot-averaged reduced density matrix: rho = sum_k w_k Tr_traced |psi_k><psi_k| over `shots` statevectors of ONE
 * circuit, in one batched launch of the tensor-core contraction with the accumulation over shots fused into its
 * split reduction.  Replaces the Python loop that calls trace_out_qubits per shot and averages the results over the
 * per-shot statevectors simulate() collects (src/bosonic_qiskit/util.py:523-532;
 * tutorials/bosonic-kitten-code/bosonic-kitten-code.ipynb cell 14) -- the mixed-state input of the Wigner routine
 * under noise.  weights: `shots` doubles, or NULL for the plain mean (w_k = 1 / shots).  rho_out: D x D.          */
int bq_partial_trace_sv_accumulate_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                        int shots, int64_t psi_stride, const double *weights, double *rho_out);
int bq_partial_trace_sv_accumulate_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                                       int shots, int64_t psi_stride, const double *d_weights, double *d_rho_out,
                                       void *stream);

/* DensityMatrix branch: rho_keep[i,j] = sum_t rho[(i,t),(j,t)]  (reached through
 * cv_partial_trace / avg_photon_num with a DensityMatrix, src/bosonic_qiskit/util.py:619,693).  */
int bq_partial_trace_dm_host(bq_handle_t h, const double *rho, int n_qubits, const int *traced, int n_traced,
                             int batch, double *rho_out);
int bq_partial_trace_dm_dev(bq_handle_t h, const double *d_rho, int n_qubits, const int *traced, int n_traced,
                            int batch, double *d_rho_out, void *stream);

/* <n> = sum_n n rho_nn / sum_n rho_nn per batch element, unrounded; out[2*b] real, out[2*b+1] imag
 * (util.qumode_avg_photon_num, src/bosonic_qiskit/util.py:699-736, checks the imaginary part).  */
int bq_photon_number_host(bq_handle_t h, const double *rho, int D, int batch, double *out);
int bq_photon_number_dev(bq_handle_t h, const double *d_rho, int D, int batch, double *d_out, void *stream);

/* ---- Wigner grid: replaces qutip.wigner(Qobj(rho), xvec, yvec, g, method) -------------------- */
/*   rho   batch x N x N complex128, row-major with leading dimension `ld` (complex elements) and
 *         `rho_stride` complex elements between frames; only the upper triangle is read
 *   W_out batch x ny x nx float64, W[b][iy][ix]  (src/bosonic_qiskit/wigner.py:178-190)          */
int bq_wigner_host(bq_handle_t h, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                   const double *xvec, int nx, const double *yvec, int ny, double g, int method,
                   double *W_out);
int bq_wigner_dev(bq_handle_t h, const double *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                  const double *d_xvec, int nx, const double *d_yvec, int ny, double g, int method,
                  double *d_W_out, void *stream);

/* method="fft" of qutip.wigner (qutip/wigner.py::_wigner_fourier, reached through
 * src/bosonic_qiskit/wigner.py:190 with method="fft"): the Fourier transform of <x - s|rho|x + s> along s on the
 * x grid `xvec` (uniform spacing, nx >= 2), evaluated for all eigenvectors of rho at once (no eigensolver, see
 * csrc/wigner_fft_kernels.cu).  The whole of rho is read (it need not be Hermitian).
 *   xvec      HOST pointer in both variants (the p axis is computed from it on the host)
 *   W_out     batch x nx x nx float64, W[b][ip][ix]
 *   yvec_out  HOST, nx float64: the p axis of W_out (what qutip returns as the second tuple element)          */
int bq_wigner_fft_host(bq_handle_t h, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                       const double *xvec, int nx, double g, double *W_out, double *yvec_out);
int bq_wigner_fft_dev(bq_handle_t h, const double *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                      const double *xvec, int nx, double g, double *d_W_out, double *yvec_out, void *stream);

/* ---- fused frame loop: psi_k -> rho_k -> W_k for a batch of frames --------------------------- */
/* Replaces the loop bodies at src/bosonic_qiskit/wigner.py:92-96 and animate.py:326-352
 * (trace_out_qubits + _wigner per frame).  n_traced == 0 means rho_k = |psi_k><psi_k|
 * (DensityMatrix(state), wigner.py:69,95).  rho_out / d_rho_out may be NULL.                      */
int bq_state_to_wigner_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                            int batch, int64_t psi_stride, const double *xvec, int nx, const double *yvec,
                            int ny, double g, int method, double *W_out, double *rho_out);
int bq_state_to_wigner_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                           int batch, int64_t psi_stride, const double *d_xvec, int nx, const double *d_yvec,
                           int ny, double g, int method, double *d_W_out, double *d_rho_out, void *stream);

/* Several traces of ONE statevector, each followed by its Wigner grid, in one call: per-site reduced density
 * matrices + per-site Wigner functions (the per-qumode loop of src/bosonic_qiskit/util.py:688-694 followed by
 * wigner.py:190 per site; BASELINE config 4).  `traced` holds n_sets lists of n_traced positions; W_out is
 * n_sets x ny x nx, rho_out (optional) n_sets x D x D.  One kernel launch when the traces are small.             */
int bq_state_to_wigner_multi_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                  int n_sets, const double *xvec, int nx, const double *yvec, int ny, double g,
                                  int method, double *W_out, double *rho_out);
int bq_state_to_wigner_multi_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                                 int n_sets, const double *d_xvec, int nx, const double *d_yvec, int ny, double g,
                                 int method, double *d_W_out, double *d_rho_out, void *stream);

/* per-frame (min, max) of a finished batch of grids: the colour-scale reduction of
 * animate._animate (src/bosonic_qiskit/animate.py:265-311).  out[2*b] = min, out[2*b+1] = max.    */
int bq_frame_minmax_dev(bq_handle_t h, const double *d_W, int64_t points_per_frame, int batch, double *d_out,
                        void *stream);
/* The colour matplotlib gives every grid node of a frame in animate._animate / wigner.plot
 * (src/bosonic_qiskit/animate.py:278-291, wigner.py:260-269): colour levels linspace(-A, A, levels) with
 * A = max(amax, |amin|) of the frame (zero_abs_max if that is 0: 5 in _animate, 1 in plot) and
 * contourf(..., levels, cmap="RdBu").  d_rgba_out: batch x points x 4 bytes (R, G, B, A = 255), same point order as W. */
int bq_frame_rgba_dev(bq_handle_t h, const double *d_W, int64_t points_per_frame, int batch, int levels,
                      double zero_abs_max, unsigned char *d_rgba_out, void *stream);
/* the 256 x 4 byte lookup table it uses (matplotlib's "RdBu", restated); host only, needs no device */
int bq_rdbu_lut(unsigned char *lut_out);

/* ---- measurement helpers (bench.py only) ----------------------------------------------------- */
/* Runs a register-resident DFMA loop on every SM and reports the sustained FP64 rate: the
 * roofline denominator for the Wigner kernel (MEASURED_PEAKS.json has no fp64 entry).           */
int bq_fp64_peak(bq_handle_t h, double seconds, double *tflops);
/* device time (ms) the most recent `_host` call spent between its first H2D and last D2H */
int bq_last_device_ms(bq_handle_t h, double *ms);
/* When enabled, every launch of the Wigner grid kernel (K3) is bracketed by CUDA events on the
 * stream it is launched on; bq_profile_read synchronises those streams, returns the summed device
 * time and the number of launches since the last read, and clears the record.                    */
int bq_profile_enable(bq_handle_t h, int on);
/* Which machine code of the Wigner kernels is launched: mode 1 = the build's re-scheduled SASS (same
 * instructions as nvcc's output, DFMAs re-ordered into operand-reuse chains; the default when the build
 * carries it, overridable with the environment variable BQ_B200_SASS=plain or =0), mode 0 = exactly what nvcc
 * scheduled.  Results are bit-identical.  *active (may be NULL) receives the mode in effect.            */
int bq_set_sass_mode(bq_handle_t h, int mode, int *active);
/* Result of the load-time self-check of the re-scheduled SASS: when the first handle on a device is created, every
 * patched kernel variant and its nvcc-scheduled twin evaluate the same probe grids and are compared bit for bit
 * (environment BQ_B200_SASS_VERIFY=0 skips it).  *bad_mask: bit v set = variant v differed (it runs nvcc's schedule
 * from then on); *checked: variants compared.  Either pointer may be NULL.                                        */
int bq_sass_self_check(bq_handle_t h, uint64_t *bad_mask, int *checked);
int bq_profile_read(bq_handle_t h, double *total_ms, int *launches);

/* ---- mirror-symmetric grids ------------------------------------------------------------------
 * The reference only ever evaluates W on `xvec = linspace(-a, a, S)`, `yvec = xvec`
 * (src/bosonic_qiskit/wigner.py:135,187-188).  On such a grid the eight points (+-x_a, +-x_b),
 * (+-x_b, +-x_a) share |A2|^2, hence the whole inner Clenshaw recurrence; the library then runs one
 * recurrence per pair of grid lines and eight Horner accumulators (same results to ~1e-15, 6-8x fewer
 * fp64 instructions).  mode 0: never (always the point-by-point kernels); mode 1 (default): the `_host`
 * entry points test the grid they are given (bq_grid_is_mirror), the `_dev` entry points -- which cannot
 * read a device grid without a synchronisation -- use the point-by-point kernels; mode 2: as 1, and the
 * caller vouches that the grids passed to `_dev` entry points are mirror-symmetric with yvec == xvec
 * until the mode is changed (the Python binding tests device grids once and caches the verdict).      */
int bq_set_grid_symmetry(bq_handle_t h, int mode);
/* One large mirror-symmetric grid shared by several GPUs (BASELINE config 5): the grid's bq_grid_units(nx)
 * units (pairs of grid lines, 8 points each) all cost the same, so each rank takes a contiguous range of
 * them: the next Wigner calls on this handle evaluate units [first, first + count) only and leave every other
 * point of W untouched (zero the buffer first; the ranks' results then add up exactly, one reduce replaces
 * the gather).  count < 0 restores the whole grid.  A call that cannot use the symmetric-grid kernels while a
 * range is set fails with BQ_ERR_UNSUPPORTED (shard by row slab instead).                               */
int bq_set_unit_range(bq_handle_t h, int64_t first, int64_t count);
int64_t bq_grid_units(int nx);
/* Tuning / test aid: restrict the symmetric-grid kernels to ONE variant of csrc/wigner_kernels.cu's variant table
 * (its id; -1 = the library's own choice by predicted duration).  A variant that cannot run the call (cutoff class,
 * shared memory) leaves the point-by-point kernels to do it.  Same as the environment variable BQ_B200_SYM_VARIANT. */
int bq_set_sym_variant(bq_handle_t h, int variant);
/* id of the kernel variant the last Wigner call launched (bench.py names the kernel it reports a roofline for) */
int bq_last_wigner_variant(bq_handle_t h, int *variant);
/* 1 if nx == ny >= 2, yvec[i] == xvec[i] and |xvec[i] + xvec[nx-1-i]| <= 4 ulp of max|xvec|; else 0 */
int bq_grid_is_mirror(const double *xvec, int nx, const double *yvec, int ny);

/* ---- several GPUs from ONE process ------------------------------------------------------------
 * The reference fans its frame loop out from inside a plain function call (multiprocessing.Pool.starmap over
 * frames, src/bosonic_qiskit/animate.py:166-169,191-206; the sequential loops at wigner.py:92-96 and
 * animate.py:326-352).  A bq_multi_t owns one handle + one worker thread per GPU of the box; the sharded entry
 * points take HOST pointers like the `_host` entry points and return when W_out is complete.  No collective on
 * the data path:
 *   BQ_SHARD_FRAMES  contiguous frame blocks per GPU, every GPU copies its finished grids straight into its slice
 *                    of W_out over its own PCIe link (page-locked W_out, e.g. from bq_host_alloc, makes those
 *                    copies concurrent);
 *   BQ_SHARD_GRID    ONE grid (batch == 1): a mirror-symmetric grid is shared out by equal unit ranges, every GPU
 *                    stores its points into one frame buffer on the first GPU through NVLink peer memory, one copy
 *                    to the host; any other grid by row slabs.
 * bq_multi_init: devs == NULL / n_dev <= 0 takes every visible GPU (n_dev > 0 with devs == NULL: the first n_dev);
 * a device listed more than once takes that many shares.                                                          */
typedef struct bq_multi_s *bq_multi_t;
enum { BQ_SHARD_FRAMES = 0, BQ_SHARD_GRID = 1 };
int bq_multi_init(bq_multi_t *out, int n_dev, const int *devs);
int bq_multi_finalize(bq_multi_t m);
int bq_multi_device_count(bq_multi_t m, int *count);
/* the per-device handle (index < device count), e.g. for bq_set_sass_mode / bq_profile_enable; owned by `m` */
int bq_multi_handle(bq_multi_t m, int index, bq_handle_t *out);
/* kernels launched by all of m's handles since creation */
int bq_multi_launch_count(bq_multi_t m, uint64_t *count);
/* sharded bq_state_to_wigner_host (frame loops of wigner.py:92-96, animate.py:191-206,326-352) */
int bq_state_to_wigner_sharded_host(bq_multi_t m, const double *psi, int n_qubits, const int *traced, int n_traced,
                                    int batch, int64_t psi_stride, const double *xvec, int nx, const double *yvec,
                                    int ny, double g, int method, int shard_mode, double *W_out, double *rho_out);
/* sharded bq_wigner_host (qutip.wigner per frame, wigner.py:190) */
int bq_wigner_sharded_host(bq_multi_t m, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                           const double *xvec, int nx, const double *yvec, int ny, double g, int method, int shard_mode,
                           double *W_out);

/* ---- CUDA IPC plumbing for the multi-GPU tile gather ---------------------------------------- */
/* One process per GPU: the root exports its gather buffer, peers open it and let the Wigner
 * kernel store finished tiles straight into it over NVLink (replaces the pickled return of
 * multiprocessing.Pool.starmap, src/bosonic_qiskit/animate.py:191-206).  handle64 = 64 bytes.   */
/* plain cudaMalloc / cudaFree on the handle's device: IPC handles name whole allocations, so the
 * gather buffer must not be a sub-range of a caching allocator's block */
int bq_device_alloc(bq_handle_t h, size_t bytes, void **d_ptr);
int bq_device_free(bq_handle_t h, void *d_ptr);
int bq_ipc_export(bq_handle_t h, void *d_ptr, unsigned char *handle64);
int bq_ipc_open(bq_handle_t h, const unsigned char *handle64, void **d_ptr);
int bq_ipc_close(bq_handle_t h, void *d_ptr);
/* asynchronous device-to-device copy on `stream` (unified addressing: either side may be a peer GPU's buffer
 * opened with bq_ipc_open).  Used when a finished block of grids was computed locally by the symmetric-grid
 * kernels (scattered stores: kept inside the GPU) and then goes to the gather buffer as one NVLink copy.  */
int bq_memcpy_dev(bq_handle_t h, void *d_dst, const void *d_src, size_t bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BQ_B200_H */
