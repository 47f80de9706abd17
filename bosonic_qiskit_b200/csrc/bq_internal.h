// Internal host-side declarations shared by the kernel translation units and the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bq {

constexpr int kResidentMaxN = 64;  // stream fits twice per SM in shared memory up to this cutoff
constexpr int kMaxQubits = 40;
constexpr int kKernelSlots = 64;

// records in the coefficient stream of cutoff N, and the first record of the diagonal with K = N-L
// coefficients (K >= 2); layout described in bq_common.cuh
__host__ __device__ inline int64_t stream_records(int N) { return (int64_t)N * (N + 1) / 2 + N - 1; }
__host__ __device__ inline int64_t diag_pos(int K) { return ((int64_t)K * K + K - 4) / 2; }

// Whether a mirror-symmetric grid of `units` units (all frames of the launch) runs on the recurrence-sharing
// kernels: always for the resident cutoffs, for the ring kernels only when there are enough units to fill the GPU
// (below that the point-by-point kernels finish first).  One rule for the single-GPU launcher and for the
// multi-GPU unit-range sharding, so that both evaluate a given grid with the same kernels (bit-identical results).
inline bool sym_kernels_cover(int N, int64_t units, int sm_count)
{
    return N <= kResidentMaxN || units >= 2LL * 256 * (sm_count / 2);
}

struct DeviceInfo {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    // per-kernel launch cache (attribute set once, occupancy looked up once per shared-memory size)
    int pts8 = 0;       // use the 8-points-per-thread resident kernel for large grids
    int sym_force = -1; // >= 0: tuning aid (env BQ_B200_SYM_VARIANT), only this symmetric-kernel variant is eligible
    int sym_split_main = -1, sym_split_tail = -1;  // tuning aid (env BQ_B200_SYM_SPLIT=main:tail): whole rounds / remainder
    int sym_phase = 1;  // symmetric kernels: stagger the warps of a scheduler (env BQ_B200_SYM_PHASE=0 turns it off)
    int last_variant = -1;  // variant id (kKernelNames index) of the last Wigner kernel launched
    int sass_mode = 1;  // 1: launch the re-scheduled (patched) SASS of the Wigner kernels when the build has it
    uint64_t patched_bad = 0;  // bit v: the patched image of variant v failed the load-time self-check -> nvcc's schedule
    bool configured[kKernelSlots] = {};
    int occ[kKernelSlots] = {};
    size_t occ_smem[kKernelSlots] = {};
};

struct TraceMap {
    // bit positions (ascending significance) of kept and traced qubits in the statevector index
    int keep[kMaxQubits];
    int trace[kMaxQubits];
    int n_keep, n_trace;
};

struct WigParams {
    const double2 *cs;   // [batch][R_tot] rho-dependent stream
    const double4 *tab;  // [R_tot] recurrence constants
    const double *xvec, *yvec;
    double *W;  // [batch][ny][nx]
    int64_t w_stride, cs_stride;
    int N, R_tot, nx, ny, batch;
    int chunks_per_row, tiles_per_frame, total_tiles, grab, vec_ok;
    double g, scale;
    unsigned int *ctr;  // [0] next tile, [1] CTAs that left, [2] CTAs that finished the fused prologue, [3] ... its first pass
    // Fused prologue (one launch per step instead of trace + split reduce + pack + Wigner): before any tile is
    // started the CTAs of the launch build the coefficient stream `cs` themselves -- from rho (pro_mode 1, what
    // pack_rho_kernel does) or straight from the statevectors (pro_mode 2: the partial trace rho = Psi Psi^dagger of
    // every frame, optionally also written out as pro_rho_out) -- and meet at a grid-wide barrier (all CTAs of these
    // persistent launches are co-resident).  pro_mode 0: `cs` was filled by an earlier launch.
    int pro_mode;
    double2 *cs_w;                    // writable alias of cs
    const double2 *pro_rho;           // mode 1: [batch][N][ld]
    int64_t pro_ld, pro_rho_stride;
    const double2 *pro_psi;           // mode 2: pro_psi_batch statevectors, pro_psi_stride apart; frame f = set * pro_psi_batch + b
    int64_t pro_psi_stride;           //         is the trace of statevector b with map `set` (layout of launch_trace_sv)
    const TraceMap *pro_maps;
    int pro_n_keep, pro_n_trace, pro_psi_batch;
    double2 *pro_rho_out;             // mode 2, optional: [batch][N][N]
    // mode 4 (cutoff <= 32 kept, 512 .. 8192 traced indices; resident kernels only): a CTA per (frame, slice of the traced
    // range), the slice of Psi staged in shared memory behind the kernel's own
    int pro_slices;                   //         slices of the traced range (a power of two; slice length 32 .. 256)
    double2 *pro_partial;             //         scratch [batch][N (N + 1) / 2][pro_slices]
    int pro_smem_bytes, pro_smem_off; //         bytes of shared memory the prologue needs (host) and where they start (launcher)
    // mirror-symmetric grids (wigner_sym_* kernels): U = ceil(S/2) mirror classes, U(U+1)/2 units per frame
    int sym, sym_U, sym_phase;
    int64_t sym_units;  // units this launch evaluates per frame ...
    int64_t sym_first;  // ... starting at this unit (a rank's share of one large grid; 0 = the whole grid)
    int64_t sym_want_first, sym_want_count;  // caller's unit range (count < 0: all)
};

cudaError_t launch_build_tab(double4 *tab, int N, cudaStream_t st);
cudaError_t launch_pack_rho(const double2 *rho, int64_t ld, int64_t rho_stride, int N, int batch, double2 *cs,
                            int64_t cs_stride, cudaStream_t st);
cudaError_t launch_wigner(WigParams p, DeviceInfo &dev, cudaStream_t st, int *launches);
bool patched_sass_available();
// Load-time self-check of the re-scheduled SASS (once per device and process): every patched kernel variant and its
// nvcc-scheduled twin evaluate the same probe grid; a variant whose results differ in any bit is reported in the
// returned mask and runs nvcc's schedule from then on.  *checked receives the number of variants compared.
cudaError_t verify_patched_kernels(DeviceInfo &dev, uint64_t *bad_mask, int *checked);
// method="fft": scratch A [N][S] doubles, M [N][S], R [S][S], T [2S] complex; W is [S][S] (indexed [ip][ix])
cudaError_t launch_wigner_fft(const double2 *rho, int N, int64_t ld, const double *d_xvec, int S, double g, double scale,
                              double *A, double2 *M, double2 *R, double2 *T, double *W, cudaStream_t st, int *launches);

// K1: rho[set][b] = Psi Psi^dagger.  `maps` is a device array of n_sets TraceMaps (same n_keep / n_trace);
// rho_out is [n_sets][batch][D][D]; `partial` is scratch of at least trace_sv_scratch_bytes().
size_t trace_sv_scratch_bytes(int n_keep, int n_trace, int batch, int n_sets);
cudaError_t launch_trace_sv(const double2 *psi, int64_t psi_stride, int batch, const TraceMap *d_maps, int n_sets,
                            int n_keep, int n_trace, double2 *rho_out, double2 *partial, const DeviceInfo &dev,
                            cudaStream_t st, int *launches);
// shot average: rho_out[set] = sum_b w_b rho[set][b] (d_weights NULL: 1 / batch); scratch as for launch_trace_sv
cudaError_t launch_trace_sv_accumulate(const double2 *psi, int64_t psi_stride, int batch, const TraceMap *d_maps, int n_sets,
                                       int n_keep, int n_trace, const double *d_weights, double2 *rho_out, double2 *partial,
                                       const DeviceInfo &dev, cudaStream_t st, int *launches);
// K2: density-matrix branch
cudaError_t launch_trace_dm(const double2 *rho, int n_qubits, int batch, const TraceMap *d_map, int n_keep,
                            int n_trace, double2 *rho_out, cudaStream_t st, int *launches);
// K5: <n> per batch element, out[b] = (re, im)
cudaError_t launch_photon_number(const double2 *rho, int D, int batch, double2 *out, cudaStream_t st, int *launches);
// per-frame (min, max)
cudaError_t launch_frame_minmax(const double *W, int64_t points, int batch, double2 *out, cudaStream_t st,
                                int *launches);
// contourf-style RdBu colouring of frames at the grid nodes (levels symmetric colour levels); minmax from launch_frame_minmax
cudaError_t upload_rdbu_lut(const unsigned char *lut);
cudaError_t launch_frame_rgba(const double *W, int64_t points, int batch, const double2 *minmax, int levels,
                              double zero_abs_max, unsigned char *rgba, cudaStream_t st, int *launches);
// FP64 DFMA peak probe: runs `iters` rounds of 8 independent chains x 16 FMAs per thread
cudaError_t launch_fp64_probe(double *sink, int iters, int grid, cudaStream_t st);

}  // namespace bq
