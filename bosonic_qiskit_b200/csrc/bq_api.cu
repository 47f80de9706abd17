// C ABI of the library (include/bq_b200.h): argument checking, per-device workspaces, host <-> device
// staging for the `_host` entry points, and the orchestration psi -> rho -> coefficient stream -> W.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "../../include/bq_b200.h"
#include "bq_internal.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return BQ_ERR_CUDA;
}
#define CU(call)                                       \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

struct NvtxRange {  // one range per public call around its enqueues (K1 / K3 / K4 launches and copies)
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Buffer {  // grow-only device scratch
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (ptr) {
            cudaError_t e = cudaFree(ptr);  // implicit device sync: nothing in flight still reads it
            ptr = nullptr;
            bytes = 0;
            if (e != cudaSuccess) return e;
        }
        const size_t cap = std::max(need, (size_t)4096);
        cudaError_t e = cudaMalloc(&ptr, cap);
        if (e == cudaSuccess) bytes = cap;
        return e;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

constexpr int kCtrSlots = 64;

}  // namespace

struct bq_handle_s {
    bq::DeviceInfo dev;
    std::mutex mu;
    cudaStream_t stream = nullptr;  // used by the _host entry points
    cudaEvent_t ev_order = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaStream_t copy_stream = nullptr;  // device -> host copies of finished frame blocks, overlapped with the next block
    cudaEvent_t ev_block[4] = {}, ev_copied = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;

    std::map<int, double4 *> tabs;                          // cutoff -> constant table
    std::map<std::vector<int>, bq::TraceMap *> maps;        // (n_qubits, n_traced, sets...) -> device maps
    Buffer cs, partial, rho_tmp;                            // kernel scratch
    Buffer in_dev, out_dev, rho_dev, grid_dev, small_dev;   // staging for _host calls
    Buffer fft_a, fft_m, fft_r, fft_t, fft_x;               // method="fft" scratch
    Buffer mm_dev;                                          // per-frame (min, max) for bq_frame_rgba_dev
    bool lut_ready = false;
    std::vector<double> grid_host;                          // what grid_dev currently holds
    bool grid_host_mirror = false;                          // ... and whether it is a mirror-symmetric square grid
    int64_t unit_first = 0, unit_count = -1;                // bq_set_unit_range (count < 0: whole grid)
    int grid_mode = 1;  // 0: never share recurrences, 1: _host calls test their grid, 2: + caller vouches for _dev grids
    unsigned int *ctr = nullptr;
    int ctr_next = 0;
    uint64_t launches = 0;
    double last_ms = 0.0;
    bool profile = false;
    int sass_checked = 0;  // kernel variants the load-time SASS self-check compared on this device
    bool fused = true;  // the Wigner launch builds its coefficient stream in its own prologue (BQ_B200_FUSED=0: separate launches)
    bool local_prologue = true;  // tiny traces: per CTA, into shared memory (BQ_B200_LOCAL_PROLOGUE=0: the grid-wide prologue)
    bool staged_trace = true;    // few kept / many traced indices: slices of psi through shared memory (BQ_B200_STAGED_TRACE=0: a warp per element)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Scratch is shared by consecutive calls; if the caller hops streams, chain them.
int order_stream(bq_handle_t h, cudaStream_t st)
{
    if (h->have_last && h->last_stream != st) {
        CU(cudaEventRecord(h->ev_order, h->last_stream));
        CU(cudaStreamWaitEvent(st, h->ev_order, 0));
    }
    h->last_stream = st;
    h->have_last = true;
    return BQ_OK;
}

int get_tab(bq_handle_t h, int N, cudaStream_t st, const double4 **out)
{
    auto it = h->tabs.find(N);
    if (it == h->tabs.end()) {
        double4 *tab = nullptr;
        CU(cudaMalloc(&tab, (size_t)bq::stream_records(N) * sizeof(double4)));
        cudaError_t e = bq::launch_build_tab(tab, N, st);
        if (e != cudaSuccess) {
            cudaFree(tab);
            return fail_cuda(e, "build_tab");
        }
        ++h->launches;
        // later calls may come on another stream: make the table globally visible now
        CU(cudaStreamSynchronize(st));
        it = h->tabs.emplace(N, tab).first;
    }
    *out = it->second;
    return BQ_OK;
}

int build_map(int n_qubits, const int *traced, int n_traced, bq::TraceMap *m)
{
    if (n_qubits < 0 || n_qubits > bq::kMaxQubits) return fail(BQ_ERR_INVALID, "n_qubits out of range [0, 40]");
    if (n_traced < 0 || n_traced > n_qubits) return fail(BQ_ERR_INVALID, "n_traced out of range");
    if (n_traced > 0 && !traced) return fail(BQ_ERR_INVALID, "traced is NULL");
    bool is_tr[bq::kMaxQubits] = {};
    for (int i = 0; i < n_traced; ++i) {
        const int q = traced[i];
        if (q < 0 || q >= n_qubits) return fail(BQ_ERR_INVALID, "traced qubit index out of range");
        if (is_tr[q]) return fail(BQ_ERR_INVALID, "duplicate traced qubit index");
        is_tr[q] = true;
    }
    std::memset(m, 0, sizeof(*m));
    for (int q = 0; q < n_qubits; ++q) {
        if (is_tr[q])
            m->trace[m->n_trace++] = q;
        else
            m->keep[m->n_keep++] = q;
    }
    return BQ_OK;
}

int get_maps(bq_handle_t h, int n_qubits, const int *traced, int n_traced, int n_sets, const bq::TraceMap **out)
{
    std::vector<int> key;
    key.reserve(3 + (size_t)n_traced * n_sets);
    key.push_back(n_qubits);
    key.push_back(n_traced);
    key.push_back(n_sets);
    for (int i = 0; i < n_traced * n_sets; ++i) key.push_back(traced[i]);
    auto it = h->maps.find(key);
    if (it == h->maps.end()) {
        std::vector<bq::TraceMap> host((size_t)n_sets);
        for (int s = 0; s < n_sets; ++s) {
            int rc = build_map(n_qubits, traced ? traced + (size_t)s * n_traced : nullptr, n_traced, &host[s]);
            if (rc) return rc;
        }
        if (h->maps.size() >= 512) {  // bounded cache
            CU(cudaDeviceSynchronize());
            for (auto &kv : h->maps) cudaFree(kv.second);
            h->maps.clear();
        }
        bq::TraceMap *d = nullptr;
        CU(cudaMalloc(&d, sizeof(bq::TraceMap) * n_sets));
        cudaError_t e = cudaMemcpy(d, host.data(), sizeof(bq::TraceMap) * n_sets, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return fail_cuda(e, "upload trace map");
        }
        it = h->maps.emplace(std::move(key), d).first;
    }
    *out = it->second;
    return BQ_OK;
}

int check_method(int method)
{
    switch (method) {
    case BQ_WIGNER_CLENSHAW:
    case BQ_WIGNER_ITERATIVE:
    case BQ_WIGNER_LAGUERRE:
        return BQ_OK;
    case BQ_WIGNER_FFT:
        return fail(BQ_ERR_UNSUPPORTED, "method 'fft' is not implemented by the B200 path");
    default:
        return fail(BQ_ERR_INVALID, "method must be either 'iterative', 'laguerre', 'fft' or 'clenshaw'.");
    }
}

// Where the Wigner kernel's coefficient stream comes from: density matrices, or statevectors that the kernel's fused
// prologue traces itself (csrc/wigner_kernels.cu, fused_prologue).
struct WigSource {
    const double2 *rho = nullptr;  // [batch][N][ld]
    int64_t ld = 0, rho_stride = 0;
    const double2 *psi = nullptr;  // psi_batch statevectors; frame f = set * psi_batch + b
    int64_t psi_stride = 0;
    int psi_batch = 0;
    const bq::TraceMap *maps = nullptr;
    int n_keep = 0, n_trace = 0;
    double2 *rho_out = nullptr;  // optional, [batch][N][N]
};

// device-side core: rho or psi (device) -> W (device)
int wigner_core(bq_handle_t h, const WigSource &src, int N, int batch, const double *d_x, int nx, const double *d_y, int ny,
                double g, double *d_W, cudaStream_t st, bool mirror_grid)
{
    NvtxRange nvtx_range("K3 wigner (coefficient stream + Clenshaw launch)");
    if (batch == 0 || nx == 0 || ny == 0) return BQ_OK;
    const double4 *tab = nullptr;
    int rc = get_tab(h, N, st, &tab);
    if (rc) return rc;
    const int64_t R = bq::stream_records(N);
    CU(h->cs.reserve((size_t)R * batch * sizeof(double2)));
    int nl = 0;
    bq::WigParams p{};
    p.cs = (const double2 *)h->cs.ptr;
    p.cs_w = (double2 *)h->cs.ptr;
    if (src.psi) {
        p.pro_mode = 2;
        // tiny traces at a resident cutoff on a mirror-symmetric grid: every CTA traces its frame into its own shared memory
        // (local_prologue): no global stream, no grid-wide barrier
        const int64_t ntri = (int64_t)N * (N + 1) / 2, T = (int64_t)1 << src.n_trace;
        // (few frames only: with hundreds of frames the grid-wide prologue's barrier is noise and recomputing a frame's trace
        // in every CTA that touches it is not -- 512 frames of cutoff 32: 2.97 ms against 3.00 ms)
        // few kept, many traced indices (per-site reduced density matrices): rows of rho x slices of the traced range
        // (up to 8192 traced indices: beyond, a CTA walks several slices one after the other and the second pass adds
        // hundreds of partial sums per element -- 4 sites x cutoff 32 (32768 traced): 229 us against 145 us for the DMMA trace
        // launch + the Wigner launch; 4 sites x cutoff 16 (4096): 50 against 69 us)
        if (h->staged_trace && N <= 32 && T >= 512 && T <= 8192) {
            // slices of 32 .. 256 traced indices (128 at cutoff 32: the slice of Psi has to fit in shared memory next to
            // the kernel's own, <= 66 KB here); about one (frame, slice) per SM
            int64_t slices = std::max<int64_t>(1, T / (N <= 16 ? 256 : 128));
            while (slices * 2 * batch <= h->dev.sm_count && T / (slices * 2) >= 32) slices *= 2;
            const int64_t Tc = T / slices;
            CU(h->partial.reserve((size_t)batch * ntri * slices * sizeof(double2)));
            p.pro_mode = 4;
            p.pro_slices = (int)slices;
            p.pro_partial = (double2 *)h->partial.ptr;
            p.pro_smem_bytes = (int)((size_t)N * (Tc + 1) * sizeof(double2) + (size_t)(Tc + N) * sizeof(uint64_t));
        }
        if (h->local_prologue && mirror_grid && nx == ny && nx >= 2 && N >= 2 && N <= bq::kResidentMaxN && h->unit_count < 0 && T < 32 &&
            ntri * T <= 4096 && batch <= 8)
            p.pro_mode = 3;
        p.pro_psi = src.psi;
        p.pro_psi_stride = src.psi_stride;
        p.pro_psi_batch = src.psi_batch;
        p.pro_maps = src.maps;
        p.pro_n_keep = src.n_keep;
        p.pro_n_trace = src.n_trace;
        p.pro_rho_out = src.rho_out;
    } else if (h->fused) {
        p.pro_mode = 1;
        p.pro_rho = src.rho;
        p.pro_ld = src.ld;
        p.pro_rho_stride = src.rho_stride;
    } else {
        for (int b0 = 0; b0 < batch; b0 += 65535) {
            const int nb = std::min(65535, batch - b0);
            cudaError_t e = bq::launch_pack_rho(src.rho + (int64_t)b0 * src.rho_stride, src.ld, src.rho_stride, N, nb,
                                                (double2 *)h->cs.ptr + (int64_t)b0 * R, R, st);
            if (e != cudaSuccess) return fail_cuda(e, "pack_rho");
            ++nl;
        }
    }
    p.tab = tab;
    p.xvec = d_x;
    p.yvec = d_y;
    p.W = d_W;
    p.w_stride = (int64_t)nx * ny;
    p.cs_stride = R;
    p.N = N;
    p.R_tot = (int)R;
    p.nx = nx;
    p.ny = ny;
    p.batch = batch;
    p.g = g;
    p.scale = g * g * 0.5 / 3.14159265358979323846;
    p.sym = (mirror_grid && nx == ny && nx >= 2 && N >= 2) ? 1 : 0;
    p.sym_want_first = h->unit_first;
    p.sym_want_count = h->unit_count;
    if (h->unit_count >= 0 && !p.sym)
        return fail(BQ_ERR_UNSUPPORTED, "a unit range needs a mirror-symmetric grid (bq_set_grid_symmetry / bq_grid_is_mirror)");
    p.ctr = h->ctr + 4 * h->ctr_next;
    h->ctr_next = (h->ctr_next + 1) % kCtrSlots;
    cudaEvent_t pa = nullptr, pb = nullptr;
    if (h->profile) {
        CU(cudaEventCreate(&pa));
        CU(cudaEventCreate(&pb));
        CU(cudaEventRecord(pa, st));
    }
    cudaError_t e = bq::launch_wigner(p, h->dev, st, &nl);
    h->launches += nl;
    if (h->profile) {
        cudaEventRecord(pb, st);
        h->prof_events.emplace_back(pa, pb);
    }
    if (e == cudaErrorNotSupported) {
        cudaGetLastError();
        return fail(BQ_ERR_UNSUPPORTED, "no symmetric-grid kernel for this cutoff / grid: shard by row slab instead");
    }
    if (e != cudaSuccess) return fail_cuda(e, "launch_wigner");
    return BQ_OK;
}

int wigner_core(bq_handle_t h, const double2 *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                const double *d_x, int nx, const double *d_y, int ny, double g, double *d_W, cudaStream_t st,
                bool mirror_grid)
{
    WigSource src;
    src.rho = d_rho;
    src.ld = ld;
    src.rho_stride = rho_stride;
    return wigner_core(h, src, N, batch, d_x, nx, d_y, ny, g, d_W, st, mirror_grid);
}

int check_wigner_args(const void *rho, int N, int64_t ld, int64_t rho_stride, int batch, const void *x, int nx,
                      const void *y, int ny, double g, const void *W)
{
    if (N < 1 || N > 16384) return fail(BQ_ERR_INVALID, "cutoff N out of range [1, 16384]");
    if (batch < 0 || nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (ld < N) return fail(BQ_ERR_INVALID, "leading dimension smaller than N");
    if (batch > 1 && rho_stride < 0) return fail(BQ_ERR_INVALID, "negative rho_stride");
    if (!std::isfinite(g)) return fail(BQ_ERR_INVALID, "g is not finite");
    if (batch > 0 && nx > 0 && ny > 0 && (!rho || !x || !y || !W)) return fail(BQ_ERR_INVALID, "NULL pointer");
    return BQ_OK;
}

// uploads (xvec, yvec) unless the device copy already holds exactly these values
// yvec == xvec and xvec[S-1-i] == -xvec[i] to within 4 ulp of max|x|: the grids wigner.py:135,187-188 builds
// (numpy.linspace(-a, a, S) is symmetric to ~1 ulp, not bit-exactly).  The symmetric kernels take the mirrored
// coordinate as exactly -x, which moves a grid point by at most that much.
bool grid_is_mirror(const double *x, int nx, const double *y, int ny)
{
    if (nx != ny || nx < 2) return false;
    double amax = 0.0;
    for (int i = 0; i < nx; ++i) {
        if (!(x[i] == y[i]) || !std::isfinite(x[i])) return false;
        amax = std::max(amax, std::fabs(x[i]));
    }
    const double tol = 4.0 * 2.220446049250313e-16 * amax;
    for (int i = 0; i < nx / 2; ++i)
        if (std::fabs(x[i] + x[nx - 1 - i]) > tol) return false;
    if ((nx & 1) && std::fabs(x[nx / 2]) > tol) return false;
    return true;
}

int stage_grid(bq_handle_t h, const double *x, int nx, const double *y, int ny, const double **dx, const double **dy)
{
    const size_t n = (size_t)nx + ny;
    bool same = h->grid_host.size() == n + 2 && h->grid_host[0] == nx && h->grid_host[1] == ny &&
                std::memcmp(h->grid_host.data() + 2, x, sizeof(double) * nx) == 0 &&
                std::memcmp(h->grid_host.data() + 2 + nx, y, sizeof(double) * ny) == 0;
    if (!same) {
        CU(cudaStreamSynchronize(h->stream));
        CU(h->grid_dev.reserve(sizeof(double) * std::max<size_t>(n, 1)));
        h->grid_host.assign(n + 2, 0.0);
        h->grid_host[0] = nx;
        h->grid_host[1] = ny;
        std::memcpy(h->grid_host.data() + 2, x, sizeof(double) * nx);
        std::memcpy(h->grid_host.data() + 2 + nx, y, sizeof(double) * ny);
        if (n) CU(cudaMemcpy(h->grid_dev.ptr, h->grid_host.data() + 2, sizeof(double) * n, cudaMemcpyHostToDevice));
        h->grid_host_mirror = grid_is_mirror(x, nx, y, ny);
    }
    *dx = (const double *)h->grid_dev.ptr;
    *dy = (const double *)h->grid_dev.ptr + nx;
    return BQ_OK;
}

int trace_sv_core(bq_handle_t h, const double2 *d_psi, int n_qubits, const int *traced, int n_traced, int n_sets,
                  int batch, int64_t psi_stride, double2 *d_rho, cudaStream_t st)
{
    NvtxRange nvtx_range("K1 partial trace (DMMA)");
    const bq::TraceMap *maps = nullptr;
    int rc = get_maps(h, n_qubits, traced, n_traced, n_sets, &maps);
    if (rc) return rc;
    const int n_keep = n_qubits - n_traced;
    if (n_keep > 15) return fail(BQ_ERR_INVALID, "reduced density matrix larger than 2^15 x 2^15");
    const int64_t D = (int64_t)1 << n_keep;
    const int zmax = std::max(1, 65535 / n_sets);
    for (int b0 = 0; b0 < batch; b0 += zmax) {
        const int nb = std::min(zmax, batch - b0);
        CU(h->partial.reserve(bq::trace_sv_scratch_bytes(n_keep, n_traced, nb, n_sets)));
        int nl = 0;
        // output layout [set][batch][D][D]: with several sets and a sliced batch the slices interleave,
        // so slicing is only done for the single-set case (multi-set calls are batch == 1)
        cudaError_t e = bq::launch_trace_sv(d_psi + (int64_t)b0 * psi_stride, psi_stride, nb, maps, n_sets, n_keep,
                                            n_traced, d_rho + (int64_t)b0 * D * D, (double2 *)h->partial.ptr, h->dev,
                                            st, &nl);
        h->launches += nl;
        if (e != cudaSuccess) return fail_cuda(e, "launch_trace_sv");
    }
    return BQ_OK;
}

int check_trace_args(const void *psi, int n_qubits, int n_traced, int batch, int64_t psi_stride, const void *out)
{
    if (n_qubits < 0 || n_qubits > bq::kMaxQubits) return fail(BQ_ERR_INVALID, "n_qubits out of range [0, 40]");
    if (n_traced < 0 || n_traced > n_qubits) return fail(BQ_ERR_INVALID, "n_traced out of range");
    if (batch < 0) return fail(BQ_ERR_INVALID, "negative batch");
    if (batch > 1 && psi_stride < ((int64_t)1 << n_qubits)) return fail(BQ_ERR_INVALID, "psi_stride smaller than 2^n_qubits");
    if (batch > 0 && (!psi || !out)) return fail(BQ_ERR_INVALID, "NULL pointer");
    return BQ_OK;
}

// If `p` is page-locked host memory that the device can address (cudaHostAlloc / cudaHostRegister under
// UVA), returns the device-side alias: kernels then store results straight into host memory over PCIe
// (posted writes that overlap the fp64 math) and the separate D2H copy disappears.
double *mapped_alias(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (attr.type == cudaMemoryTypeHost && attr.devicePointer) return (double *)attr.devicePointer;
    return nullptr;
}

struct HostTimer {
    bq_handle_t h;
    explicit HostTimer(bq_handle_t hh) : h(hh) { cudaEventRecord(h->ev_t0, h->stream); }
    int finish()
    {
        CU(cudaEventRecord(h->ev_t1, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
        h->last_ms = ms;
        return BQ_OK;
    }
};

}  // namespace

// lets the other translation units of the library (bq_multi.cu) set the calling thread's bq_last_error()
extern "C" __attribute__((visibility("hidden"))) int bq_internal_fail(int code, const char *msg) { return fail(code, msg ? msg : ""); }

extern "C" {

int bq_version(void) { return BQ_B200_VERSION; }
const char *bq_last_error(void) { return g_err.c_str(); }

int bq_device_count(int *count)
{
    if (!count) return fail(BQ_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail_cuda(e, "cudaGetDeviceCount");
    }
    *count = n;
    return BQ_OK;
}

int bq_create(bq_handle_t *out, int device)
{
    if (!out) return fail(BQ_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(BQ_ERR_NO_DEVICE, std::string("no CUDA device available (there is no CPU fallback): ") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(BQ_ERR_NO_DEVICE, "device ordinal out of range");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(BQ_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(BQ_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200) class");
    bq_handle_s *h = new bq_handle_s();
    h->dev.device = device;
    h->dev.sm_count = prop.multiProcessorCount;
    h->dev.smem_optin = prop.sharedMemPerBlockOptin;
    {
        const char *p8 = std::getenv("BQ_B200_PTS8");
        h->dev.pts8 = (p8 && p8[0] == '0') ? 0 : 1;  // default on: measured 87.6 % vs 82.7 % of fp64 peak at N=64
        const char *sv = std::getenv("BQ_B200_SYM_VARIANT");
        h->dev.sym_force = sv ? std::atoi(sv) : -1;
        if (const char *sp = std::getenv("BQ_B200_SYM_SPLIT")) std::sscanf(sp, "%d:%d", &h->dev.sym_split_main, &h->dev.sym_split_tail);
        if (const char *ph = std::getenv("BQ_B200_SYM_PHASE")) h->dev.sym_phase = std::atoi(ph);
        const char *fu = std::getenv("BQ_B200_FUSED");
        h->fused = !(fu && fu[0] == '0');
        const char *lp = std::getenv("BQ_B200_LOCAL_PROLOGUE");
        h->local_prologue = !(lp && lp[0] == '0');
        const char *stg = std::getenv("BQ_B200_STAGED_TRACE");
        h->staged_trace = !(stg && stg[0] == '0');
        const char *env = std::getenv("BQ_B200_SASS");
        h->dev.sass_mode = (env && (std::strcmp(env, "plain") == 0 || std::strcmp(env, "0") == 0)) ? 0 : (bq::patched_sass_available() ? 1 : 0);
    }
    if (bq::patched_sass_available()) {
        // load-time self-check of the re-scheduled SASS, once per device and process (BQ_B200_SASS_VERIFY=0 skips it)
        static std::mutex vmu;
        static std::map<int, std::pair<uint64_t, int>> verdicts;
        std::lock_guard<std::mutex> vlk(vmu);
        auto it = verdicts.find(device);
        if (it == verdicts.end()) {
            uint64_t bad = 0;
            int checked = 0;
            const char *vf = std::getenv("BQ_B200_SASS_VERIFY");
            if (!(vf && vf[0] == '0')) {
                cudaError_t ve = bq::verify_patched_kernels(h->dev, &bad, &checked);
                if (ve != cudaSuccess) {
                    cudaGetLastError();
                    bad = ~0ull;  // could not compare: nvcc's schedule for everything
                    std::fprintf(stderr, "libbq_b200: SASS self-check could not run (%s); using nvcc's schedule\n", cudaGetErrorString(ve));
                } else if (bad) {
                    std::fprintf(stderr, "libbq_b200: re-scheduled SASS of kernel variants 0x%llx differs from nvcc's schedule; "
                                         "those variants run nvcc's schedule\n", (unsigned long long)bad);
                }
            }
            it = verdicts.emplace(device, std::make_pair(bad, checked)).first;
        }
        h->dev.patched_bad = it->second.first;
        h->sass_checked = it->second.second;
    }
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_block[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_block[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_block[2], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_block[3], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreate(&h->ev_t0) == cudaSuccess && cudaEventCreate(&h->ev_t1) == cudaSuccess &&
              cudaMalloc(&h->ctr, sizeof(unsigned int) * 4 * kCtrSlots) == cudaSuccess &&
              cudaMemset(h->ctr, 0, sizeof(unsigned int) * 4 * kCtrSlots) == cudaSuccess;
    if (!ok) {
        cudaError_t le = cudaGetLastError();
        bq_destroy(h);
        return fail_cuda(le, "bq_create resources");
    }
    *out = h;
    return BQ_OK;
}

int bq_destroy(bq_handle_t h)
{
    if (!h) return BQ_OK;
    {
        DeviceGuard guard(h->dev.device);
        cudaDeviceSynchronize();
        for (auto &kv : h->tabs) cudaFree(kv.second);
        for (auto &kv : h->maps) cudaFree(kv.second);
        Buffer *bufs[] = {&h->cs, &h->partial, &h->rho_tmp, &h->in_dev, &h->out_dev, &h->rho_dev, &h->grid_dev, &h->small_dev,
                          &h->fft_a, &h->fft_m, &h->fft_r, &h->fft_t, &h->fft_x, &h->mm_dev};
        for (Buffer *b : bufs) b->release();
        for (auto &pr : h->prof_events) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        if (h->ctr) cudaFree(h->ctr);
        if (h->ev_order) cudaEventDestroy(h->ev_order);
        if (h->ev_t0) cudaEventDestroy(h->ev_t0);
        if (h->ev_t1) cudaEventDestroy(h->ev_t1);
        for (cudaEvent_t e : h->ev_block)
            if (e) cudaEventDestroy(e);
        if (h->ev_copied) cudaEventDestroy(h->ev_copied);
        if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
    return BQ_OK;
}

int bq_device(bq_handle_t h, int *device)
{
    if (!h || !device) return fail(BQ_ERR_INVALID, "NULL argument");
    *device = h->dev.device;
    return BQ_OK;
}

int bq_launch_count(bq_handle_t h, uint64_t *count)
{
    if (!h || !count) return fail(BQ_ERR_INVALID, "NULL argument");
    *count = h->launches;
    return BQ_OK;
}

int bq_synchronize(bq_handle_t h)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    DeviceGuard guard(h->dev.device);
    CU(cudaStreamSynchronize(h->stream));
    if (h->have_last) CU(cudaStreamSynchronize(h->last_stream));
    return BQ_OK;
}

int bq_host_alloc(bq_handle_t h, size_t bytes, void **out)
{
    if (!h || !out) return fail(BQ_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->dev.device);
    CU(cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable));
    return BQ_OK;
}

int bq_host_free(bq_handle_t h, void *ptr)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (!ptr) return BQ_OK;
    DeviceGuard guard(h->dev.device);
    CU(cudaFreeHost(ptr));
    return BQ_OK;
}

// ---- partial trace -----------------------------------------------------------------------------
int bq_partial_trace_sv_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                            int batch, int64_t psi_stride, double *d_rho_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_partial_trace_sv_dev");
    int rc = check_trace_args(d_psi, n_qubits, n_traced, batch, psi_stride, d_rho_out);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    return trace_sv_core(h, (const double2 *)d_psi, n_qubits, traced, n_traced, 1, batch, psi_stride,
                         (double2 *)d_rho_out, st);
}

int bq_partial_trace_sv_multi_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced,
                                  int n_traced, int n_sets, double *d_rho_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_partial_trace_sv_multi_dev");
    int rc = check_trace_args(d_psi, n_qubits, n_traced, 1, 0, d_rho_out);
    if (rc) return rc;
    if (n_sets < 0 || n_sets > 4096) return fail(BQ_ERR_INVALID, "n_sets out of range [0, 4096]");
    if (n_sets == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    return trace_sv_core(h, (const double2 *)d_psi, n_qubits, traced, n_traced, n_sets, 1, 0, (double2 *)d_rho_out, st);
}

static int trace_sv_host_common(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                int n_sets, int batch, int64_t psi_stride, double *rho_out)
{
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    int rc = order_stream(h, h->stream);
    if (rc) return rc;
    const int64_t dim = (int64_t)1 << n_qubits;
    const int64_t D = (int64_t)1 << (n_qubits - n_traced);
    const size_t in_elems = batch ? (size_t)((int64_t)(batch - 1) * psi_stride + dim) : 0;
    const size_t out_elems = (size_t)n_sets * batch * D * D;
    if (!in_elems || !out_elems) return BQ_OK;
    CU(h->in_dev.reserve(in_elems * sizeof(double2)));
    CU(h->rho_dev.reserve(out_elems * sizeof(double2)));
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, psi, in_elems * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    rc = trace_sv_core(h, (const double2 *)h->in_dev.ptr, n_qubits, traced, n_traced, n_sets, batch, psi_stride,
                       (double2 *)h->rho_dev.ptr, h->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(rho_out, h->rho_dev.ptr, out_elems * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

int bq_partial_trace_sv_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                             int batch, int64_t psi_stride, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    int rc = check_trace_args(psi, n_qubits, n_traced, batch, psi_stride, rho_out);
    if (rc) return rc;
    return trace_sv_host_common(h, psi, n_qubits, traced, n_traced, 1, batch, psi_stride, rho_out);
}

int bq_partial_trace_sv_multi_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced,
                                   int n_traced, int n_sets, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    int rc = check_trace_args(psi, n_qubits, n_traced, 1, 0, rho_out);
    if (rc) return rc;
    if (n_sets < 0 || n_sets > 4096) return fail(BQ_ERR_INVALID, "n_sets out of range [0, 4096]");
    if (n_sets == 0) return BQ_OK;
    return trace_sv_host_common(h, psi, n_qubits, traced, n_traced, n_sets, 1, 0, rho_out);
}

// ---- shot-averaged partial trace ---------------------------------------------------------------
static int trace_accumulate_core(bq_handle_t h, const double2 *d_psi, int n_qubits, const int *traced, int n_traced, int shots,
                                 int64_t psi_stride, const double *d_weights, double2 *d_rho, cudaStream_t st)
{
    NvtxRange nvtx_range("K1 partial trace, shot average");
    const bq::TraceMap *maps = nullptr;
    int rc = get_maps(h, n_qubits, traced, n_traced, 1, &maps);
    if (rc) return rc;
    const int n_keep = n_qubits - n_traced;
    if (n_keep > 15) return fail(BQ_ERR_INVALID, "reduced density matrix larger than 2^15 x 2^15");
    if (shots > 65535) return fail(BQ_ERR_INVALID, "more than 65535 shots in one call");
    CU(h->partial.reserve(bq::trace_sv_scratch_bytes(n_keep, n_traced, shots, 1)));
    int nl = 0;
    cudaError_t e = bq::launch_trace_sv_accumulate(d_psi, psi_stride, shots, maps, 1, n_keep, n_traced, d_weights, d_rho,
                                                   (double2 *)h->partial.ptr, h->dev, st, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_trace_sv_accumulate");
    return BQ_OK;
}

int bq_partial_trace_sv_accumulate_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                                       int shots, int64_t psi_stride, const double *d_weights, double *d_rho_out,
                                       void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    int rc = check_trace_args(d_psi, n_qubits, n_traced, shots, psi_stride, d_rho_out);
    if (rc) return rc;
    if (shots == 0) return fail(BQ_ERR_INVALID, "the average of zero shots is undefined");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    NvtxRange range("bq_partial_trace_sv_accumulate_dev");
    return trace_accumulate_core(h, (const double2 *)d_psi, n_qubits, traced, n_traced, shots, psi_stride, d_weights,
                                 (double2 *)d_rho_out, st);
}

int bq_partial_trace_sv_accumulate_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                        int shots, int64_t psi_stride, const double *weights, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    int rc = check_trace_args(psi, n_qubits, n_traced, shots, psi_stride, rho_out);
    if (rc) return rc;
    if (shots == 0) return fail(BQ_ERR_INVALID, "the average of zero shots is undefined");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    if ((rc = order_stream(h, h->stream))) return rc;
    const int64_t dim = (int64_t)1 << n_qubits, D = (int64_t)1 << (n_qubits - n_traced);
    const size_t in_elems = (size_t)((int64_t)(shots - 1) * psi_stride + dim);
    CU(h->in_dev.reserve(in_elems * sizeof(double2)));
    CU(h->rho_dev.reserve((size_t)D * D * sizeof(double2)));
    if (weights) CU(h->small_dev.reserve((size_t)shots * sizeof(double)));
    NvtxRange range("bq_partial_trace_sv_accumulate_host");
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, psi, in_elems * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    if (weights) CU(cudaMemcpyAsync(h->small_dev.ptr, weights, (size_t)shots * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = trace_accumulate_core(h, (const double2 *)h->in_dev.ptr, n_qubits, traced, n_traced, shots, psi_stride,
                               weights ? (const double *)h->small_dev.ptr : nullptr, (double2 *)h->rho_dev.ptr, h->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(rho_out, h->rho_dev.ptr, (size_t)D * D * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

static int trace_dm_core(bq_handle_t h, const double2 *d_rho, int n_qubits, const int *traced, int n_traced, int batch,
                         double2 *d_out, cudaStream_t st)
{
    NvtxRange nvtx_range("K2 partial trace of a density matrix");
    const bq::TraceMap *maps = nullptr;
    int rc = get_maps(h, n_qubits, traced, n_traced, 1, &maps);
    if (rc) return rc;
    int nl = 0;
    cudaError_t e = bq::launch_trace_dm(d_rho, n_qubits, batch, maps, n_qubits - n_traced, n_traced, d_out, st, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_trace_dm");
    return BQ_OK;
}

static int check_dm_args(const void *rho, int n_qubits, int n_traced, int batch, const void *out)
{
    if (n_qubits < 0 || n_qubits > 15) return fail(BQ_ERR_INVALID, "density matrix n_qubits out of range [0, 15]");
    if (n_traced < 0 || n_traced > n_qubits) return fail(BQ_ERR_INVALID, "n_traced out of range");
    if (batch < 0) return fail(BQ_ERR_INVALID, "negative batch");
    if (batch > 0 && (!rho || !out)) return fail(BQ_ERR_INVALID, "NULL pointer");
    return BQ_OK;
}

int bq_partial_trace_dm_dev(bq_handle_t h, const double *d_rho, int n_qubits, const int *traced, int n_traced,
                            int batch, double *d_rho_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_partial_trace_dm_dev");
    int rc = check_dm_args(d_rho, n_qubits, n_traced, batch, d_rho_out);
    if (rc) return rc;
    if (batch == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    return trace_dm_core(h, (const double2 *)d_rho, n_qubits, traced, n_traced, batch, (double2 *)d_rho_out, st);
}

int bq_partial_trace_dm_host(bq_handle_t h, const double *rho, int n_qubits, const int *traced, int n_traced,
                             int batch, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_partial_trace_dm_host");
    int rc = check_dm_args(rho, n_qubits, n_traced, batch, rho_out);
    if (rc) return rc;
    if (batch == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    if ((rc = order_stream(h, h->stream))) return rc;
    const int64_t F = (int64_t)1 << n_qubits, D = (int64_t)1 << (n_qubits - n_traced);
    const size_t in_b = (size_t)batch * F * F * sizeof(double2), out_b = (size_t)batch * D * D * sizeof(double2);
    CU(h->in_dev.reserve(in_b));
    CU(h->rho_dev.reserve(out_b));
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, rho, in_b, cudaMemcpyHostToDevice, h->stream));
    rc = trace_dm_core(h, (const double2 *)h->in_dev.ptr, n_qubits, traced, n_traced, batch, (double2 *)h->rho_dev.ptr,
                       h->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(rho_out, h->rho_dev.ptr, out_b, cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

int bq_photon_number_dev(bq_handle_t h, const double *d_rho, int D, int batch, double *d_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_photon_number_dev");
    if (D < 1 || batch < 0) return fail(BQ_ERR_INVALID, "bad size");
    if (batch == 0) return BQ_OK;
    if (!d_rho || !d_out) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = order_stream(h, st);
    if (rc) return rc;
    int nl = 0;
    cudaError_t e = bq::launch_photon_number((const double2 *)d_rho, D, batch, (double2 *)d_out, st, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_photon_number");
    return BQ_OK;
}

int bq_photon_number_host(bq_handle_t h, const double *rho, int D, int batch, double *out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_photon_number_host");
    if (D < 1 || batch < 0) return fail(BQ_ERR_INVALID, "bad size");
    if (batch == 0) return BQ_OK;
    if (!rho || !out) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    int rc = order_stream(h, h->stream);
    if (rc) return rc;
    const size_t in_b = (size_t)batch * D * D * sizeof(double2), out_b = (size_t)batch * sizeof(double2);
    CU(h->in_dev.reserve(in_b));
    CU(h->small_dev.reserve(out_b));
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, rho, in_b, cudaMemcpyHostToDevice, h->stream));
    int nl = 0;
    cudaError_t e = bq::launch_photon_number((const double2 *)h->in_dev.ptr, D, batch, (double2 *)h->small_dev.ptr,
                                             h->stream, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_photon_number");
    CU(cudaMemcpyAsync(out, h->small_dev.ptr, out_b, cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

// ---- Wigner ------------------------------------------------------------------------------------
int bq_wigner_dev(bq_handle_t h, const double *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                  const double *d_xvec, int nx, const double *d_yvec, int ny, double g, int method,
                  double *d_W_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_wigner_dev");
    int rc = check_method(method);
    if (rc) return rc;
    if ((rc = check_wigner_args(d_rho, N, ld, rho_stride, batch, d_xvec, nx, d_yvec, ny, g, d_W_out))) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    return wigner_core(h, (const double2 *)d_rho, N, ld, rho_stride, batch, d_xvec, nx, d_yvec, ny, g, d_W_out, st,
                       h->grid_mode == 2);
}

int bq_wigner_host(bq_handle_t h, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                   const double *xvec, int nx, const double *yvec, int ny, double g, int method, double *W_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_wigner_host");
    int rc = check_method(method);
    if (rc) return rc;
    if ((rc = check_wigner_args(rho, N, ld, rho_stride, batch, xvec, nx, yvec, ny, g, W_out))) return rc;
    if (batch == 0 || nx == 0 || ny == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    if ((rc = order_stream(h, h->stream))) return rc;
    const double *dx, *dy;
    if ((rc = stage_grid(h, xvec, nx, yvec, ny, &dx, &dy))) return rc;
    const size_t in_elems = (size_t)((int64_t)(batch - 1) * rho_stride + (int64_t)(N - 1) * ld + N);
    const size_t out_b = (size_t)batch * nx * ny * sizeof(double);
    CU(h->in_dev.reserve(in_elems * sizeof(double2)));
    // the symmetric kernels scatter 8-byte stores over the grid: fine for HBM/L2, not for posted writes over PCIe
    const bool mirror = h->grid_mode >= 1 && h->grid_host_mirror;
    double *d_out = mirror ? nullptr : mapped_alias(W_out);
    const bool direct = d_out != nullptr;
    if (!direct) {
        CU(h->out_dev.reserve(out_b));
        d_out = (double *)h->out_dev.ptr;
    }
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, rho, in_elems * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    rc = wigner_core(h, (const double2 *)h->in_dev.ptr, N, ld, rho_stride, batch, dx, nx, dy, ny, g, d_out, h->stream,
                     mirror);
    if (rc) return rc;
    if (!direct) CU(cudaMemcpyAsync(W_out, h->out_dev.ptr, out_b, cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

// ---- method="fft" ------------------------------------------------------------------------------
namespace {

// p axis of qutip's _wigner_fft / _psi_wigner_fft, same operation order as the NumPy expressions:
//   xs = xvec * g / sqrt(2);  p = arange(-n/4, n/4) * pi / (n * (xs[1] - xs[0]));  yvec = p * sqrt(2) / g
int fft_axis(const double *xvec, int nx, double g, double *yvec_out, double *scale)
{
    const double n = 2.0 * nx;
    const double dx = xvec[1] * g / std::sqrt(2.0) - xvec[0] * g / std::sqrt(2.0);
    if (!(dx != 0.0) || !std::isfinite(dx)) return fail(BQ_ERR_INVALID, "method='fft' needs a uniformly spaced xvec with xvec[1] != xvec[0]");
    auto p_at = [&](int c) { return (-n / 4.0 + c) * 3.14159265358979323846 / (n * dx); };
    if (yvec_out)
        for (int c = 0; c < nx; ++c) yvec_out[c] = p_at(c) * std::sqrt(2.0) / g;
    *scale = (0.5 * g * g) / (p_at(1) - p_at(0)) / n;
    return BQ_OK;
}

int wigner_fft_core(bq_handle_t h, const double2 *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                    const double *xvec, int nx, double g, double *d_W, double *yvec_out, cudaStream_t st)
{
    NvtxRange nvtx_range("K6 wigner method=fft");
    if (N < 1 || N > 16384) return fail(BQ_ERR_INVALID, "cutoff N out of range [1, 16384]");
    if (batch < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (nx < 2) return fail(BQ_ERR_INVALID, "method='fft' needs at least two grid points (the p axis comes from xvec[1] - xvec[0])");
    if (nx > 14000) return fail(BQ_ERR_UNSUPPORTED, "method='fft': grids above 14000 points per axis do not fit the band buffer");
    if (ld < N) return fail(BQ_ERR_INVALID, "leading dimension smaller than N");
    if (!std::isfinite(g) || g == 0.0) return fail(BQ_ERR_INVALID, "g must be finite and non-zero");
    if (!xvec || (batch > 0 && (!d_rho || !d_W))) return fail(BQ_ERR_INVALID, "NULL pointer");
    double scale = 0.0;
    int rc = fft_axis(xvec, nx, g, yvec_out, &scale);
    if (rc || batch == 0) return rc;
    const size_t S = (size_t)nx;
    CU(h->fft_x.reserve(S * sizeof(double)));
    CU(h->fft_a.reserve((size_t)N * S * sizeof(double)));
    CU(h->fft_m.reserve((size_t)N * S * sizeof(double2)));
    CU(h->fft_r.reserve(S * S * sizeof(double2)));
    CU(h->fft_t.reserve(2 * S * sizeof(double2)));
    CU(cudaMemcpyAsync(h->fft_x.ptr, xvec, S * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // xvec is the caller's (possibly pageable) memory: do not return before it is read
    for (int b = 0; b < batch; ++b) {
        int nl = 0;
        cudaError_t e = bq::launch_wigner_fft(d_rho + (int64_t)b * rho_stride, N, ld, (const double *)h->fft_x.ptr, nx, g, scale,
                                              (double *)h->fft_a.ptr, (double2 *)h->fft_m.ptr, (double2 *)h->fft_r.ptr,
                                              (double2 *)h->fft_t.ptr, d_W + (int64_t)b * nx * nx, st, &nl);
        h->launches += nl;
        if (e != cudaSuccess) return fail_cuda(e, "launch_wigner_fft");
    }
    return BQ_OK;
}

}  // namespace

int bq_wigner_fft_dev(bq_handle_t h, const double *d_rho, int N, int64_t ld, int64_t rho_stride, int batch,
                      const double *xvec, int nx, double g, double *d_W_out, double *yvec_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_wigner_fft_dev");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = order_stream(h, st);
    if (rc) return rc;
    return wigner_fft_core(h, (const double2 *)d_rho, N, ld, rho_stride, batch, xvec, nx, g, d_W_out, yvec_out, st);
}

int bq_wigner_fft_host(bq_handle_t h, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                       const double *xvec, int nx, double g, double *W_out, double *yvec_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_wigner_fft_host");
    if (N < 1 || ld < N || batch < 0 || nx < 0) return fail(BQ_ERR_INVALID, "bad size");
    if (batch > 0 && (!rho || !W_out)) return fail(BQ_ERR_INVALID, "NULL pointer");
    if (batch > 1 && rho_stride < 0) return fail(BQ_ERR_INVALID, "negative rho_stride");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    int rc = order_stream(h, h->stream);
    if (rc) return rc;
    const size_t in_elems = batch ? (size_t)((int64_t)(batch - 1) * rho_stride + (int64_t)(N - 1) * ld + N) : 0;
    const size_t out_b = (size_t)batch * nx * nx * sizeof(double);
    CU(h->in_dev.reserve(std::max<size_t>(in_elems, 1) * sizeof(double2)));
    CU(h->out_dev.reserve(std::max<size_t>(out_b, 8)));
    HostTimer timer(h);
    if (in_elems) CU(cudaMemcpyAsync(h->in_dev.ptr, rho, in_elems * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    rc = wigner_fft_core(h, (const double2 *)h->in_dev.ptr, N, ld, rho_stride, batch, xvec, nx, g, (double *)h->out_dev.ptr,
                         yvec_out, h->stream);
    if (rc) return rc;
    if (out_b) CU(cudaMemcpyAsync(W_out, h->out_dev.ptr, out_b, cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

// ---- fused frame loop --------------------------------------------------------------------------
// psi -> (rho) -> W for n_sets traces of each of psi_batch statevectors (frame f = set * psi_batch + b).  Small traces
// are done by the Wigner launch itself (fused prologue, one launch per call); large ones by the DMMA kernel first.
static bool trace_fits_prologue(const bq_handle_t h, int n_keep, int n_traced, int64_t frames)
{
    if (!h->fused || n_keep > 12) return false;
    const int64_t D = (int64_t)1 << n_keep, T = (int64_t)1 << n_traced;
    if (h->staged_trace && D <= 32 && T >= 512 && T <= 8192 && frames <= 256) return true;  // slices of Psi through shared memory (pro_mode 4)
    const int64_t items = frames * (D * (D + 1) / 2);
    // serial length of the prologue per warp (thread) on a conservatively small grid of 512 warps
    const int64_t est = T < 32 ? ((items + 16383) / 16384) * T : ((items + 511) / 512) * (T / 32);
    return est <= 256;
}

static int state_to_wigner_core(bq_handle_t h, const double2 *d_psi, int n_qubits, const int *traced, int n_traced,
                                int n_sets, int batch, int64_t psi_stride, const double *dx, int nx, const double *dy,
                                int ny, double g, double *d_W, double2 *d_rho_opt, cudaStream_t st, bool mirror_grid)
{
    const int n_keep = n_qubits - n_traced;
    if (n_keep > 14) return fail(BQ_ERR_INVALID, "reduced density matrix larger than 2^14 x 2^14");
    const int64_t D = (int64_t)1 << n_keep;
    const int64_t frames = (int64_t)batch * n_sets;
    if (frames > 0x7fffffffLL) return fail(BQ_ERR_INVALID, "too many frames");
    if (trace_fits_prologue(h, n_keep, n_traced, frames)) {
        WigSource src;
        int rc = get_maps(h, n_qubits, traced, n_traced, n_sets, &src.maps);
        if (rc) return rc;
        src.psi = d_psi;
        src.psi_stride = psi_stride;
        src.psi_batch = batch;
        src.n_keep = n_keep;
        src.n_trace = n_traced;
        src.rho_out = d_rho_opt;
        return wigner_core(h, src, (int)D, (int)frames, dx, nx, dy, ny, g, d_W, st, mirror_grid);
    }
    double2 *d_rho = d_rho_opt;
    if (!d_rho) {
        CU(h->rho_tmp.reserve((size_t)frames * D * D * sizeof(double2)));
        d_rho = (double2 *)h->rho_tmp.ptr;
    }
    int rc = trace_sv_core(h, d_psi, n_qubits, traced, n_traced, n_sets, batch, psi_stride, d_rho, st);
    if (rc) return rc;
    return wigner_core(h, d_rho, (int)D, D, D * D, (int)frames, dx, nx, dy, ny, g, d_W, st, mirror_grid);
}

int bq_state_to_wigner_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                           int batch, int64_t psi_stride, const double *d_xvec, int nx, const double *d_yvec,
                           int ny, double g, int method, double *d_W_out, double *d_rho_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_state_to_wigner_dev");
    int rc = check_method(method);
    if (rc) return rc;
    if ((rc = check_trace_args(d_psi, n_qubits, n_traced, batch, psi_stride, d_W_out))) return rc;
    if (nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (batch > 0 && nx > 0 && ny > 0 && (!d_xvec || !d_yvec)) return fail(BQ_ERR_INVALID, "NULL pointer");
    if (!std::isfinite(g)) return fail(BQ_ERR_INVALID, "g is not finite");
    if (batch == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    return state_to_wigner_core(h, (const double2 *)d_psi, n_qubits, traced, n_traced, 1, batch, psi_stride, d_xvec, nx,
                                d_yvec, ny, g, d_W_out, (double2 *)d_rho_out, st, h->grid_mode == 2);
}

// host pointers in, host pointers out; n_sets traces of each statevector (n_sets > 1 only with batch == 1)
static int state_to_wigner_host_impl(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                     int n_sets, int batch, int64_t psi_stride, const double *xvec, int nx,
                                     const double *yvec, int ny, double g, int method, double *W_out, double *rho_out)
{
    int rc = check_method(method);
    if (rc) return rc;
    if ((rc = check_trace_args(psi, n_qubits, n_traced, batch, psi_stride, W_out))) return rc;
    if (nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (batch > 0 && nx > 0 && ny > 0 && (!xvec || !yvec)) return fail(BQ_ERR_INVALID, "NULL pointer");
    if (!std::isfinite(g)) return fail(BQ_ERR_INVALID, "g is not finite");
    if (batch == 0 || n_sets == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    if ((rc = order_stream(h, h->stream))) return rc;
    const double *dx, *dy;
    if ((rc = stage_grid(h, xvec, nx, yvec, ny, &dx, &dy))) return rc;
    const int64_t dim = (int64_t)1 << n_qubits, D = (int64_t)1 << (n_qubits - n_traced);
    const int64_t frames = (int64_t)batch * n_sets;
    const size_t in_elems = (size_t)((int64_t)(batch - 1) * psi_stride + dim);
    const size_t out_b = (size_t)frames * nx * ny * sizeof(double);
    const size_t rho_b = (size_t)frames * D * D * sizeof(double2);
    CU(h->in_dev.reserve(in_elems * sizeof(double2)));
    const bool mirror = h->grid_mode >= 1 && h->grid_host_mirror;
    double *d_out = (out_b && !mirror) ? mapped_alias(W_out) : nullptr;
    const bool direct = d_out != nullptr;
    if (!direct) {
        CU(h->out_dev.reserve(std::max<size_t>(out_b, 8)));
        d_out = (double *)h->out_dev.ptr;
    }
    if (rho_out) CU(h->rho_dev.reserve(rho_b));
    double2 *d_rho = rho_out ? (double2 *)h->rho_dev.ptr : nullptr;
    NvtxRange range("bq_state_to_wigner_host");
    HostTimer timer(h);
    CU(cudaMemcpyAsync(h->in_dev.ptr, psi, in_elems * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    // Many frames, staged output (symmetric-grid kernels): the frames go in up to 4 blocks and a finished block's
    // device -> host copy runs on a second stream under the next block's kernel -- the copy (PCIe) is the longer leg
    // of such a call, so what is hidden is the compute.  Blocks keep >= 2 whole rounds of resident CTAs each.
    int blocks = 1;
    if (!direct && n_sets == 1 && out_b >= ((size_t)8 << 20) && batch >= 8 && nx == ny) {
        const int64_t U = (nx + 1) / 2, units = (int64_t)batch * U * (U + 1) / 2;
        const int64_t round = (int64_t)h->dev.sm_count * 768;  // units one round of the 3-unit tile covers
        blocks = (int)std::max<int64_t>(1, std::min<int64_t>({4, units / (2 * round), (int64_t)batch / 2}));
    }
    if (blocks > 1 && n_qubits - n_traced <= 14) {
        const int64_t frame_pts = (int64_t)nx * ny;
        for (int k = 0; k < blocks; ++k) {
            const int b0 = (int)((int64_t)batch * k / blocks), b1 = (int)((int64_t)batch * (k + 1) / blocks);
            rc = state_to_wigner_core(h, (const double2 *)h->in_dev.ptr + (int64_t)b0 * psi_stride, n_qubits, traced, n_traced, 1,
                                      b1 - b0, psi_stride, dx, nx, dy, ny, g, d_out + b0 * frame_pts,
                                      d_rho ? d_rho + (int64_t)b0 * D * D : nullptr, h->stream, mirror);
            if (rc) return rc;
            CU(cudaEventRecord(h->ev_block[k], h->stream));
            CU(cudaStreamWaitEvent(h->copy_stream, h->ev_block[k], 0));
            CU(cudaMemcpyAsync(W_out + b0 * frame_pts, d_out + b0 * frame_pts, (size_t)(b1 - b0) * frame_pts * sizeof(double),
                               cudaMemcpyDeviceToHost, h->copy_stream));
        }
        if (rho_out) CU(cudaMemcpyAsync(rho_out, h->rho_dev.ptr, rho_b, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->ev_copied, h->copy_stream));
        CU(cudaStreamWaitEvent(h->stream, h->ev_copied, 0));
        return timer.finish();
    }
    rc = state_to_wigner_core(h, (const double2 *)h->in_dev.ptr, n_qubits, traced, n_traced, n_sets, batch, psi_stride, dx, nx,
                              dy, ny, g, d_out, d_rho, h->stream, mirror);
    if (rc) return rc;
    if (rho_out) CU(cudaMemcpyAsync(rho_out, h->rho_dev.ptr, rho_b, cudaMemcpyDeviceToHost, h->stream));
    if (out_b && !direct) CU(cudaMemcpyAsync(W_out, h->out_dev.ptr, out_b, cudaMemcpyDeviceToHost, h->stream));
    return timer.finish();
}

int bq_state_to_wigner_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                            int batch, int64_t psi_stride, const double *xvec, int nx, const double *yvec,
                            int ny, double g, int method, double *W_out, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    return state_to_wigner_host_impl(h, psi, n_qubits, traced, n_traced, 1, batch, psi_stride, xvec, nx, yvec, ny, g, method,
                                     W_out, rho_out);
}

int bq_state_to_wigner_multi_host(bq_handle_t h, const double *psi, int n_qubits, const int *traced, int n_traced,
                                  int n_sets, const double *xvec, int nx, const double *yvec, int ny, double g,
                                  int method, double *W_out, double *rho_out)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (n_sets < 0 || n_sets > 4096) return fail(BQ_ERR_INVALID, "n_sets out of range [0, 4096]");
    return state_to_wigner_host_impl(h, psi, n_qubits, traced, n_traced, n_sets, 1, 0, xvec, nx, yvec, ny, g, method, W_out,
                                     rho_out);
}

int bq_state_to_wigner_multi_dev(bq_handle_t h, const double *d_psi, int n_qubits, const int *traced, int n_traced,
                                 int n_sets, const double *d_xvec, int nx, const double *d_yvec, int ny, double g,
                                 int method, double *d_W_out, double *d_rho_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    int rc = check_method(method);
    if (rc) return rc;
    if ((rc = check_trace_args(d_psi, n_qubits, n_traced, 1, 0, d_W_out))) return rc;
    if (n_sets < 0 || n_sets > 4096) return fail(BQ_ERR_INVALID, "n_sets out of range [0, 4096]");
    if (nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (n_sets > 0 && nx > 0 && ny > 0 && (!d_xvec || !d_yvec)) return fail(BQ_ERR_INVALID, "NULL pointer");
    if (!std::isfinite(g)) return fail(BQ_ERR_INVALID, "g is not finite");
    if (n_sets == 0) return BQ_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = order_stream(h, st))) return rc;
    NvtxRange range("bq_state_to_wigner_multi_dev");
    return state_to_wigner_core(h, (const double2 *)d_psi, n_qubits, traced, n_traced, n_sets, 1, 0, d_xvec, nx, d_yvec, ny, g,
                                d_W_out, (double2 *)d_rho_out, st, h->grid_mode == 2);
}

int bq_frame_minmax_dev(bq_handle_t h, const double *d_W, int64_t points_per_frame, int batch, double *d_out,
                        void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_frame_minmax_dev");
    if (points_per_frame < 0 || batch < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (batch == 0) return BQ_OK;
    if (!d_W || !d_out) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = order_stream(h, st);
    if (rc) return rc;
    int nl = 0;
    cudaError_t e = bq::launch_frame_minmax(d_W, points_per_frame, batch, (double2 *)d_out, st, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_frame_minmax");
    return BQ_OK;
}

// matplotlib's LinearSegmentedColormap.from_list("RdBu", 11 ColorBrewer anchors, N=256) lookup table as bytes, same
// operations in the same order as its _create_lookup_table (volatile: no FMA contraction by the host compiler)
static void build_rdbu_lut(unsigned char *lut)
{
    static const int anchors[11][3] = {{103, 0, 31}, {178, 24, 43}, {214, 96, 77}, {244, 165, 130}, {253, 219, 199}, {247, 247, 247},
                                       {209, 229, 240}, {146, 197, 222}, {67, 147, 195}, {33, 102, 172}, {5, 48, 97}};
    double x[11], xind[256];
    for (int i = 0; i < 11; ++i) x[i] = i == 10 ? 1.0 : i * (1.0 / 10.0);
    for (int i = 0; i < 256; ++i) xind[i] = i == 255 ? 1.0 : i * (1.0 / 255.0);
    for (int ch = 0; ch < 3; ++ch) {
        double y[11];
        for (int i = 0; i < 11; ++i) y[i] = anchors[i][ch] / 255.0;
        for (int i = 0; i < 256; ++i) {
            double v;
            if (i == 0) {
                v = y[0];
            } else if (i == 255) {
                v = y[10];
            } else {
                int ind = 0;
                while (ind < 11 && x[ind] < xind[i]) ++ind;  // searchsorted, side="left"
                volatile double num = xind[i] - x[ind - 1], den = x[ind] - x[ind - 1];
                volatile double distance = num / den;
                volatile double prod = distance * (y[ind] - y[ind - 1]);
                v = prod + y[ind - 1];
            }
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            volatile double scaled = v * 255.0;
            lut[4 * i + ch] = (unsigned char)scaled;
        }
    }
    for (int i = 0; i < 256; ++i) lut[4 * i + 3] = 255;
}

int bq_rdbu_lut(unsigned char *lut_out)
{
    if (!lut_out) return fail(BQ_ERR_INVALID, "NULL pointer");
    build_rdbu_lut(lut_out);
    return BQ_OK;
}

int bq_frame_rgba_dev(bq_handle_t h, const double *d_W, int64_t points_per_frame, int batch, int levels,
                      double zero_abs_max, unsigned char *d_rgba_out, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    NvtxRange nvtx_call("bq_frame_rgba_dev");
    if (points_per_frame < 0 || batch < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (levels < 2 || levels > 4096) return fail(BQ_ERR_INVALID, "levels out of range [2, 4096]");
    if (batch == 0 || points_per_frame == 0) return BQ_OK;
    if (!d_W || !d_rgba_out) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = order_stream(h, st);
    if (rc) return rc;
    if (!h->lut_ready) {
        unsigned char lut[256 * 4];
        build_rdbu_lut(lut);
        CU(bq::upload_rdbu_lut(lut));
        h->lut_ready = true;
    }
    CU(h->mm_dev.reserve((size_t)batch * sizeof(double2)));
    int nl = 0;
    cudaError_t e = bq::launch_frame_minmax(d_W, points_per_frame, batch, (double2 *)h->mm_dev.ptr, st, &nl);
    if (e == cudaSuccess)
        e = bq::launch_frame_rgba(d_W, points_per_frame, batch, (const double2 *)h->mm_dev.ptr, levels, zero_abs_max,
                                  d_rgba_out, st, &nl);
    h->launches += nl;
    if (e != cudaSuccess) return fail_cuda(e, "launch_frame_rgba");
    return BQ_OK;
}

// ---- measurement helpers -----------------------------------------------------------------------
int bq_fp64_peak(bq_handle_t h, double seconds, double *tflops)
{
    if (!h || !tflops) return fail(BQ_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    CU(h->small_dev.reserve(512));
    CU(cudaMemsetAsync(h->small_dev.ptr, 0, 512, h->stream));
    const int grid = h->dev.sm_count * 8;
    const int iters = 4096;
    const double flops_per_launch = (double)grid * 256.0 * iters * 16.0 * 8.0 * 2.0;
    // warm-up
    for (int i = 0; i < 3; ++i) CU(bq::launch_fp64_probe((double *)h->small_dev.ptr, iters, grid, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    double best = 0.0, total_ms = 0.0;
    int rounds = 0;
    while (total_ms < seconds * 1e3 && rounds < 100000) {
        CU(cudaEventRecord(h->ev_t0, h->stream));
        for (int i = 0; i < 4; ++i) CU(bq::launch_fp64_probe((double *)h->small_dev.ptr, iters, grid, h->stream));
        CU(cudaEventRecord(h->ev_t1, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
        total_ms += ms;
        ++rounds;
        // sustained figure: the LAST rounds (clocks settled under load), not the best burst
        best = 4.0 * flops_per_launch / (ms * 1e-3) / 1e12;
    }
    *tflops = best;
    return BQ_OK;
}

int bq_last_device_ms(bq_handle_t h, double *ms)
{
    if (!h || !ms) return fail(BQ_ERR_INVALID, "NULL argument");
    *ms = h->last_ms;
    return BQ_OK;
}

int bq_memcpy_dev(bq_handle_t h, void *d_dst, const void *d_src, size_t bytes, void *stream)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (bytes == 0) return BQ_OK;
    if (!d_dst || !d_src) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    int rc = order_stream(h, (cudaStream_t)stream);  // the source is usually this handle's own output of a moment ago
    if (rc) return rc;
    CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return BQ_OK;
}

int bq_set_grid_symmetry(bq_handle_t h, int mode)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (mode < 0 || mode > 2) return fail(BQ_ERR_INVALID, "grid symmetry mode must be 0, 1 or 2");
    std::lock_guard<std::mutex> lk(h->mu);
    h->grid_mode = mode;
    return BQ_OK;
}

int bq_set_unit_range(bq_handle_t h, int64_t first, int64_t count)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (count >= 0 && first < 0) return fail(BQ_ERR_INVALID, "negative first unit");
    std::lock_guard<std::mutex> lk(h->mu);
    h->unit_first = count >= 0 ? first : 0;
    h->unit_count = count >= 0 ? count : -1;
    return BQ_OK;
}

int bq_set_sym_variant(bq_handle_t h, int variant)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->dev.sym_force = variant < 0 ? -1 : variant;
    return BQ_OK;
}

int bq_last_wigner_variant(bq_handle_t h, int *variant)
{
    if (!h || !variant) return fail(BQ_ERR_INVALID, "NULL argument");
    *variant = h->dev.last_variant;
    return BQ_OK;
}

int64_t bq_grid_units(int nx)
{
    const int64_t U = ((int64_t)nx + 1) / 2;
    return nx >= 2 ? U * (U + 1) / 2 : 0;
}

int bq_grid_is_mirror(const double *xvec, int nx, const double *yvec, int ny)
{
    if (nx < 0 || ny < 0 || ((nx > 0 || ny > 0) && (!xvec || !yvec))) return 0;
    return grid_is_mirror(xvec, nx, yvec, ny) ? 1 : 0;
}

int bq_set_sass_mode(bq_handle_t h, int mode, int *active)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (mode != 0 && mode != 1) return fail(BQ_ERR_INVALID, "sass mode must be 0 or 1");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    h->dev.sass_mode = (mode == 1 && bq::patched_sass_available()) ? 1 : 0;
    if (active) *active = h->dev.sass_mode;
    return BQ_OK;
}

int bq_sass_self_check(bq_handle_t h, uint64_t *bad_mask, int *checked)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (bad_mask) *bad_mask = h->dev.patched_bad;
    if (checked) *checked = h->sass_checked;
    return BQ_OK;
}

int bq_profile_enable(bq_handle_t h, int on)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->profile = on != 0;
    return BQ_OK;
}

int bq_profile_read(bq_handle_t h, double *total_ms, int *launches)
{
    if (!h || !total_ms || !launches) return fail(BQ_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->dev.device);
    double tot = 0.0;
    int n = 0;
    for (auto &pr : h->prof_events) {
        CU(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        ++n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    h->prof_events.clear();
    *total_ms = tot;
    *launches = n;
    return BQ_OK;
}

// ---- CUDA IPC ----------------------------------------------------------------------------------
int bq_device_alloc(bq_handle_t h, size_t bytes, void **d_ptr)
{
    if (!h || !d_ptr) return fail(BQ_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->dev.device);
    CU(cudaMalloc(d_ptr, std::max<size_t>(bytes, 1)));
    return BQ_OK;
}

int bq_device_free(bq_handle_t h, void *d_ptr)
{
    if (!h) return fail(BQ_ERR_INVALID, "NULL handle");
    if (!d_ptr) return BQ_OK;
    DeviceGuard guard(h->dev.device);
    CU(cudaFree(d_ptr));
    return BQ_OK;
}

int bq_ipc_export(bq_handle_t h, void *d_ptr, unsigned char *handle64)
{
    if (!h || !d_ptr || !handle64) return fail(BQ_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    DeviceGuard guard(h->dev.device);
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, d_ptr));
    std::memcpy(handle64, &mh, 64);
    return BQ_OK;
}

int bq_ipc_open(bq_handle_t h, const unsigned char *handle64, void **d_ptr)
{
    if (!h || !handle64 || !d_ptr) return fail(BQ_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->dev.device);
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, handle64, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    return BQ_OK;
}

int bq_ipc_close(bq_handle_t h, void *d_ptr)
{
    if (!h || !d_ptr) return fail(BQ_ERR_INVALID, "NULL argument");
    DeviceGuard guard(h->dev.device);
    CU(cudaIpcCloseMemHandle(d_ptr));
    return BQ_OK;
}

}  // extern "C"
