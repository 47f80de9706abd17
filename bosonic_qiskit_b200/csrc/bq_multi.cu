// Single-process multi-GPU entry points of the C ABI (include/bq_b200.h, "several GPUs from one process").
//
// The reference fans its frame loop out from inside a plain function call -- multiprocessing.Pool.starmap over
// frames, src/bosonic_qiskit/animate.py:166-169,191-206; sequentially at wigner.py:92-96 and animate.py:326-352.
// bq_multi_* does the same from one host process over the GPUs of one box: one bq_handle_t and one worker thread
// per device, no collective on the data path.
//
//   frames    (BQ_SHARD_FRAMES): contiguous frame blocks per GPU; every GPU copies its own block of statevectors in
//             and its finished grids out, straight into its slice of the caller's host array over its own PCIe
//             link (nothing is funnelled through GPU 0: 512 frames of 400 x 400 are 655 MB).
//   one grid  (BQ_SHARD_GRID): a single large grid.  A mirror-symmetric grid is shared out by equal ranges of its
//             units (pairs of grid lines, 8 points each, all of equal cost); every GPU stores its points directly
//             into ONE frame buffer on the first GPU through NVLink peer memory (the shares touch disjoint points:
//             no zeroing, no reduction), which then goes to the host as one copy.  Any other grid is cut into row
//             slabs that each GPU writes to its rows of the host array.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/bq_b200.h"
#include "bq_internal.h"

extern "C" __attribute__((visibility("hidden"))) int bq_internal_fail(int code, const char *msg);

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = 0;
    std::string err;
};

void worker_main(Worker *w)
{
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        int rc = job();
        std::string err = rc ? bq_last_error() : "";  // bq_last_error is per thread: hand it to the caller
        lk.lock();
        w->rc = rc;
        w->err = std::move(err);
        w->done = true;
        w->cv.notify_all();
    }
}

}  // namespace

struct bq_multi_s {
    std::vector<int> devices;
    std::vector<bq_handle_t> handles;
    std::vector<Worker *> workers;
    std::vector<cudaStream_t> streams;   // per device, for the `_dev` calls of the one-grid mode
    std::vector<void *> d_in, d_grid;    // per device: staged rho / psi and the grid axes (grow-only)
    std::vector<size_t> d_in_bytes, d_grid_bytes;
    void *frame0 = nullptr;              // one-grid mode: the frame buffer on devices[0] every GPU stores into
    size_t frame0_bytes = 0;
    bool peer_ok = true;                 // every GPU can address devices[0]
    std::mutex mu;                       // one sharded call at a time
};

namespace {

int fail(int code, const std::string &msg) { return bq_internal_fail(code, msg.c_str()); }

int fail_cuda(cudaError_t e, const char *what)
{
    return fail(BQ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define CUM(call)                                             \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

void submit(Worker *w, std::function<int()> job)
{
    std::lock_guard<std::mutex> lk(w->mu);
    w->job = std::move(job);
    w->has_job = true;
    w->done = false;
    w->cv.notify_all();
}

// waits for every worker; returns the first failure (with its message as the calling thread's bq_last_error)
int wait_all(bq_multi_s *m)
{
    int rc = BQ_OK;
    std::string err;
    for (size_t k = 0; k < m->workers.size(); ++k) {
        Worker *w = m->workers[k];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc && !rc) {
            rc = w->rc;
            err = "device " + std::to_string(m->devices[k]) + ": " + w->err;
        }
        w->rc = 0;
        w->err.clear();
    }
    return rc ? fail(rc, err) : BQ_OK;
}

// contiguous block [lo, hi) of n units owned by part k of parts; the first n % parts get one more
void shard_range(int64_t n, int k, int parts, int64_t *lo, int64_t *hi)
{
    const int64_t base = n / parts, extra = n % parts;
    *lo = k * base + std::min<int64_t>(k, extra);
    *hi = *lo + base + (k < extra ? 1 : 0);
}

int reserve(bq_multi_s *m, std::vector<void *> &ptrs, std::vector<size_t> &sizes, size_t k, size_t bytes)
{
    if (bytes <= sizes[k]) return BQ_OK;
    if (ptrs[k]) {
        int rc = bq_device_free(m->handles[k], ptrs[k]);
        ptrs[k] = nullptr;
        sizes[k] = 0;
        if (rc) return rc;
    }
    int rc = bq_device_alloc(m->handles[k], bytes, &ptrs[k]);
    if (rc) return rc;
    sizes[k] = bytes;
    return BQ_OK;
}

struct SetDevice {
    int prev = -1;
    explicit SetDevice(int dev)
    {
        cudaGetDevice(&prev);
        cudaSetDevice(dev);
    }
    ~SetDevice()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// One large grid on all GPUs.  `rho` (host, N x N with leading dimension ld) or `psi` (host, one statevector) is
// replicated; W_out is the host grid [ny][nx].
int one_grid(bq_multi_s *m, const double *rho, int N, int64_t ld, const double *psi, int n_qubits, const int *traced,
             int n_traced, const double *xvec, int nx, const double *yvec, int ny, double g, int method, double *W_out,
             double *rho_out)
{
    const size_t n_dev = m->devices.size();
    const bool mirror = bq_grid_is_mirror(xvec, nx, yvec, ny) != 0;
    const int64_t units = bq_grid_units(nx);
    int sm_count = 0;
    CUM(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, m->devices[0]));
    // unit ranges only where ONE GPU would run the recurrence-sharing kernels on this grid too: same kernels, same bits
    const bool by_units = mirror && m->peer_ok && N >= 2 && bq::sym_kernels_cover(N, units, sm_count);
    const size_t in_bytes = psi ? sizeof(double) * 2 * ((size_t)1 << n_qubits) : sizeof(double) * 2 * ((size_t)(N - 1) * ld + N);
    const size_t frame_bytes = sizeof(double) * (size_t)nx * ny;
    std::vector<int> traced_v(traced, traced + (psi ? n_traced : 0));
    if (by_units) {
        if (frame_bytes > m->frame0_bytes) {
            if (m->frame0) bq_device_free(m->handles[0], m->frame0);
            m->frame0 = nullptr;
            m->frame0_bytes = 0;
            int rc = bq_device_alloc(m->handles[0], frame_bytes, &m->frame0);
            if (rc) return rc;
            m->frame0_bytes = frame_bytes;
        }
    }
    for (size_t k = 0; k < n_dev; ++k) {
        submit(m->workers[k], [=]() -> int {
            bq_handle_t h = m->handles[k];
            SetDevice sd(m->devices[k]);
            cudaStream_t st = m->streams[k];
            int rc = reserve(m, m->d_in, m->d_in_bytes, k, in_bytes);
            if (rc) return rc;
            if ((rc = reserve(m, m->d_grid, m->d_grid_bytes, k, sizeof(double) * ((size_t)nx + ny)))) return rc;
            double *d_x = (double *)m->d_grid[k], *d_y = d_x + nx;
            CUM(cudaMemcpyAsync(m->d_in[k], psi ? psi : rho, in_bytes, cudaMemcpyHostToDevice, st));
            CUM(cudaMemcpyAsync(d_x, xvec, sizeof(double) * nx, cudaMemcpyHostToDevice, st));
            CUM(cudaMemcpyAsync(d_y, yvec, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
            double *d_rho_out = nullptr;
            void *rho_tmp = nullptr;
            if (psi && rho_out && k == 0) {
                const size_t D = (size_t)1 << (n_qubits - n_traced);
                if ((rc = bq_device_alloc(h, sizeof(double) * 2 * D * D, &rho_tmp))) return rc;
                d_rho_out = (double *)rho_tmp;
            }
            int64_t lo, hi;
            double *d_W;
            void *slab = nullptr;
            if (by_units) {
                shard_range(units, (int)k, (int)n_dev, &lo, &hi);
                if ((rc = bq_set_grid_symmetry(h, 2))) return rc;  // this call vouches for the grid it tested above
                if ((rc = bq_set_unit_range(h, lo, hi - lo))) return rc;
                d_W = (double *)m->frame0;  // peer memory for k > 0
                if (psi)
                    rc = bq_state_to_wigner_dev(h, (const double *)m->d_in[k], n_qubits, traced_v.data(), n_traced, 1, 0, d_x, nx,
                                                d_y, ny, g, method, d_W, d_rho_out, st);
                else
                    rc = bq_wigner_dev(h, (const double *)m->d_in[k], N, ld, 0, 1, d_x, nx, d_y, ny, g, method, d_W, st);
                bq_set_unit_range(h, 0, -1);
                bq_set_grid_symmetry(h, 1);
            } else {
                // row slab [lo, hi) of the grid with the point-by-point kernels, into this GPU's own buffer
                shard_range(ny, (int)k, (int)n_dev, &lo, &hi);
                if ((rc = bq_set_grid_symmetry(h, 0))) return rc;
                rc = BQ_OK;
                if (hi > lo) {
                    if ((rc = bq_device_alloc(h, sizeof(double) * (size_t)(hi - lo) * nx, &slab))) return rc;
                    d_W = (double *)slab;
                    if (psi)
                        rc = bq_state_to_wigner_dev(h, (const double *)m->d_in[k], n_qubits, traced_v.data(), n_traced, 1, 0, d_x,
                                                    nx, d_y + lo, (int)(hi - lo), g, method, d_W, d_rho_out, st);
                    else
                        rc = bq_wigner_dev(h, (const double *)m->d_in[k], N, ld, 0, 1, d_x, nx, d_y + lo, (int)(hi - lo), g, method,
                                           d_W, st);
                    if (!rc) {
                        cudaError_t e = cudaMemcpyAsync(W_out + lo * nx, d_W, sizeof(double) * (size_t)(hi - lo) * nx,
                                                        cudaMemcpyDeviceToHost, st);
                        if (e != cudaSuccess) rc = fail_cuda(e, "slab copy to host");
                    }
                }
                bq_set_grid_symmetry(h, 1);
            }
            if (!rc && d_rho_out) {
                const size_t D = (size_t)1 << (n_qubits - n_traced);
                cudaError_t e = cudaMemcpyAsync(rho_out, d_rho_out, sizeof(double) * 2 * D * D, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) rc = fail_cuda(e, "rho copy to host");
            }
            cudaError_t e = cudaStreamSynchronize(st);
            if (slab) bq_device_free(h, slab);
            if (rho_tmp) bq_device_free(h, rho_tmp);
            if (rc) return rc;
            if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize");
            return BQ_OK;
        });
    }
    int rc = wait_all(m);
    if (rc) return rc;
    if (by_units) {
        SetDevice sd(m->devices[0]);
        CUM(cudaMemcpy(W_out, m->frame0, frame_bytes, cudaMemcpyDeviceToHost));
    }
    return BQ_OK;
}

int check_multi(bq_multi_t m, int shard_mode, int batch)
{
    if (!m) return fail(BQ_ERR_INVALID, "NULL multi-GPU handle");
    if (shard_mode != BQ_SHARD_FRAMES && shard_mode != BQ_SHARD_GRID)
        return fail(BQ_ERR_INVALID, "shard_mode must be BQ_SHARD_FRAMES or BQ_SHARD_GRID");
    if (batch < 0) return fail(BQ_ERR_INVALID, "negative batch");
    if (shard_mode == BQ_SHARD_GRID && batch > 1) return fail(BQ_ERR_INVALID, "BQ_SHARD_GRID shares ONE grid out: batch must be 1");
    return BQ_OK;
}

}  // namespace

extern "C" {

int bq_multi_init(bq_multi_t *out, int n_dev, const int *devs)
{
    if (!out) return fail(BQ_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    int rc = bq_device_count(&count);
    if (rc) return rc;
    if (count == 0) return fail(BQ_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    std::vector<int> devices;
    if (n_dev <= 0 || !devs) {
        for (int d = 0; d < (n_dev > 0 ? std::min(n_dev, count) : count); ++d) devices.push_back(d);
    } else {
        for (int k = 0; k < n_dev; ++k) {
            if (devs[k] < 0 || devs[k] >= count) return fail(BQ_ERR_NO_DEVICE, "device ordinal out of range");
            devices.push_back(devs[k]);  // (a device may be listed more than once: it then takes that many shares)
        }
    }
    bq_multi_s *m = new bq_multi_s();
    m->devices = devices;
    const size_t n = devices.size();
    m->handles.assign(n, nullptr);
    m->streams.assign(n, nullptr);
    m->d_in.assign(n, nullptr);
    m->d_grid.assign(n, nullptr);
    m->d_in_bytes.assign(n, 0);
    m->d_grid_bytes.assign(n, 0);
    for (size_t k = 0; k < n; ++k) {
        if ((rc = bq_create(&m->handles[k], devices[k]))) {
            bq_multi_finalize(m);
            return rc;
        }
        SetDevice sd(devices[k]);
        if (cudaStreamCreateWithFlags(&m->streams[k], cudaStreamNonBlocking) != cudaSuccess) {
            cudaError_t e = cudaGetLastError();
            bq_multi_finalize(m);
            return fail_cuda(e, "cudaStreamCreate");
        }
        if (k > 0) {  // the one-grid mode stores into devices[0]'s memory
            int can = 0;
            if (devices[k] == devices[0]) {
                can = 2;  // the same GPU: its own memory
            } else {
                cudaDeviceCanAccessPeer(&can, devices[k], devices[0]);
            }
            if (can == 1) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) {
                    cudaGetLastError();
                    can = 0;
                }
            }
            if (!can) m->peer_ok = false;
        }
    }
    for (size_t k = 0; k < n; ++k) {
        Worker *w = new Worker();
        m->workers.push_back(w);
        w->th = std::thread(worker_main, w);
    }
    *out = m;
    return BQ_OK;
}

int bq_multi_finalize(bq_multi_t m)
{
    if (!m) return BQ_OK;
    for (Worker *w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t k = 0; k < m->handles.size(); ++k) {
        if (!m->handles[k]) continue;
        if (m->d_in[k]) bq_device_free(m->handles[k], m->d_in[k]);
        if (m->d_grid[k]) bq_device_free(m->handles[k], m->d_grid[k]);
        if (k == 0 && m->frame0) bq_device_free(m->handles[0], m->frame0);
        if (m->streams[k]) {
            SetDevice sd(m->devices[k]);
            cudaStreamDestroy(m->streams[k]);
        }
        bq_destroy(m->handles[k]);
    }
    delete m;
    return BQ_OK;
}

int bq_multi_device_count(bq_multi_t m, int *count)
{
    if (!m || !count) return fail(BQ_ERR_INVALID, "NULL argument");
    *count = (int)m->devices.size();
    return BQ_OK;
}

int bq_multi_handle(bq_multi_t m, int index, bq_handle_t *out)
{
    if (!m || !out) return fail(BQ_ERR_INVALID, "NULL argument");
    if (index < 0 || index >= (int)m->handles.size()) return fail(BQ_ERR_INVALID, "device index out of range");
    *out = m->handles[index];
    return BQ_OK;
}

int bq_multi_launch_count(bq_multi_t m, uint64_t *count)
{
    if (!m || !count) return fail(BQ_ERR_INVALID, "NULL argument");
    uint64_t tot = 0;
    for (bq_handle_t h : m->handles) {
        uint64_t c = 0;
        int rc = bq_launch_count(h, &c);
        if (rc) return rc;
        tot += c;
    }
    *count = tot;
    return BQ_OK;
}

int bq_state_to_wigner_sharded_host(bq_multi_t m, const double *psi, int n_qubits, const int *traced, int n_traced,
                                    int batch, int64_t psi_stride, const double *xvec, int nx, const double *yvec,
                                    int ny, double g, int method, int shard_mode, double *W_out, double *rho_out)
{
    int rc = check_multi(m, shard_mode, batch);
    if (rc) return rc;
    if (batch == 0) return BQ_OK;
    if (n_qubits < 0 || n_qubits > 40 || n_traced < 0 || n_traced > n_qubits) return fail(BQ_ERR_INVALID, "bad qubit counts");
    if (!psi || !W_out || (n_traced > 0 && !traced)) return fail(BQ_ERR_INVALID, "NULL pointer");
    if (nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "negative size");
    if (nx > 0 && ny > 0 && (!xvec || !yvec)) return fail(BQ_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> lk(m->mu);
    nvtxRangePushA("bq_state_to_wigner_sharded_host");
    if (shard_mode == BQ_SHARD_GRID) {
        if (nx == 0 || ny == 0) {
            nvtxRangePop();
            return BQ_OK;
        }
        rc = one_grid(m, nullptr, 1 << (n_qubits - n_traced), 0, psi, n_qubits, traced, n_traced, xvec, nx, yvec, ny, g, method,
                      W_out, rho_out);
        nvtxRangePop();
        return rc;
    }
    const size_t n_dev = m->devices.size();
    const int64_t frame_pts = (int64_t)nx * ny, D = (int64_t)1 << (n_qubits - n_traced);
    std::vector<int> traced_v(traced, traced + n_traced);
    for (size_t k = 0; k < n_dev; ++k) {
        int64_t lo, hi;
        shard_range(batch, (int)k, (int)n_dev, &lo, &hi);
        if (hi <= lo) continue;
        submit(m->workers[k], [=]() -> int {
            return bq_state_to_wigner_host(m->handles[k], psi + 2 * lo * psi_stride, n_qubits, traced_v.data(), n_traced,
                                           (int)(hi - lo), psi_stride, xvec, nx, yvec, ny, g, method, W_out + lo * frame_pts,
                                           rho_out ? rho_out + 2 * lo * D * D : nullptr);
        });
    }
    rc = wait_all(m);
    nvtxRangePop();
    return rc;
}

int bq_wigner_sharded_host(bq_multi_t m, const double *rho, int N, int64_t ld, int64_t rho_stride, int batch,
                           const double *xvec, int nx, const double *yvec, int ny, double g, int method, int shard_mode,
                           double *W_out)
{
    int rc = check_multi(m, shard_mode, batch);
    if (rc) return rc;
    if (batch == 0 || nx == 0 || ny == 0) return BQ_OK;
    if (N < 1 || N > 16384 || ld < N) return fail(BQ_ERR_INVALID, "bad cutoff / leading dimension");
    if (!rho || !W_out || !xvec || !yvec || nx < 0 || ny < 0) return fail(BQ_ERR_INVALID, "NULL pointer / negative size");
    std::lock_guard<std::mutex> lk(m->mu);
    nvtxRangePushA("bq_wigner_sharded_host");
    if (shard_mode == BQ_SHARD_GRID) {
        rc = one_grid(m, rho, N, ld, nullptr, 0, nullptr, 0, xvec, nx, yvec, ny, g, method, W_out, nullptr);
        nvtxRangePop();
        return rc;
    }
    const size_t n_dev = m->devices.size();
    const int64_t frame_pts = (int64_t)nx * ny;
    for (size_t k = 0; k < n_dev; ++k) {
        int64_t lo, hi;
        shard_range(batch, (int)k, (int)n_dev, &lo, &hi);
        if (hi <= lo) continue;
        submit(m->workers[k], [=]() -> int {
            return bq_wigner_host(m->handles[k], rho + 2 * lo * rho_stride, N, ld, rho_stride, (int)(hi - lo), xvec, nx, yvec, ny, g,
                                  method, W_out + lo * frame_pts);
        });
    }
    rc = wait_all(m);
    nvtxRangePop();
    return rc;
}

}  // extern "C"
