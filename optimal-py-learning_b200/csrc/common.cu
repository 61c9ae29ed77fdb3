// Error plumbing, device queries and the n-vector helpers of the optimizer stack.
#include <stdarg.h>

#include "common.cuh"

namespace opl {

std::string &last_error() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev & 63;
}

int sm_count() {
    static int cached[64] = {};
    int &slot = cached[current_device_slot()];
    if (slot == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            slot = n;
        else
            slot = 148;
    }
    return slot;
}

// ---------------------------------------------------------------------------------------
// vector kernels.  Products and sums are rounded separately (__dmul_rn/__dadd_rn block FMA
// contraction) so `x + a*d` is bit-identical to NumPy's two-step evaluation
// (optimizer.py:97,136-141,253,431; linesearch.py:394).
// ---------------------------------------------------------------------------------------
__global__ void axpy_kernel(int64_t n, const double *__restrict__ x, double a, const double *__restrict__ d,
                            double b, const double *__restrict__ p, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = __dadd_rn(x[i], __dmul_rn(a, d[i]));
        if (p) v = __dadd_rn(v, __dmul_rn(b, p[i]));
        out[i] = v;
    }
}

// step = fl(a * d); out = x + step: `prev_step = step_size * step_dir; parameters + prev_step` (optimizer.py:251-253,
// 429-431) in one launch, each product and sum rounded once as NumPy does
__global__ void step_kernel(int64_t n, const double *__restrict__ x, double a, const double *__restrict__ d,
                            double *__restrict__ step, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = __dmul_rn(a, d[i]);
        step[i] = t;
        out[i] = __dadd_rn(x[i], t);
    }
}

__global__ void scale_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __dmul_rn(a, x[i]);
}

// stage 1 of a deterministic dot: one partial per block
__global__ void dot_partial_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                   double *__restrict__ partial) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += x[i] * y[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void dot_final_kernel(int nparts, const double *__restrict__ partial, double *__restrict__ out) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[0] = s;
}

// short vectors: the whole dot product in ONE block; the scalar goes to mapped pinned memory followed by a sequence number,
// so the host needs neither a second launch nor a memcpy (one line-search step of a small model calls this once)
__global__ void __launch_bounds__(1024) dot_single_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                         volatile double *host, double seq) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * y[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        host[0] = s;
        __threadfence_system();
        host[1] = seq;
    }
}

// stage 1 of the penalty norms: sum|w| (kind 1) or sum w^2 (kind 2)
__global__ void penalty_partial_kernel(int64_t n, const double *__restrict__ w, int kind, double *__restrict__ partial) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += kind == 1 ? fabs(w[i]) : w[i] * w[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// every block folds the partials identically, then adds its slice of the penalty gradient
__global__ void penalty_apply_kernel(int64_t n, const double *__restrict__ w, int kind, double lambda, int nparts,
                                     const double *__restrict__ partial, double *grad, double *value_out) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
    s = block_sum(s, scratch);
    const double norm = kind == 1 ? s : sqrt(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) value_out[0] = lambda * norm;
    if (!grad) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double wi = w[i];
        const double d = kind == 1 ? (wi > 0.0 ? 1.0 : (wi < 0.0 ? -1.0 : 0.0)) : wi / norm;   // numpy.sign / w / ||w||
        grad[i] += lambda * d;
    }
}

// dst[i][:] = src[index[i]][:]: one warp per 32-column segment of a row, rows strided over the grid
__global__ void gather_rows_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, const int64_t *__restrict__ index,
                                   int64_t count, double *__restrict__ dst) {
    const int64_t total = count * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const int64_t from = index[r];
        dst[i] = (from >= 0 && from < rows) ? src[from * cols + c] : __longlong_as_double(0x7ff8000000000000LL);
    }
}

static int vec_grid(int64_t n, int threads) {
    int64_t g = ceil_div(n, threads);
    int64_t cap = 4 * (int64_t)sm_count();
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace opl

using namespace opl;

extern "C" {

const char *opl_last_error(void) { return last_error().c_str(); }

int opl_version(void) { return 100; }

int opl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int opl_vec_axpy_device(int64_t n, const double *x, double a, const double *d, double b, const double *p,
                        double *out, void *cuda_stream) {
    if (n < 0 || (n > 0 && (!x || !d || !out))) return fail(OPL_E_ARG, "opl_vec_axpy_device: null pointer");
    if (n == 0) return OPL_OK;
    axpy_kernel<<<vec_grid(n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(n, x, a, d, b, p, out);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

int opl_gather_rows_device(const double *src_dev, int64_t rows, int64_t cols, const int64_t *index_dev, int64_t count,
                           double *dst_dev, void *cuda_stream) {
    if (rows < 1 || cols < 1 || count < 0 || !src_dev || (count > 0 && (!index_dev || !dst_dev)))
        return fail(OPL_E_ARG, "opl_gather_rows_device: bad arguments");
    if (count == 0) return OPL_OK;
    gather_rows_kernel<<<vec_grid(count * cols, 256), 256, 0, (cudaStream_t)cuda_stream>>>(src_dev, rows, cols, index_dev, count, dst_dev);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

int opl_vec_step_device(int64_t n, const double *x, double a, const double *d, double *step_out, double *out, void *cuda_stream) {
    if (n < 0 || (n > 0 && (!x || !d || !step_out || !out))) return fail(OPL_E_ARG, "opl_vec_step_device: null pointer");
    if (n == 0) return OPL_OK;
    step_kernel<<<vec_grid(n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(n, x, a, d, step_out, out);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

int opl_vec_scale_device(int64_t n, double a, const double *x, double *out, void *cuda_stream) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(OPL_E_ARG, "opl_vec_scale_device: null pointer");
    if (n == 0) return OPL_OK;
    scale_kernel<<<vec_grid(n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(n, a, x, out);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

}  // extern "C"

namespace {
// Per-device reduction scratch and page-locked scalar for the synchronous vector calls: the calls block until their
// scalar is on the host, so one buffer per device is enough.  (A cudaMallocAsync / cudaFreeAsync pair plus a download into
// a pageable stack variable made one dot product cost 0.4 ms -- as much as half an objective evaluation.)
struct VecScratch {
    double *dev = nullptr;
    double *host = nullptr;
    double *mapped_host = nullptr, *mapped_dev = nullptr;   // value, sequence number (dot_single_kernel)
    double seq = 0.0;
    int blocks = 0;
};
int vec_scratch(int blocks, VecScratch **out) {
    static VecScratch table[64];
    int device = 0;
    OPL_CUDA(cudaGetDevice(&device));
    if (device < 0 || device >= 64) return fail(OPL_E_CUDA, "device index %d out of range", device);
    VecScratch &v = table[device];
    if (v.blocks < blocks) {
        cudaFree(v.dev);
        v.dev = nullptr;
        v.blocks = 0;
        OPL_CUDA(cudaMalloc(&v.dev, sizeof(double) * (size_t)(blocks + 1)));
        v.blocks = blocks;
    }
    if (!v.host) OPL_CUDA(cudaMallocHost(&v.host, sizeof(double) * 2));
    if (!v.mapped_host) {
        OPL_CUDA(cudaHostAlloc(&v.mapped_host, sizeof(double) * 2, cudaHostAllocMapped));
        v.mapped_host[0] = v.mapped_host[1] = 0.0;
        OPL_CUDA(cudaHostGetDevicePointer(&v.mapped_dev, v.mapped_host, 0));
    }
    *out = &v;
    return OPL_OK;
}
}  // namespace

extern "C" {

int opl_vec_penalty_device(int64_t n, const double *w, int32_t kind, double lambda, double *grad_inout,
                           double *host_out, void *cuda_stream) {
    if (n < 1 || !w || !host_out || (kind != 1 && kind != 2)) return fail(OPL_E_ARG, "opl_vec_penalty_device: bad arguments");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int blocks = vec_grid(n, 256);
    VecScratch *v = nullptr;
    int rc = vec_scratch(blocks, &v);
    if (rc != OPL_OK) return rc;
    penalty_partial_kernel<<<blocks, 256, 0, st>>>(n, w, kind, v->dev);
    penalty_apply_kernel<<<blocks, 256, 0, st>>>(n, w, kind, lambda, blocks, v->dev, grad_inout, v->dev + blocks);
    OPL_CUDA(cudaMemcpyAsync(v->host, v->dev + blocks, sizeof(double), cudaMemcpyDeviceToHost, st));
    OPL_CUDA(cudaStreamSynchronize(st));
    OPL_LAUNCH_CHECK();
    *host_out = v->host[0];
    return OPL_OK;
}

int opl_vec_dot_device(int64_t n, const double *x, const double *y, double *host_out, void *cuda_stream) {
    if (n < 0 || !host_out || (n > 0 && (!x || !y))) return fail(OPL_E_ARG, "opl_vec_dot_device: null pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int blocks = n > 0 ? vec_grid(n, 256) : 1;
    VecScratch *v = nullptr;
    int rc = vec_scratch(blocks, &v);
    if (rc != OPL_OK) return rc;
    if (n <= 65536) {
        v->seq += 1.0;
        const double seq = v->seq;
        dot_single_kernel<<<1, n <= 4096 ? 256 : 1024, 0, st>>>(n, x, y, v->mapped_dev, seq);
        OPL_LAUNCH_CHECK();
        volatile double *m = v->mapped_host;
        bool seen = false;
        for (int spin = 0; spin < 2000000 && !seen; ++spin) seen = m[1] == seq;
        if (!seen) {
            OPL_CUDA(cudaStreamSynchronize(st));
            if (m[1] != seq) return fail(OPL_E_CUDA, "opl_vec_dot_device: the kernel did not publish its result");
        }
        *host_out = m[0];
        return OPL_OK;
    }
    dot_partial_kernel<<<blocks, 256, 0, st>>>(n, x, y, v->dev);
    dot_final_kernel<<<1, 256, 0, st>>>(blocks, v->dev, v->dev + blocks);
    OPL_CUDA(cudaMemcpyAsync(v->host, v->dev + blocks, sizeof(double), cudaMemcpyDeviceToHost, st));
    OPL_CUDA(cudaStreamSynchronize(st));
    OPL_LAUNCH_CHECK();
    *host_out = v->host[0];
    return OPL_OK;
}

}  // extern "C"
