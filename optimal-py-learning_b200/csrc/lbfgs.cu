// L-BFGS two-loop recursion as ONE persistent cooperative kernel.
//
// Replaces LBFGS._update_diffs / _lbfgs_step_dir / _initial_inv_hessian_scalar
// (optimize/optimizer.py:433-508): 2m dot products that each gate an axpy.  The reference
// makes 4m+2 passes over n-vectors with a Python round trip between them; here the grid stays
// resident (cooperative launch, grid.sync between dependent phases), every CTA folds the same
// per-CTA partial sums in the same order (bit-reproducible, identical in all CTAs) and each
// thread owns a fixed slice of q, so no other synchronisation is needed.
//
// Reference quirks kept: `rho` stores y.s and is divided by (:464-468); pairs with y.s == 0 are
// skipped in both loops (:465-470,486); gamma comes from the OLDEST stored pair (:507-508);
// products and sums are rounded separately (no FMA contraction) to mirror NumPy's temporaries.
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace opl {

constexpr int LB_THREADS = 256;
constexpr int LB_MAX_SMEM_PAIRS = 64;   // ys/alpha kept in shared memory up to this history length

struct LbfgsArgs {
    int64_t n;
    int m;                       // pairs in the history, oldest first
    const double *const *S;      // device table of m pointers
    const double *const *Y;
    const double *g;
    double *q;                   // work vector (n)
    double *dir;                 // out (n)
    double *partial;             // [2][grid][4]
    double *pair_scratch;        // [2][m] global ys/alpha when m > LB_MAX_SMEM_PAIRS
    int gamma_mode;
};

// fold per-CTA partials (k values each) in a fixed order; result in every thread
template <int K>
__device__ __forceinline__ void fold_partials(const double *partial, int nblocks, double (&out)[K], double *scratch) {
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] += partial[4 * i + k];
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = block_sum(a[k], scratch);
}

__global__ void __launch_bounds__(LB_THREADS) lbfgs_two_loop_kernel(const LbfgsArgs p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double scratch[33];
    __shared__ double ys_s[LB_MAX_SMEM_PAIRS], alpha_s[LB_MAX_SMEM_PAIRS];
    double *ys_tab = p.m <= LB_MAX_SMEM_PAIRS ? ys_s : p.pair_scratch;
    double *alpha_tab = p.m <= LB_MAX_SMEM_PAIRS ? alpha_s : p.pair_scratch + p.m;
    const bool table_shared = p.m <= LB_MAX_SMEM_PAIRS;

    const int64_t first = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int nb = gridDim.x;
    int phase = 0;

    for (int64_t i = first; i < p.n; i += stride) p.q[i] = p.g[i];   // newton_grad = copy(jacobian)

    double gamma = 1.0;
    // ---- first loop: newest -> oldest (optimizer.py:460-474) ----
    for (int j = p.m - 1; j >= 0; --j) {
        const double *s = p.S[j], *y = p.Y[j];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int64_t i = first; i < p.n; i += stride) {
            const double yi = y[i], si = s[i];
            a0 += yi * si;        // rho (stores y.s)
            a1 += si * p.q[i];    // s.q
            a2 += yi * yi;        // only used for the oldest pair (gamma)
        }
        double *part = p.partial + (size_t)(phase & 1) * nb * 4;
        a0 = block_sum(a0, scratch);
        a1 = block_sum(a1, scratch);
        a2 = block_sum(a2, scratch);
        if (threadIdx.x == 0) {
            part[4 * blockIdx.x + 0] = a0;
            part[4 * blockIdx.x + 1] = a1;
            part[4 * blockIdx.x + 2] = a2;
        }
        grid.sync();
        double t[3];
        fold_partials<3>(part, nb, t, scratch);
        ++phase;
        const double ys = t[0];
        double alpha = 0.0;
        if (ys != 0.0) {
            alpha = t[1] / ys;
            for (int64_t i = first; i < p.n; i += stride) p.q[i] = __dadd_rn(p.q[i], -__dmul_rn(alpha, y[i]));
        }
        if (table_shared) {
            if (threadIdx.x == 0) {
                ys_tab[j] = ys;
                alpha_tab[j] = alpha;
            }
        } else if (blockIdx.x == 0 && threadIdx.x == 0) {
            ys_tab[j] = ys;
            alpha_tab[j] = alpha;
        }
        if (j == 0 && p.gamma_mode == OPL_LBFGS_GAMMA_OLDEST)   // optimizer.py:346-362 on pair [0]
            gamma = t[2] == 0.0 ? 1.0 : nan_to_num(ys / t[2]);
    }
    if (!table_shared) grid.sync();
    else __syncthreads();

    // ---- scale by the initial inverse-Hessian scalar (optimizer.py:478) ----
    for (int64_t i = first; i < p.n; i += stride) p.q[i] = __dmul_rn(p.q[i], gamma);

    // ---- second loop: oldest -> newest (optimizer.py:479-488) ----
    for (int j = 0; j < p.m; ++j) {
        const double ys = ys_tab[j];
        if (ys == 0.0) continue;   // uniform across the grid
        const double *s = p.S[j], *y = p.Y[j];
        double a0 = 0.0;
        for (int64_t i = first; i < p.n; i += stride) a0 += y[i] * p.q[i];
        double *part = p.partial + (size_t)(phase & 1) * nb * 4;
        a0 = block_sum(a0, scratch);
        if (threadIdx.x == 0) part[4 * blockIdx.x] = a0;
        grid.sync();
        double t[1];
        fold_partials<1>(part, nb, t, scratch);
        ++phase;
        const double coeff = alpha_tab[j] - t[0] / ys;
        for (int64_t i = first; i < p.n; i += stride) p.q[i] = __dadd_rn(p.q[i], __dmul_rn(s[i], coeff));
    }

    for (int64_t i = first; i < p.n; i += stride) p.dir[i] = -p.q[i];
}

__global__ void lbfgs_sub_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                                 double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] - b[i];
}

}  // namespace opl

using namespace opl;

struct opl_lbfgs {
    int64_t n = 0;
    int memory = 0;             // <= 0: unbounded
    cudaStream_t stream = nullptr;
    std::vector<double *> s_hist, y_hist;   // oldest first
    std::vector<double *> spare;            // recycled buffers
    double *q = nullptr;
    double *partial = nullptr;
    double *pair_scratch = nullptr;
    int pair_capacity = 0;
    const double **table = nullptr;         // device: [2][table_capacity]
    int table_capacity = 0;
    int grid = 0;
    int64_t launches = 0;
};

namespace {

int grab(opl_lbfgs *l, double **out) {
    if (!l->spare.empty()) {
        *out = l->spare.back();
        l->spare.pop_back();
        return OPL_OK;
    }
    OPL_CUDA(cudaMalloc(out, sizeof(double) * (size_t)l->n));
    return OPL_OK;
}

}  // namespace

extern "C" {

int opl_lbfgs_create(opl_lbfgs **out, int64_t n, int32_t memory, void *cuda_stream) {
    if (!out || n < 1) return fail(OPL_E_ARG, "opl_lbfgs_create: bad arguments");
    if (opl_device_count() < 1) return fail(OPL_E_CUDA, "no CUDA device: opl_b200 has no CPU fallback");
    opl_lbfgs *l = new opl_lbfgs();
    l->n = n;
    l->memory = memory;
    l->stream = (cudaStream_t)cuda_stream;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_two_loop_kernel, LB_THREADS, 0);
    if (e != cudaSuccess || per_sm < 1) {
        delete l;
        return fail(OPL_E_CUDA, "opl_lbfgs_create: occupancy query failed: %s", cudaGetErrorString(e));
    }
    const int64_t resident = (int64_t)sm_count() * std::min(per_sm, 2);
    l->grid = (int)std::max<int64_t>(1, std::min<int64_t>(resident, ceil_div(n, LB_THREADS)));
    e = cudaMalloc(&l->q, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&l->partial, sizeof(double) * 2 * 4 * (size_t)l->grid);
    if (e != cudaSuccess) {
        opl_lbfgs_destroy(l);
        return fail(OPL_E_CUDA, "opl_lbfgs_create: %s", cudaGetErrorString(e));
    }
    *out = l;
    return OPL_OK;
}

int opl_lbfgs_destroy(opl_lbfgs *l) {
    if (!l) return OPL_OK;
    cudaStreamSynchronize(l->stream);
    for (double *p : l->s_hist) cudaFree(p);
    for (double *p : l->y_hist) cudaFree(p);
    for (double *p : l->spare) cudaFree(p);
    cudaFree(l->q);
    cudaFree(l->partial);
    cudaFree(l->pair_scratch);
    cudaFree((void *)l->table);
    delete l;
    return OPL_OK;
}

int opl_lbfgs_reset(opl_lbfgs *l) {
    if (!l) return fail(OPL_E_ARG, "opl_lbfgs_reset: null");
    for (double *p : l->s_hist) l->spare.push_back(p);
    for (double *p : l->y_hist) l->spare.push_back(p);
    l->s_hist.clear();
    l->y_hist.clear();
    return OPL_OK;
}

int opl_lbfgs_set_stream(opl_lbfgs *l, void *cuda_stream) {
    if (!l) return fail(OPL_E_ARG, "opl_lbfgs_set_stream: null");
    if (l->stream != (cudaStream_t)cuda_stream) {
        OPL_CUDA(cudaStreamSynchronize(l->stream));
        l->stream = (cudaStream_t)cuda_stream;
    }
    return OPL_OK;
}

int opl_lbfgs_push_pair_device(opl_lbfgs *l, const double *s_dev, const double *y_dev) {
    if (!l || !s_dev || !y_dev) return fail(OPL_E_ARG, "opl_lbfgs_push_pair_device: null pointer");
    if (l->memory > 0 && (int)l->s_hist.size() >= l->memory) {
        l->spare.push_back(l->s_hist.front());
        l->spare.push_back(l->y_hist.front());
        l->s_hist.erase(l->s_hist.begin());
        l->y_hist.erase(l->y_hist.begin());
    }
    double *s = nullptr, *y = nullptr;
    int rc = grab(l, &s);
    if (rc == OPL_OK) rc = grab(l, &y);
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaMemcpyAsync(s, s_dev, sizeof(double) * (size_t)l->n, cudaMemcpyDeviceToDevice, l->stream));
    OPL_CUDA(cudaMemcpyAsync(y, y_dev, sizeof(double) * (size_t)l->n, cudaMemcpyDeviceToDevice, l->stream));
    l->s_hist.push_back(s);
    l->y_hist.push_back(y);
    return OPL_OK;
}

int opl_lbfgs_get_pair_device(const opl_lbfgs *l, int32_t index, double *s_out_dev, double *y_out_dev) {
    if (!l || !s_out_dev || !y_out_dev) return fail(OPL_E_ARG, "opl_lbfgs_get_pair_device: null pointer");
    if (index < 0 || index >= (int32_t)l->s_hist.size()) return fail(OPL_E_ARG, "opl_lbfgs_get_pair_device: index %d outside the history", index);
    OPL_CUDA(cudaMemcpyAsync(s_out_dev, l->s_hist[index], sizeof(double) * (size_t)l->n, cudaMemcpyDeviceToDevice, l->stream));
    OPL_CUDA(cudaMemcpyAsync(y_out_dev, l->y_hist[index], sizeof(double) * (size_t)l->n, cudaMemcpyDeviceToDevice, l->stream));
    return OPL_OK;
}

int32_t opl_lbfgs_history_len(const opl_lbfgs *l) { return l ? (int32_t)l->s_hist.size() : -1; }
int64_t opl_lbfgs_launch_count(const opl_lbfgs *l) { return l ? l->launches : -1; }

int opl_lbfgs_direction_device(opl_lbfgs *l, const double *grad_dev, const double *prev_step_dev,
                               const double *prev_grad_dev, int32_t gamma_mode, double *dir_out_dev) {
    if (!l || !grad_dev || !dir_out_dev) return fail(OPL_E_ARG, "opl_lbfgs_direction_device: null pointer");
    if (prev_step_dev && !prev_grad_dev) return fail(OPL_E_ARG, "opl_lbfgs_direction_device: prev_grad missing");
    const int64_t n = l->n;
    // _update_diffs (optimizer.py:433-450): pop the oldest if full, then append the new pair
    if (l->memory > 0 && (int)l->s_hist.size() >= l->memory) {
        l->spare.push_back(l->s_hist.front());
        l->spare.push_back(l->y_hist.front());
        l->s_hist.erase(l->s_hist.begin());
        l->y_hist.erase(l->y_hist.begin());
    }
    if (prev_step_dev) {
        double *s = nullptr, *y = nullptr;
        int rc = grab(l, &s);
        if (rc == OPL_OK) rc = grab(l, &y);
        if (rc != OPL_OK) return rc;
        OPL_CUDA(cudaMemcpyAsync(s, prev_step_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, l->stream));
        int g = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 4 * (int64_t)sm_count()));
        lbfgs_sub_kernel<<<g, 256, 0, l->stream>>>(n, grad_dev, prev_grad_dev, y);
        OPL_LAUNCH_CHECK();
        ++l->launches;
        l->s_hist.push_back(s);
        l->y_hist.push_back(y);
    }
    const int m = (int)l->s_hist.size();
    if (m > l->table_capacity) {
        const int cap = std::max(16, 2 * m);
        cudaFree((void *)l->table);
        l->table = nullptr;
        OPL_CUDA(cudaMalloc((void **)&l->table, sizeof(double *) * 2 * (size_t)cap));
        l->table_capacity = cap;
    }
    if (m > LB_MAX_SMEM_PAIRS && m > l->pair_capacity) {
        cudaFree(l->pair_scratch);
        l->pair_scratch = nullptr;
        OPL_CUDA(cudaMalloc(&l->pair_scratch, sizeof(double) * 2 * (size_t)(2 * m)));
        l->pair_capacity = 2 * m;
    }
    if (m > 0) {
        std::vector<const double *> host(2 * (size_t)m);
        for (int j = 0; j < m; ++j) {
            host[j] = l->s_hist[j];
            host[m + j] = l->y_hist[j];
        }
        // pageable source: the runtime stages it before returning, so `host` may die after this call
        OPL_CUDA(cudaMemcpyAsync((void *)l->table, host.data(), sizeof(double *) * 2 * (size_t)m,
                                 cudaMemcpyHostToDevice, l->stream));
    }
    LbfgsArgs a = {};
    a.n = n;
    a.m = m;
    a.S = (const double *const *)l->table;
    a.Y = (const double *const *)(l->table ? l->table + m : nullptr);
    a.g = grad_dev;
    a.q = l->q;
    a.dir = dir_out_dev;
    a.partial = l->partial;
    a.pair_scratch = l->pair_scratch;
    a.gamma_mode = gamma_mode;
    void *params[] = {(void *)&a};
    OPL_CUDA(cudaLaunchCooperativeKernel((const void *)lbfgs_two_loop_kernel, dim3((unsigned)l->grid),
                                         dim3(LB_THREADS), params, 0, l->stream));
    ++l->launches;
    return OPL_OK;
}

}  // extern "C"
