// RBF network front end: cluster distances / Gaussian similarities and the SOM that places the
// cluster centres.  The RBF output layer itself is the single-layer case of the MLP kernel schedule (its flat layout
// [bias ; W.ravel()], architecture/rbf.py:229-231, IS the MLP layout), so only these two pieces are new.
//
//   rbf_similarity_kernel  SOM.activate (architecture/som.py:61-85: sqrt(sum_k (x_k - w_jk)^2) for every pattern and
//                          neuron) + calculate.gaussian (calculate.py:101-102: exp(-(d^2 / variance))) + the similarity
//                          scaling of RBF.activate (architecture/rbf.py:139-146: divide by the row sum, NaN -> 1/clusters).
//                          Data-parallel over patterns: one warp per pattern, lanes over neurons; HBM-bound
//                          (rows * (attrs + neurons) * 8 bytes), centres through L1/L2.
//   som_train_kernel       the whole of Model.train_step's per-pattern loop (base.py:289-331) around
//                          SOM._train_increment / _move_neurons (som.py:87-113) for `passes` passes over the dataset.
//                          The algorithm is sequential by construction (every pattern moves the neurons the next pattern
//                          competes against), so it runs as ONE resident CTA with the neuron weights in shared memory; the
//                          parallelism is inside a pattern (neurons x attributes).  Products and sums of the update are
//                          rounded separately (w + fl(rate * fl(x - w))) as NumPy's temporaries are.
#include <algorithm>

#include "common.cuh"

namespace opl {

namespace {

constexpr int SIM_WARPS = 8;

__global__ void __launch_bounds__(SIM_WARPS * 32) rbf_similarity_kernel(const double *__restrict__ x, int64_t rows, int attrs,
                                                                         const double *__restrict__ centers, int neurons, int mode,
                                                                         double variance, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * SIM_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double *xr = x + row * attrs;
    double *o = out + row * neurons;
    double total = 0.0;
    for (int j = lane; j < neurons; j += 32) {
        const double *c = centers + (int64_t)j * attrs;
        double s = 0.0;
        for (int k = 0; k < attrs; ++k) {
            const double d = __dsub_rn(xr[k], c[k]);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
        double v = sqrt(s);
        if (mode != 0) v = exp(-(__dmul_rn(v, v) / variance));
        o[j] = v;
        total += v;
    }
    if (mode == 2) {
        total = warp_sum(total);
        const double uniform = 1.0 / (double)neurons;
        __syncwarp();
        for (int j = lane; j < neurons; j += 32) {
            const double v = o[j] / total;
            o[j] = isnan(v) ? uniform : v;          // 0 / 0 (every similarity underflowed): uniform vector, rbf.py:144-146
        }
    }
}

struct SomArgs {
    const double *x;
    int64_t rows;
    int attrs;
    double *weights;       // neurons x attrs, in / out
    int neurons;
    int64_t passes;
    const double *rates;   // neighborhood + 1 final move rates, by |i - winner| (computed by the host mirror exactly as
                           // the reference does: calculate.gaussian(distance, neighbor_move_rate) * move_rate)
    int neighborhood;
    int in_smem;           // weights fit in shared memory
};

__global__ void __launch_bounds__(256) som_train_kernel(const SomArgs p) {
    extern __shared__ double som_sm[];
    double *xs = som_sm;                              // attrs
    double *dist = xs + p.attrs;                      // neurons
    double *w = p.in_smem ? dist + p.neurons : p.weights;
    __shared__ int winner;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int64_t wcount = (int64_t)p.neurons * p.attrs;
    if (p.in_smem)
        for (int64_t i = tid; i < wcount; i += nthreads) w[i] = p.weights[i];
    __syncthreads();
    for (int64_t pass = 0; pass < p.passes; ++pass) {
        for (int64_t r = 0; r < p.rows; ++r) {
            for (int k = tid; k < p.attrs; k += nthreads) xs[k] = p.x[r * p.attrs + k];
            __syncthreads();
            // distances to every neuron (som.py:66-71)
            for (int j = tid; j < p.neurons; j += nthreads) {
                const double *wj = w + (int64_t)j * p.attrs;
                double s = 0.0;
                for (int k = 0; k < p.attrs; ++k) {
                    const double d = __dsub_rn(xs[k], wj[k]);
                    s = __dadd_rn(s, __dmul_rn(d, d));
                }
                dist[j] = sqrt(s);
            }
            __syncthreads();
            // numpy.argmin (som.py:102): the FIRST index of the minimum; a NaN counts as the minimum (key -inf: distances
            // are never negative)
            if (tid < 32) {
                double best = INFINITY;
                int at = 0x7fffffff;
                for (int j = tid; j < p.neurons; j += 32) {
                    const double d = dist[j], key = isnan(d) ? -INFINITY : d;
                    if (at == 0x7fffffff || key < best) {
                        best = key;
                        at = j;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, at, o);
                    if (oa != 0x7fffffff && (at == 0x7fffffff || ob < best || (ob == best && oa < at))) {
                        best = ob;
                        at = oa;
                    }
                }
                if (tid == 0) winner = at;
            }
            __syncthreads();
            // move the winner and its neighbours towards the pattern (som.py:104-113)
            const int lo = max(0, winner - p.neighborhood), hi = min(p.neurons - 1, winner + p.neighborhood);
            const int span = (hi - lo + 1) * p.attrs;
            for (int i = tid; i < span; i += nthreads) {
                const int j = lo + i / p.attrs, k = i - (j - lo) * p.attrs;
                const double rate = p.rates[abs(j - winner)];
                double *wp = w + (int64_t)j * p.attrs + k;
                *wp = __dadd_rn(*wp, __dmul_rn(rate, __dsub_rn(xs[k], *wp)));
            }
            __syncthreads();
        }
    }
    if (p.in_smem)
        for (int64_t i = tid; i < wcount; i += nthreads) p.weights[i] = w[i];
}

}  // namespace

}  // namespace opl

extern "C" {

int opl_rbf_similarity_device(const double *x_dev, int64_t rows, int32_t attrs, const double *centers_dev, int32_t neurons,
                              int32_t mode, double variance, double *out_dev, void *cuda_stream) {
    using namespace opl;
    if (rows < 0 || attrs < 1 || neurons < 1 || mode < 0 || mode > 2 || !centers_dev || (rows > 0 && (!x_dev || !out_dev)))
        return fail(OPL_E_ARG, "opl_rbf_similarity_device: bad arguments");
    if (rows == 0) return OPL_OK;
    rbf_similarity_kernel<<<(unsigned)ceil_div(rows, SIM_WARPS), SIM_WARPS * 32, 0, (cudaStream_t)cuda_stream>>>(
        x_dev, rows, attrs, centers_dev, neurons, mode, variance, out_dev);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

int opl_som_train_device(const double *x_dev, int64_t rows, int32_t attrs, double *weights_dev, int32_t neurons, int64_t passes,
                         const double *rates_dev, int32_t neighborhood, void *cuda_stream) {
    using namespace opl;
    if (rows < 0 || attrs < 1 || neurons < 1 || passes < 0 || neighborhood < 0 || !weights_dev || !rates_dev || (rows > 0 && !x_dev))
        return fail(OPL_E_ARG, "opl_som_train_device: bad arguments");
    if (rows == 0 || passes == 0) return OPL_OK;
    SomArgs a;
    a.x = x_dev;
    a.rows = rows;
    a.attrs = attrs;
    a.weights = weights_dev;
    a.neurons = neurons;
    a.passes = passes;
    a.rates = rates_dev;
    a.neighborhood = neighborhood;
    size_t smem = sizeof(double) * ((size_t)attrs + neurons);
    const size_t with_w = smem + sizeof(double) * (size_t)attrs * neurons;
    a.in_smem = with_w <= 200 * 1024 ? 1 : 0;
    if (a.in_smem) smem = with_w;
    if (smem > 200 * 1024) return fail(OPL_E_ARG, "opl_som_train_device: attributes + neurons = %lld exceed the shared staging area",
                                       (long long)attrs + neurons);
    if (smem > 48 * 1024) OPL_CUDA(cudaFuncSetAttribute(som_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    som_train_kernel<<<1, 256, smem, (cudaStream_t)cuda_stream>>>(a);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

}  // extern "C"
