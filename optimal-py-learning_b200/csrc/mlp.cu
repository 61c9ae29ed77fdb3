// MLP objective / gradient workspace (parity mode, FP64).
//
// Replaces MLP.activate / _get_obj / _get_obj_jac / _get_jacobians of the reference
// (learning/architecture/mlp.py:134-163, 195-276) with a fixed kernel schedule per evaluation:
//
//   trial point   w = x + alpha*dir                                  (linesearch.py:394)
//   forward       A_{l+1} = f_l(A_l W_l [+ b]);  D_l = f_l'(Z_l)     NN GEMM, fused epilogue
//   output rows   softmax / elementwise transfer, MSE / cross entropy, delta_L, loss partials
//   backward      dW_l = A_l^T delta_l        TN split-K GEMM -> fixed-order reduce into the
//                                             flat gradient (mlp.py:313-315 layout)
//                 delta_{l-1} = (delta_l W_l^T) o D_{l-1}            NT GEMM, fused epilogue
//                 db = 1^T delta_0                                   two-stage column sum
//   loss          obj = -(sum/N)  or  sum/(N*C)                      -> grad_out[n]
//
// Everything is enqueued on one stream; the only host sync is the scalar readback in
// opl_mlp_probe_scalars / the *_host entry points.
#include <algorithm>
#include <cstring>
#include <vector>

#include "gemm_f64.cuh"
#include "gemm_tf32.cuh"
#include "ozaki.cuh"
#include "oz_transcend.cuh"

namespace opl {

__global__ void axpy_kernel(int64_t n, const double *__restrict__ x, double a, const double *__restrict__ d,
                            double b, const double *__restrict__ p, double *__restrict__ out);

// ---------------------------------------------------------------------------------------
// row-wise kernels: `group` lanes (power of two <= 32) cooperate on one pattern
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double group_sum(double v, int group) {
    for (int o = group >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double group_max(double v, int group) {
    for (int o = group >> 1; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ double transfer_value(int tid, double tparam, double z) {
    switch (tid) {
        case OPL_TRANSFER_TANH: return tanh(z);
        case OPL_TRANSFER_SOFTPLUS: {
            double r = log(1.0 + exp(z));
            return isinf(r) ? z : r;
        }
        case OPL_TRANSFER_GAUSSIAN: return exp(-(z * z / tparam));
        case OPL_TRANSFER_LOGIT: return 1.0 / (1.0 + exp(-z));   // calculate.py:65-67
        default: return z;
    }
}
__device__ __forceinline__ double transfer_slope(int tid, double tparam, double z, double a) {
    switch (tid) {
        case OPL_TRANSFER_TANH: return 1.0 - a * a;
        case OPL_TRANSFER_SOFTPLUS: return 1.0 / (1.0 + exp(-z));
        case OPL_TRANSFER_GAUSSIAN: return -2.0 * z * a / tparam;
        case OPL_TRANSFER_LOGIT: {   // calculate.dlogit (calculate.py:70-88): e^z/(e^z+1)^2, and 1 where e^z overflows
            const double e = exp(z);
            return isinf(e) ? 1.0 : e / ((e + 1.0) * (e + 1.0));
        }
        default: return 1.0;
    }
}

// T = storage type of the activation-side matrices: double (parity mode) or float (fast mode; values that
// feed a TF32 GEMM are stored pre-rounded to TF32 so the tensor core's truncation is exact).
template <typename T>
struct RowArgsT {
    const T *Z;        // rows x C pre-activations, row stride ld
    const double *Y;   // rows x C targets (dense, stride C), or null (activate only)
    const T *G;        // explicit upstream gradient instead of an error function, or null (stride ld)
    T *A;              // activations out, or null (may alias Z) (stride ld)
    T *D;              // delta out, or null (stride ld)
    int64_t rows;
    int64_t ld;
    int C;
    int tid;           // output transfer
    double tparam;
    int eid;           // error id
    double ce_div;     // -(double)N          (error.py:115-116)
    double mse_mul;    // 2.0 / (N * C)       (error.py:62)
    double *loss_partial;   // one per block, or null
    int *flag;         // set to 1 on log(negative)
    int group;
};
using RowArgs = RowArgsT<double>;

__device__ __forceinline__ void store_operand(double *p, double v) { *p = v; }
__device__ __forceinline__ void store_operand(float *p, double v) { *p = round_tf32((float)v); }

// error term for one output element: returns the upstream gradient, adds to `loss`
template <typename T>
__device__ __forceinline__ double error_term(const RowArgsT<T> &r, double a, double y, double &loss, bool &bad) {
    if (r.eid == OPL_ERROR_MSE) {   // error.py:58-64
        double d = a - y;
        loss += d * d;
        return d * r.mse_mul;
    }
    // cross entropy, error.py:89-116
    if (a < 0.0) bad = true;
    loss += nan_to_num(log(a)) * y;
    return nan_to_num(y / a) / r.ce_div;
}

template <typename T>
__global__ void __launch_bounds__(256) output_rows_kernel(const RowArgsT<T> r) {
    __shared__ double scratch[33];
    const int group = r.group;
    const int gl = threadIdx.x & (group - 1);
    const int rows_per_block = blockDim.x / group;
    const int slot = threadIdx.x / group;
    const bool softmax = r.tid == OPL_TRANSFER_SOFTMAX;
    double loss = 0.0;
    bool bad = false;
    const int64_t nblocks = (r.rows + rows_per_block - 1) / rows_per_block;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row = blk * rows_per_block + slot;
        const bool live = row < r.rows;
        const T *z = r.Z + (live ? row : 0) * r.ld;
        const double *y = r.Y ? r.Y + (live ? row : 0) * (int64_t)r.C : nullptr;
        double mx = -INFINITY, den = 1.0;
        if (softmax) {   // calculate.py:152-153
            if (live)
                for (int c = gl; c < r.C; c += group) mx = fmax(mx, (double)z[c]);
            mx = group_max(mx, group);
            double s = 0.0;
            if (live)
                for (int c = gl; c < r.C; c += group) s += exp((double)z[c] - mx);
            den = group_sum(s, group);
        }
        double t = 0.0;
        if (live) {
            for (int c = gl; c < r.C; c += group) {
                const double zc = (double)z[c];
                const double a = softmax ? exp(zc - mx) / den : transfer_value(r.tid, r.tparam, zc);
                if (y || r.G) {
                    const double g = r.G ? (double)r.G[row * r.ld + c] : error_term(r, a, y[c], loss, bad);
                    if (softmax) t += g * a;
                    else if (r.D) store_operand(&r.D[row * r.ld + c], g * transfer_slope(r.tid, r.tparam, zc, a));
                }
                if (r.A && !(softmax && r.D && r.A == r.Z)) store_operand(&r.A[row * r.ld + c], a);
            }
        }
        if (softmax && r.D && (y || r.G)) {   // dsoftmax contracted with the upstream gradient (mlp.py:295): softmax_delta()
            t = group_sum(t, group);
            if (live) {
                double dummy = 0.0;
                bool dummy_bad = false;
                for (int c = gl; c < r.C; c += group) {
                    const double a = exp((double)z[c] - mx) / den;
                    const double g = r.G ? (double)r.G[row * r.ld + c] : error_term(r, a, y[c], dummy, dummy_bad);
                    store_operand(&r.D[row * r.ld + c], softmax_delta(a, g, t));
                    if (r.A && r.A == r.Z) store_operand(&r.A[row * r.ld + c], a);
                }
            }
        }
    }
    if (bad) *r.flag = 1;
    if (r.loss_partial) {
        loss = block_sum(loss, scratch);
        if (threadIdx.x == 0) r.loss_partial[blockIdx.x] = loss;
    }
}

// Same stage for narrow outputs (C <= MAXC): one thread owns a whole pattern in registers -- no
// shuffles, no re-reads, every transcendental evaluated once.  Classification heads (3..32
// classes) take this path.
template <int MAXC, typename T>
__global__ void __launch_bounds__(256) output_rows_thread_kernel(const RowArgsT<T> r) {
    __shared__ double scratch[33];
    const bool softmax = r.tid == OPL_TRANSFER_SOFTMAX;
    const int C = r.C;
    double loss = 0.0;
    bool bad = false;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < r.rows;
         row += (int64_t)gridDim.x * blockDim.x) {
        const T *zp = r.Z + row * r.ld;
        // Rolled loops on purpose (the unrolled version inlined 16 copies of exp/log/div and stalled on
        // instruction fetch); the per-pattern values live in a small thread-local array.
        double a[MAXC], g[MAXC];
        double mx = -INFINITY, den = 0.0, t = 0.0;
        if (softmax) {
#pragma unroll 1
            for (int c = 0; c < C; ++c) {
                a[c] = (double)zp[c];
                mx = fmax(mx, a[c]);
            }
#pragma unroll 1
            for (int c = 0; c < C; ++c) {
                a[c] = exp(a[c] - mx);
                den += a[c];
            }
        }
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const double zc = (double)zp[c];
            const double ac = softmax ? a[c] / den : transfer_value(r.tid, r.tparam, zc);
            a[c] = ac;
            if (r.A) store_operand(&r.A[row * r.ld + c], ac);
            if (r.Y || r.G) {
                const double gc = r.G ? (double)r.G[row * r.ld + c] : error_term(r, ac, r.Y[row * C + c], loss, bad);
                g[c] = gc;
                if (softmax) t += gc * ac;
                else if (r.D) store_operand(&r.D[row * r.ld + c], gc * transfer_slope(r.tid, r.tparam, zc, ac));
            }
        }
        if (softmax && r.D && (r.Y || r.G)) {
#pragma unroll 1
            for (int c = 0; c < C; ++c) store_operand(&r.D[row * r.ld + c], softmax_delta(a[c], g[c], t));
        }
    }
    if (bad) *r.flag = 1;
    if (r.loss_partial) {
        loss = block_sum(loss, scratch);
        if (threadIdx.x == 0) r.loss_partial[blockIdx.x] = loss;
    }
}

// hidden softmax layer backward: delta_k = softmax_delta(a_k, g_k, sum_j g_j a_j), in place over g
template <typename T>
__global__ void __launch_bounds__(256)
softmax_backward_rows_kernel(T *G, const T *A, int64_t rows, int64_t ld, int C, int group) {
    const int gl = threadIdx.x & (group - 1);
    const int rows_per_block = blockDim.x / group;
    const int slot = threadIdx.x / group;
    const int64_t nblocks = (rows + rows_per_block - 1) / rows_per_block;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row = blk * rows_per_block + slot;
        const bool live = row < rows;
        double t = 0.0;
        if (live)
            for (int c = gl; c < C; c += group) t += (double)G[row * ld + c] * (double)A[row * ld + c];
        t = group_sum(t, group);
        if (live)
            for (int c = gl; c < C; c += group)
                store_operand(&G[row * ld + c], softmax_delta((double)A[row * ld + c], (double)G[row * ld + c], t));
    }
}

// delta_prev[r][j] = (sum_c delta[r][c] * W[j][c]) * slope[r][j] for a NARROW next layer (C <= 16): the
// "GEMM" has K = C, so it is pure streaming over the rows x H slope / output matrices.  One CTA takes a
// block of rows; W (H x C) and the block's delta rows sit in shared memory; thread j walks the rows, so
// global traffic is perfectly coalesced along j with 8 independent loads in flight per thread.
// mode: 0 out = s, 1 out = s * aux (D = f'), 2 out = s * (1 - aux^2) (tanh from the activation)
template <int MAXC>
// colmax (optional): max |out[:, j]| per column, gathered with atomicMax on the bit patterns (non-negative doubles order
// like unsigned integers; a non-finite entry stores the NaN pattern, which is larger than every finite one).
__global__ void __launch_bounds__(256) narrow_backprop_kernel(const double *__restrict__ delta, const double *__restrict__ W,
                                                              const double *aux, double *out, int64_t rows, int H, int C,
                                                              int rows_per_block, int mode, double *colmax) {
    extern __shared__ double sm[];
    double *Ws = sm;                       // H x (C|1): odd stride => conflict-free across j
    const int cp = C | 1;
    double *Ds = sm + (size_t)H * cp;      // rows_per_block x C
    for (int i = threadIdx.x; i < H * C; i += blockDim.x) Ws[(i / C) * cp + (i % C)] = W[i];
    // persistent CTAs over small row blocks: W is staged once per CTA and the tail is a few rows, not a wave
    const int64_t nblocks = (rows + rows_per_block - 1) / rows_per_block;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t r0 = blk * rows_per_block;
        const int nr = (int)min((int64_t)rows_per_block, rows - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * C; i += blockDim.x) Ds[i] = delta[r0 * C + i];
        __syncthreads();
        for (int j = threadIdx.x; j < H; j += blockDim.x) {
            double w[MAXC];
            double cmax = 0.0;
            bool bad = false;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) w[c] = c < C ? Ws[j * cp + c] : 0.0;
            for (int rb = 0; rb < nr; rb += 8) {
                double x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = (mode != 0 && rb + u < nr) ? aux[(r0 + rb + u) * H + j] : 1.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (rb + u < nr) {
                        double s_ = 0.0;
#pragma unroll
                        for (int c = 0; c < MAXC; ++c)
                            if (c < C) s_ += Ds[(rb + u) * C + c] * w[c];
                        const double o = mode == 2 ? s_ * (1.0 - x[u] * x[u]) : s_ * x[u];
                        out[(r0 + rb + u) * H + j] = o;
                        const double ao = fabs(o);
                        bad |= !(ao <= 1.7976931348623157e308);
                        cmax = fmax(cmax, ao);
                    }
                }
            }
            if (colmax)
                atomicMax(reinterpret_cast<unsigned long long *>(colmax) + j,
                          bad ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(cmax));
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fused narrow head: the last layer (H hidden units -> C <= 16 outputs) of the forward pass, the output transfer and
// error function, and the head's share of back-propagation, in ONE streaming pass over the hidden activations:
//     Z = A W,  a = f(Z),  loss,  delta_L = dE/dZ,  dW += A^T delta_L,  delta_prev = (delta_L W^T) o slope
// The five separate launches re-read A twice and move the logits / deltas through HBM; here A and the slope matrix
// are read once and delta_prev is written once.  A CTA owns a fixed sequence of row blocks and a fixed-order
// accumulation, so the result is bit-reproducible; per-CTA dW partials are folded by the split-K reduce kernel.
// mode (slope of the hidden transfer): 0 none, 1 aux = f' stored by the forward GEMM, 2 aux = tanh activation.
// ---------------------------------------------------------------------------------------
// T = storage type of the activation-side matrices: double (parity mode, narrow_head_kernel below) or float (fast mode:
// TF32-rounded operands, narrow_head_tf32_kernel in mlp_fast.inl)
template <typename T>
struct HeadArgsT {
    const T *A;             // rows x H hidden activations, row stride ld
    const double *W;        // H x C
    const double *Y;        // rows x C targets
    const T *aux;           // rows x H (mode 1: f', mode 2: the activation itself) or null, row stride ld
    T *delta_prev;          // rows x H out (may alias aux), or null (objective only), row stride ld
    double *dw_partial;     // gridDim.x x (H*C), or null
    double *loss_partial;   // gridDim.x
    double *colmax;         // H: max |delta_prev[:, j]| (atomicMax on bit patterns) or null
    double *colsum_partial; // gridDim.x x H: column sums of delta_prev (the bias gradient when delta_prev is delta_0) or null
    double *delta_out;      // DA == false: rows x C delta_L written out (delta_prev is then formed by the digit kernel), else unused
    double *classmax;       // DA == false: C entries, max |delta_L[:, c]| (atomicMax on bit patterns)
    T *deriv_out;           // ZIN: rows x H slope values f'(z) out (row stride ld), or null (objective only)
    int64_t rows;
    int64_t ld;
    int H, C, R;            // R rows per block step
    int mode;
    int tid;                // output transfer
    double tparam;
    int eid;
    double ce_div, mse_mul;
    double ce_rdiv;         // 1 / ce_div
    int *flag;
    long long *trace;       // tools only (OPL_HEAD_TRACE): CTA 0 writes 8 clock64() stamps per step
};
using HeadArgs = HeadArgsT<double>;

__device__ __forceinline__ double2 head_load2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 head_load2(const float *p) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    return make_double2((double)v.x, (double)v.y);
}
// stores the pair as the operand type requires and returns what was stored
__device__ __forceinline__ double2 head_store2(double *p, double a, double b) {
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
    return make_double2(a, b);
}
__device__ __forceinline__ double2 head_store2(float *p, double a, double b) {
    const float2 v = make_float2(round_tf32((float)a), round_tf32((float)b));
    *reinterpret_cast<float2 *>(p) = v;
    return make_double2((double)v.x, (double)v.y);
}

// All three products are DMMA.8x8x4 on shared-memory fragments (classes padded to 16 with zeros):
//   logits  Z[R x 16]   = As[R x H] . Ws[H x 16]            (row tile, class tile, K half) per warp, halves meet in Zs
//   dA      D0[R x H]   = Ds[R x 16] . Ws^T[16 x H]          warp -> hidden tiles w, w+16, ..; epilogue x slope, store
//   dW      G[H x 16]  += As^T[H x R] . Ds[R x 16]           warp -> hidden tiles w, w+16, ..; accumulators stay in registers
// One CTA of 16 warps per SM, R = 32 patterns per step.  The next step's activation tile streams in with cp.async
// (double buffer) and the step's slope values are requested before the logits are computed, so every global load has
// a whole phase of compute to hide behind.  Leading dimensions (H+4, 20) make every fragment load conflict-free.
// The row stage (softmax / error / delta_L) runs one pattern per half-warp, lane = class.
// Two shapes: R = 32 patterns per step with 16 warps (one CTA per SM) or R = 16 with 8 warps (two CTAs per SM, whose
// latency-bound row stages and DMMA phases overlap each other).
// tools only (OPL_HEAD_TRACE): phase stamps of CTA 0, thread 0; evaluated in place so that nothing stays live
// (compiled in with -DOPL_HEAD_TRACE_BUILD only: the eight predicate tests per step were 5 % of the kernel's instructions)
#ifdef OPL_HEAD_TRACE_BUILD
#define HEAD_STAMP(k)                                                                       \
    do {                                                                                    \
        if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && step_no < 16) p.trace[step_no * 8 + (k)] = clock64(); \
    } while (0)
#else
#define HEAD_STAMP(k) ((void)0)
#endif
// DA == false: the head does NOT form delta_prev = (delta_L W^T) o slope (a third of its DMMA work, the slope matrix read and the
// delta_prev write): it writes delta_L (rows x C) and the per-class maxima instead, and the digit kernel of the weight-gradient
// GEMM forms delta_prev on the fly (ozaki_slice_delta_prev, csrc/ozaki.cu).  `backward` then only means "accumulate dW".
// ZIN: `A` holds the hidden PRE-activations z (the forward digit GEMM ran with a plain-store epilogue) and the head applies the
// hidden softplus itself: every landed tile is transformed in place (same routine, same arguments, hence the same bits as the GEMM
// epilogue would have produced) and the slope values f'(z) go out to `deriv_out` for the digit kernel.  Why here: FP64
// instructions do not overlap tcgen05.mma (profiles/r02_ozaki_epilogue_decomposition.txt) -- in the GEMM epilogue the 27 FP64
// instructions per element are exposed time on top of the MMAs, here they fill FP64 issue slots this latency-bound kernel leaves idle.
template <typename T, int MT, int R, int HEAD_WARPS, bool CSUM, bool DA = true, bool ZIN = false>    // MT = hidden tiles (8 units) per warp: H <= 8 MT HEAD_WARPS
__global__ void __launch_bounds__(HEAD_WARPS * 32, R == 16 ? 2 : 1) narrow_head_kernel(const HeadArgsT<T> p) {
    static_assert(!ZIN || (sizeof(T) == 8 && !DA), "the late hidden transfer exists for the FP64 head without delta_prev");
    extern __shared__ double hsm[];
    __shared__ double scratch[33];
    constexpr int RM = R / 8, HEAD_THREADS = HEAD_WARPS * 32;
    constexpr int EPC = 16 / (int)sizeof(T);   // elements per 16-byte chunk
    const int H = p.H, C = p.C;
    const int ldA = H + (sizeof(T) == 8 ? 4 : 8);   // conflict-free fragment loads for 8-byte / 4-byte elements
    constexpr int ldW = 20, ldD = 20;
    double *Ws = hsm;                          // H x 20
    double *Zs = Ws + (size_t)H * ldW;         // R x 20: logits, then delta_L
    T *As0 = reinterpret_cast<T *>(Zs + R * ldD);   // 2 x R x ldA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bool softmax = p.tid == OPL_TRANSFER_SOFTMAX;
    const bool backward = DA ? p.delta_prev != nullptr : p.dw_partial != nullptr;
    const int h2 = H / EPC;                    // 16-byte chunks per row (H is a multiple of 8)
    double lmax = 0.0;                         // DA == false: max |delta_L| of this lane's class over the CTA's patterns
    bool lbad = false;
    const int64_t nblocks = (p.rows + R - 1) / R;
    auto fetch = [&](int64_t blk, int buf) {   // cp.async the block's activations (rows past the end := 0)
        // a warp per row, lanes along the row: no index division, one address per row (the i / h2 form of this loop was 38 %
        // of the kernel's instructions; profiles/r02_head_instruction_mix.txt)
        const int64_t r0 = blk * R;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(As0) + (unsigned)(buf * R * ldA) * (unsigned)sizeof(T);
#pragma unroll
        for (int r = warp; r < R; r += HEAD_WARPS) {
            const bool ok = r0 + r < p.rows;
            const T *src = p.A + (ok ? (r0 + r) * p.ld : 0) + EPC * lane;
            unsigned d = dst + (unsigned)(r * ldA + EPC * lane) * (unsigned)sizeof(T);
            for (int c2 = lane; c2 < h2; c2 += 32, src += EPC * 32, d += 512) cp_async16_s(d, ok ? src : p.A, ok);
        }
        cp_async_commit();
    };
    if ((int64_t)blockIdx.x < nblocks) fetch(blockIdx.x, 0);
    for (int i = threadIdx.x; i < H * 16; i += HEAD_THREADS) {
        const int h = i >> 4, c = i & 15;
        Ws[h * ldW + c] = c < C ? p.W[h * C + c] : 0.0;
    }
    RowArgs ra = {};                           // error_term() reads its constants from here
    ra.eid = p.eid;
    ra.ce_div = p.ce_div;
    ra.mse_mul = p.mse_mul;
    double loss = 0.0;
    bool bad = false, cbad = false;
    double dw[MT][2][2], cmax[MT][2], csum[CSUM ? MT : 1][2];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        cmax[i][0] = cmax[i][1] = 0.0;
        if (CSUM) csum[CSUM ? i : 0][0] = csum[CSUM ? i : 0][1] = 0.0;
#pragma unroll
        for (int n = 0; n < 2; ++n) dw[i][n][0] = dw[i][n][1] = 0.0;
    }
    int buf = 0, step_no = 0;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, buf ^= 1, ++step_no) {
        HEAD_STAMP(0);
        const int64_t r0 = blk * R;
        const int nr = (int)min((int64_t)R, p.rows - r0);
        const T *As = As0 + buf * R * ldA;
        // slope values of this step's delta_prev tiles: requested now, used after the row stage
        double2 x[DA ? MT : 1][RM];
        if (DA && backward && p.mode != 0) {
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int col = 8 * (warp + HEAD_WARPS * i) + 2 * t;
#pragma unroll
                for (int mt = 0; mt < RM; ++mt) {
                    const int row = 8 * mt + g;
                    x[i][mt] = (col < H && row < nr) ? head_load2(p.aux + (r0 + row) * p.ld + col) : make_double2(1.0, 1.0);
                }
            }
        }
        // this thread's target value for the row stage (one pattern per half-warp, lane = class): requested now, so the
        // row stage does not wait for DRAM
        double y_mine = 0.0;
        {
            const int r = warp * 2 + (lane >> 4), cl = lane & 15;
            if (r < nr && cl < C) y_mine = p.Y[(r0 + r) * C + cl];
        }
        for (int i = threadIdx.x; i < R * 16; i += HEAD_THREADS) Zs[(i >> 4) * ldD + (i & 15)] = 0.0;
        cp_async_wait<0>();
        __syncthreads();                       // tile `buf` landed, Zs cleared, previous step done with the other buffer
        HEAD_STAMP(1);
        if (blk + gridDim.x < nblocks) fetch(blk + gridDim.x, buf ^ 1);
        if (ZIN) {
            // ---- hidden transfer, in place: a warp per row, 8 elements (four 16-byte chunks) per lane and pass so that eight
            //      FP64 chains interleave; rows past the end stay zero ----
            double *Az = reinterpret_cast<double *>(As0) + buf * R * ldA;
            for (int r = warp; r < nr; r += HEAD_WARPS) {
                double *row = Az + r * ldA;
                double *dg = p.deriv_out ? reinterpret_cast<double *>(p.deriv_out) + (r0 + r) * p.ld : nullptr;
                for (int cb = 0; cb < H; cb += 256) {      // warp-uniform trip count (the vote below needs every lane)
                    const int c = cb + 2 * lane;
                    double zz[8], sp[8], sg[8];
                    bool regular = true;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cq = c + 64 * q;
                        const double2 v = cq < H ? *reinterpret_cast<const double2 *>(row + cq) : make_double2(0.0, 0.0);
                        zz[2 * q] = v.x;
                        zz[2 * q + 1] = v.y;
                        regular &= ozt_is_regular(v.x) && ozt_is_regular(v.y);
                    }
                    if (__all_sync(0xffffffffu, regular)) ozt_softplus_logistic_regular_n<8>(zz, sp, sg);
                    else ozt_softplus_logistic_n<8>(zz, sp, sg);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cq = c + 64 * q;
                        if (cq < H) {
                            *reinterpret_cast<double2 *>(row + cq) = make_double2(sp[2 * q], sp[2 * q + 1]);
                            if (dg) *reinterpret_cast<double2 *>(dg + cq) = make_double2(sg[2 * q], sg[2 * q + 1]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // ---- logits: warp -> (row tile, class tile, K half); the two halves add up in Zs ----
        {
            const int mt = warp % RM, nt = (warp / RM) & 1, kh = warp / (2 * RM);
            // two independent accumulation chains (four or eight measure the same: the phase is bound by the FP64 pipe
            // the two resident CTAs share, 3.9 cycles per DMMA per SM, not by the chain latency)
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            const int kq = H >> 3;                                // DMMA steps per half
            const T *ap = As + (8 * mt + g) * ldA + 4 * kh * kq + t;
            const double *bp = Ws + (4 * kh * kq + t) * ldW + 8 * nt + g;
            int k = 0;
#pragma unroll 4
            for (; k + 1 < kq; k += 2, ap += 8, bp += 8 * ldW) {
                dmma_8x8x4(c0, c1, (double)ap[0], bp[0]);
                dmma_8x8x4(e0, e1, (double)ap[4], bp[4 * ldW]);
            }
            if (k < kq) dmma_8x8x4(c0, c1, (double)ap[0], bp[0]);
            c0 += e0;
            c1 += e1;
            double *zp = Zs + (8 * mt + g) * ldD + 8 * nt + 2 * t;
            atomicAdd(zp, c0);                  // 0 + a + b == 0 + b + a: the order of the two halves does not matter
            atomicAdd(zp + 1, c1);
        }
        HEAD_STAMP(2);
        __syncthreads();
        HEAD_STAMP(3);
        // ---- output transfer + error + delta_L: one pattern per half-warp, lane = class ----
        {
            static_assert(HEAD_WARPS * 2 == R && 4 * RM == HEAD_WARPS, "one pattern per half-warp, two K halves per logits tile");
            const int hw = lane >> 4, cl = lane & 15;
            const int r = warp * 2 + hw;
            const bool live = r < nr && cl < C;
            const double z = Zs[r * ldD + cl];
            double a_out;
            if (softmax && p.eid == OPL_ERROR_CROSS_ENTROPY) {
                // parity mode, classification head.  The row stage is one dependent FP64 chain per pattern (~36 cycles per
                // link), so exp / log / reciprocal are the short table-driven forms of oz_transcend.cuh (~35 links instead
                // of ~110 for the library routines; ~1 ulp each).  Same formulas and the same numpy.nan_to_num rules as
                // error.py:89-116: log(0) -> -DBL_MAX, y/0 -> +-DBL_MAX, 0/0 and NaN -> 0.
                const double mx = group_max(live ? z : -INFINITY, 16);
                const double e = live ? ozt_exp_nonpos(z - mx) : 0.0;
                const double den = group_sum(e, 16);                         // in [1, 16] (or NaN)
                a_out = e * ozt_rcp(den);
                double gq = 0.0;
                if (live) {
                    const bool regular = a_out >= 9.332636185032189e-302;  // 2^-1000: normal, reciprocal finite
                    const double lg = regular ? ozt_log_pos(a_out) : (a_out > 0.0 ? log(a_out) : (a_out == 0.0 ? -kBig : 0.0));
                    loss += lg * y_mine;
                    const double q = regular ? y_mine * ozt_rcp(a_out) : nan_to_num(y_mine / a_out);
                    gq = q * p.ce_rdiv;
                }
                const double tt = group_sum(live ? gq * a_out : 0.0, 16);
                if (backward) Zs[r * ldD + cl] = live ? softmax_delta(a_out, gq, tt) : 0.0;
                HEAD_STAMP(4);
                __syncthreads();
                if (!backward) continue;
                goto head_backward;
            }
            if (softmax) {   // calculate.py:152-153
                const double mx = group_max(live ? z : -INFINITY, 16);
                const double e = live ? exp(z - mx) : 0.0;
                a_out = e / group_sum(e, 16);
            } else {
                a_out = live ? transfer_value(p.tid, p.tparam, z) : 0.0;
            }
            double gq = 0.0;
            if (live) gq = error_term(ra, a_out, y_mine, loss, bad);
            double d;
            if (softmax) {   // dsoftmax contracted with the upstream gradient (mlp.py:295): softmax_delta()
                const double tt = group_sum(live ? gq * a_out : 0.0, 16);
                d = softmax_delta(a_out, gq, tt);
            } else {
                d = live ? gq * transfer_slope(p.tid, p.tparam, z, a_out) : 0.0;
            }
            if (backward) Zs[r * ldD + cl] = live ? d : 0.0;
        }
        HEAD_STAMP(4);
        __syncthreads();
        if (!backward) continue;
    head_backward:
        HEAD_STAMP(5);
        if (!DA) {   // delta_L of this step (Zs, zero outside the live patterns / classes) -> global, and its per-class maxima
            const int hw = lane >> 4, cl = lane & 15, rr = warp * 2 + hw;
            if (rr < nr && cl < C) {
                const double d = Zs[rr * ldD + cl];
                p.delta_out[(r0 + rr) * C + cl] = d;
                const double ad = fabs(d);
                lbad |= !(ad <= 1.7976931348623157e308);
                lmax = fmax(lmax, ad);
            }
        }
        // ---- delta_prev = (delta_L W^T) o slope ----
#pragma unroll
        for (int i = 0; i < (DA ? MT : 0); ++i) {
            const int nt = warp + HEAD_WARPS * i;        // hidden tile
            if (8 * nt >= H) break;
            const int col = 8 * nt + 2 * t;
            const double *bp = Ws + (size_t)(8 * nt + g) * ldW + t;
#pragma unroll
            for (int mt = 0; mt < RM; ++mt) {
                const int row = 8 * mt + g;
                double c0 = 0.0, c1 = 0.0;
                const double *ap = Zs + row * ldD + t;
#pragma unroll
                for (int k = 0; k < 4; ++k) dmma_8x8x4(c0, c1, ap[4 * k], bp[4 * k]);
                if (row < nr) {
                    double o0 = c0, o1 = c1;
                    if (p.mode == 1) {
                        o0 *= x[DA ? i : 0][mt].x;
                        o1 *= x[DA ? i : 0][mt].y;
                    } else if (p.mode == 2) {
                        o0 *= 1.0 - x[DA ? i : 0][mt].x * x[DA ? i : 0][mt].x;
                        o1 *= 1.0 - x[DA ? i : 0][mt].y * x[DA ? i : 0][mt].y;
                    }
                    const double2 st = head_store2(p.delta_prev + (r0 + row) * p.ld + col, o0, o1);
                    const double a0 = fabs(st.x), a1 = fabs(st.y);
                    cbad |= !(a0 <= 1.7976931348623157e308) || !(a1 <= 1.7976931348623157e308);
                    cmax[i][0] = fmax(cmax[i][0], a0);
                    cmax[i][1] = fmax(cmax[i][1], a1);
                    if (CSUM) {
                        csum[CSUM ? i : 0][0] += st.x;
                        csum[CSUM ? i : 0][1] += st.y;
                    }
                }
            }
        }
        HEAD_STAMP(6);
        // ---- dW += A^T delta_L ----
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int mt = warp + HEAD_WARPS * i;        // hidden tile
            if (8 * mt >= H) break;
            const T *ap = As + t * ldA + 8 * mt + g;
            const int ldA4 = 4 * ldA;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const double *bp = Zs + t * ldD + 8 * n + g;
#pragma unroll
                for (int k = 0; k < R / 4; ++k) dmma_8x8x4(dw[i][n][0], dw[i][n][1], (double)ap[k * ldA4], bp[4 * k * ldD]);
            }
        }
        HEAD_STAMP(7);
        __syncthreads();                       // Zs / this A buffer are free again
    }
    cp_async_wait<0>();
    if (bad) *p.flag = 1;
    if (backward) {
        cbad = __any_sync(0xffffffffu, cbad);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int mt = warp + HEAD_WARPS * i;
            if (8 * mt >= H) break;
            double *dst = p.dw_partial + (size_t)blockIdx.x * H * C + (size_t)(8 * mt + g) * C;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int c = 8 * n + 2 * t;
                if (c < C) dst[c] = dw[i][n][0];
                if (c + 1 < C) dst[c + 1] = dw[i][n][1];
            }
            if (CSUM && p.colsum_partial) {               // fixed-order fold of the 8 row groups, one partial per CTA
                double s0 = csum[CSUM ? i : 0][0], s1 = csum[CSUM ? i : 0][1];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (g == 0) {
                    p.colsum_partial[(size_t)blockIdx.x * H + 8 * mt + 2 * t] = s0;
                    p.colsum_partial[(size_t)blockIdx.x * H + 8 * mt + 2 * t + 1] = s1;
                }
            }
            if (DA && p.colmax) {                         // columns 8 mt + 2t, +1: fold the 8 row groups, then one atomic
                double m0 = cmax[i][0], m1 = cmax[i][1];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                    m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
                }
                if (g == 0) {
                    unsigned long long *cm = reinterpret_cast<unsigned long long *>(p.colmax) + 8 * mt + 2 * t;
                    atomicMax(cm, cbad ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(m0));
                    atomicMax(cm + 1, cbad ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(m1));
                }
            }
        }
    }
    if (!DA && backward && (lane & 15) < C)
        atomicMax(reinterpret_cast<unsigned long long *>(p.classmax) + (lane & 15),
                  lbad ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(lmax));
    loss = block_sum(loss, scratch);
    if (threadIdx.x == 0) p.loss_partial[blockIdx.x] = loss;
}

// obj = -(sum/N) for cross entropy, sum/(N*C) for MSE.  out[1] (when `flag` is given) = 1.0 if the row stage met
// log(negative), else 0.0: the domain flag travels in the evaluation vector, so that a batch-sharded evaluation all-reduces it
// with the gradient and every rank raises FloatingPointError (error.py:89-91) at the same evaluation.
__global__ void loss_final_kernel(const double *partial, int nparts, int eid, double n_rows, double n_cols,
                                  double *out, const int *flag) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        out[0] = eid == OPL_ERROR_MSE ? s / (n_rows * n_cols) : -(s / n_rows);
        if (flag) out[1] = *flag ? 1.0 : 0.0;
    }
}

// The fused head's two folds in ONE launch: the per-CTA dW partials (one warp per output element, lanes over the CTAs,
// fixed order) and -- block 0, afterwards -- the per-CTA loss partials with the objective's scale and the domain flag.
// (partial2 / out2: an optional second family of per-CTA partials -- the bias gradient's column sums when the head's
// delta_prev is delta_0)
__global__ void __launch_bounds__(256) head_fold_kernel(const double *__restrict__ partial, int splits, int64_t stride,
                                                        int64_t count, double *__restrict__ out,
                                                        const double *__restrict__ loss_partial, int eid, double n_rows,
                                                        double n_cols, double *__restrict__ obj_out, const int *flag,
                                                        const double *__restrict__ partial2 = nullptr, int64_t stride2 = 0,
                                                        int64_t count2 = 0, double *__restrict__ out2 = nullptr) {
    __shared__ double scratch[33];
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < count + count2; i += nwarps) {
        const bool second = i >= count;
        const double *src = second ? partial2 + (i - count) : partial + i;
        const int64_t st = second ? stride2 : stride;
        double s = 0.0;
        for (int k = lane; k < splits; k += 32) s += src[(int64_t)k * st];
        s = warp_sum(s);
        if (lane == 0) (second ? out2 + (i - count) : out + i)[0] = s;
    }
    if (blockIdx.x == 0) {
        double s = 0.0;
        for (int i = threadIdx.x; i < splits; i += blockDim.x) s += loss_partial[i];
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) {
            obj_out[0] = eid == OPL_ERROR_MSE ? s / (n_rows * n_cols) : -(s / n_rows);
            obj_out[1] = *flag ? 1.0 : 0.0;
        }
    }
}

// column sums of a rows x cols matrix, stage 1: partial[blockIdx.y][col]
__global__ void colsum_partial_kernel(const double *__restrict__ D, int64_t rows, int cols, int64_t rows_per_block,
                                      double *__restrict__ partial) {
    __shared__ double tile[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = blockIdx.y * rows_per_block;
    const int64_t r1 = min(rows, r0 + rows_per_block);
    double s = 0.0;
    if (col < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += D[r * cols + col];
    tile[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < cols) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += tile[k][threadIdx.x];
        partial[(int64_t)blockIdx.y * cols + col] = tot;
    }
}

// rows of a (rows x cols) matrix times a per-row 0/1 factor: the input dropout mask folded into W0 / dW0
__global__ void scale_rows_kernel(double *M, int64_t rows, int64_t cols, const double *__restrict__ factor) {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        M[i] *= factor[i / cols];
}

// trial point w = x + fl(alpha*dir) (or a copy), and per-evaluation flag reset
__global__ void trial_point_kernel(int64_t n, const double *__restrict__ x, double alpha,
                                   const double *__restrict__ dir, double *__restrict__ w, int *flag) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = dir ? __dadd_rn(x[i], __dmul_rn(alpha, dir[i])) : x[i];
}

// probe scalars: partial[2b] = sum g*d, partial[2b+1] = sum g*g
__global__ void probe_partial_kernel(int64_t n, const double *__restrict__ g, const double *__restrict__ d,
                                     double *__restrict__ partial) {
    __shared__ double scratch[33];
    double gd = 0.0, gg = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double gi = g[i];
        if (d) gd += gi * d[i];
        gg += gi * gi;
    }
    gd = block_sum(gd, scratch);
    gg = block_sum(gg, scratch);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = gd;
        partial[2 * blockIdx.x + 1] = gg;
    }
}
__global__ void probe_final_kernel(int nparts, const double *__restrict__ partial, const double *obj,
                                   double *__restrict__ out) {   // obj[0] = objective, obj[1] = domain flag (summed over ranks)
    __shared__ double scratch[33];
    double gd = 0.0, gg = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        gd += partial[2 * i];
        gg += partial[2 * i + 1];
    }
    gd = block_sum(gd, scratch);
    gg = block_sum(gg, scratch);
    if (threadIdx.x == 0) {
        out[0] = obj[0];
        out[1] = gd;
        out[2] = gg;
        out[3] = obj[1];
    }
}

// per-row softmax Jacobian from the OUTPUT y (calculate.py:165-174): J[j][k] = -y_j y_k, J[j][j] = y_j (1 - y_j)
__global__ void softmax_jacobian_kernel(const double *__restrict__ y, int64_t rows, int C, double *__restrict__ jac) {
    const int64_t total = rows * (int64_t)C * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % C);
        const int j = (int)((i / C) % C);
        const int64_t row = i / ((int64_t)C * C);
        const double yj = y[row * C + j], yk = y[row * C + k];
        jac[i] = j == k ? yj * (1. - yj) : -yj * yk;
    }
}

}  // namespace opl

using namespace opl;

struct FastState;   // fast-mode (TF32) buffers, mlp_fast.inl
struct Int8State;   // parity-mode GEMMs on the INT8 tensor cores, mlp_int8.inl

// ---------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------
struct opl_mlp {
    FastState *fast = nullptr;
    Int8State *int8 = nullptr;
    int gemm_engine = OPL_ENGINE_INT8;   // which kernels run the large FP64 layer GEMMs
    bool fused_head = true;              // narrow last layer + error + its back-propagation in one streaming kernel
    int layers = 0;                    // weight matrices
    std::vector<int64_t> width;        // layers + 1
    std::vector<int> transfer;
    std::vector<double> tparam;
    std::vector<int64_t> w_off;        // offset of W_l in the flat vector
    int error_id = 0;
    int precision = 0;
    int64_t n = 0, max_rows = 0, rows = 0, norm_rows = 0;
    cudaStream_t stream = nullptr;
    // the fused head's folds (dW_{L-1}, loss) run beside the delta digit kernel on a stream of their own: nothing reads their
    // results before the evaluation ends, and their INPUT (the head's partials in ws->partials) is safe until the dW_0 GEMM
    // reuses that buffer for its split-K partials -- int8_dw_layer rejoins the streams right before it (OPL_SIDE_FOLD=0: in line)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_head = nullptr, ev_fold = nullptr;
    bool fold_pending = false, side_fold = true;

    double *x_own = nullptr, *y_own = nullptr;   // owned copies (set_data_host)
    const double *x = nullptr, *y = nullptr;     // resident data (owned or borrowed)
    double *x_scratch = nullptr;                 // activate() on fresh patterns
    std::vector<double *> act;      // act[l], l = 1..layers (act[layers] holds Z_L)
    std::vector<double *> deriv;    // D_l or null
    std::vector<double *> delta;    // delta_l (aliases deriv[l] when that exists)
    double *out_act = nullptr;      // A_L
    double *w_trial = nullptr;      // n + 2
    double *host_grad = nullptr;    // device staging for *_host calls, n + 2
    double *partials = nullptr;
    int64_t partials_len = 0;
    double *loss_partial = nullptr;
    int loss_blocks = 0;
    double *probe_partial = nullptr;
    int probe_blocks = 0;
    double *scal = nullptr;         // device: 4 doubles
    int *flag = nullptr;
    double *pinned = nullptr;       // host pinned: 4 doubles + flag
    double *pinned_vec = nullptr;   // host pinned staging for the n-vector of the *_host calls (allocated on first use)
    double *masks = nullptr;        // DropoutMLP: [d0 ; d1 ; ... ; d_{L-1}] 0/1 vectors, or null
    std::vector<int64_t> mask_off;  // offset of layer l's mask (l = 0 input, l >= 1 hidden outputs)
    int64_t launches = 0;
    // one-launch path for small networks (mlp_small.inl)
    bool direct_delta_digits = true;     // two-layer INT8 schedule: delta_0 only as digits (OPL_DIRECT_DELTA=0: materialise it)
    bool late_transfer = true;           // two-layer INT8 schedule, softplus hidden layer: the forward GEMM stores z, the fused
                                         // head applies the transfer (OPL_LATE_TRANSFER=0: in the GEMM epilogue)
    bool small_path = true;
    bool small_configured = false;
    size_t small_smem = 0;
    double *mapped_host = nullptr;  // page-locked, mapped into the device: obj, g.dir, g.g, flag, sequence number
    double *mapped_dev = nullptr;   // the same memory through its device address
    double mapped_seq = 0.0;
    unsigned int *ticket = nullptr;
    // optional per-launch timing (CUDA events on the launch stream): kind = stage*100 + layer
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_kind;
    size_t ev_used = 0;
};

enum { STAGE_TRIAL = 0, STAGE_FWD = 1, STAGE_ROWS = 2, STAGE_DW = 3, STAGE_DW_REDUCE = 4, STAGE_DA = 5,
       STAGE_BIAS = 6, STAGE_LOSS = 7, STAGE_DIGITS = 8, STAGE_END = 9, STAGE_HEAD = 10, STAGE_SMALL = 11 };

namespace {

// record "a launch of `kind` starts here" (no-op unless profiling)
void mark(opl_mlp *ws, int kind) {
    if (!ws->profiling) return;
    if (ws->ev_used == ws->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ws->ev_pool.push_back(e);
        ws->ev_kind.push_back(0);
    }
    ws->ev_kind[ws->ev_used] = kind;
    cudaEventRecord(ws->ev_pool[ws->ev_used], ws->stream);
    ++ws->ev_used;
}

int row_group(int c) {
    int g = 1;
    while (g < c && g < 32) g <<= 1;
    return g;
}

constexpr int kThreadRowMaxC = 16;

int row_grid(int64_t rows, int group) {
    const int rows_per_block = 256 / group;
    int64_t blocks = ceil_div(rows, rows_per_block);
    const int64_t cap = 8 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// pick the row kernel: thread-per-row for narrow outputs unless it works in place (hidden softmax)
// returns the grid size used (== number of loss partials written)
template <typename T>
int launch_rows(const RowArgsT<T> &r, int max_grid, cudaStream_t stream) {
    if (r.C <= kThreadRowMaxC && r.A != r.Z) {
        int64_t grid = ceil_div(r.rows, 128);
        if (grid > max_grid) grid = max_grid;
        if (grid < 1) grid = 1;
        output_rows_thread_kernel<kThreadRowMaxC, T><<<(unsigned)grid, 128, 0, stream>>>(r);
        return (int)grid;
    }
    int grid = row_grid(r.rows, r.group);
    if (grid > max_grid) grid = max_grid;
    output_rows_kernel<T><<<grid, 256, 0, stream>>>(r);
    return grid;
}

void fast_free(FastState *f);
void int8_free(Int8State *s);

int free_all(opl_mlp *ws) {
    if (ws->side) {
        cudaStreamSynchronize(ws->side);
        cudaEventDestroy(ws->ev_head);
        cudaEventDestroy(ws->ev_fold);
        cudaStreamDestroy(ws->side);
        ws->side = nullptr;
        ws->ev_head = ws->ev_fold = nullptr;
        ws->fold_pending = false;
    }
    fast_free(ws->fast);
    ws->fast = nullptr;
    int8_free(ws->int8);
    ws->int8 = nullptr;
    cudaFree(ws->x_own);
    cudaFree(ws->y_own);
    cudaFree(ws->x_scratch);
    for (size_t l = 0; l < ws->act.size(); ++l) cudaFree(ws->act[l]);
    for (size_t l = 0; l < ws->deriv.size(); ++l) cudaFree(ws->deriv[l]);
    for (size_t l = 0; l < ws->delta.size(); ++l)
        if (l >= ws->deriv.size() || ws->delta[l] != ws->deriv[l]) cudaFree(ws->delta[l]);
    cudaFree(ws->out_act);
    cudaFree(ws->w_trial);
    cudaFree(ws->host_grad);
    cudaFree(ws->partials);
    cudaFree(ws->loss_partial);
    cudaFree(ws->probe_partial);
    cudaFree(ws->scal);
    cudaFree(ws->flag);
    cudaFree(ws->masks);
    cudaFreeHost(ws->pinned);
    cudaFreeHost(ws->pinned_vec);
    cudaFreeHost(ws->mapped_host);
    cudaFree(ws->ticket);
    for (cudaEvent_t e : ws->ev_pool) cudaEventDestroy(e);
    return OPL_OK;
}

int epilogue_for(int transfer) {
    switch (transfer) {
        case OPL_TRANSFER_TANH: return EPI_TANH;
        case OPL_TRANSFER_SOFTPLUS: return EPI_SOFTPLUS;
        case OPL_TRANSFER_GAUSSIAN: return EPI_GAUSSIAN;
        default: return EPI_STORE;
    }
}

}  // namespace

#include "mlp_int8.inl"

namespace {

// forward through all layers; leaves Z_L in act[layers]
int forward(opl_mlp *ws, const double *w, const double *x, int64_t rows, bool keep_deriv, bool masked = false,
            int layer_end = -1) {
    const double *in = x;
    if (layer_end < 0) layer_end = ws->layers;
    for (int l = 0; l < layer_end; ++l) {
        GemmArgs a = {};
        a.A = in;
        a.lda = ws->width[l];
        a.B = w + ws->w_off[l];
        a.ldb = ws->width[l + 1];
        a.C = ws->act[l + 1];
        a.ldc = ws->width[l + 1];
        a.M = (int)rows;
        a.N = (int)ws->width[l + 1];
        a.K = (int)ws->width[l];
        a.k_per_split = (int)(ceil_div(a.K, GEMM_BK) * GEMM_BK);
        a.bias = l == 0 ? w : nullptr;   // only the first layer has a bias (mlp.py:150-151)
        a.ldaux = ws->width[l + 1];
        a.tparam = ws->tparam[l];
        const bool last = l == ws->layers - 1;
        a.epi = last ? EPI_STORE : epilogue_for(ws->transfer[l]);
        a.aux_out = (!last && keep_deriv) ? ws->deriv[l] : nullptr;
        a.colmask = (!last && ws->masks && masked) ? ws->masks + ws->mask_off[l + 1] : nullptr;
        int rc;
        if (int8_fwd_ok(ws, l, in, rows)) {
            rc = int8_forward_layer(ws, l, w, in, rows, keep_deriv, masked);
        } else {
            mark(ws, STAGE_FWD * 100 + l);
            rc = launch_gemm_f64(GEMM_NN, a, 1, ws->stream, &ws->launches);
        }
        if (rc != OPL_OK) return rc;
        if (!last && ws->transfer[l] == OPL_TRANSFER_SOFTMAX) {
            mark(ws, STAGE_ROWS * 100 + l);   // hidden softmax: row kernel in place
            RowArgs r = {};
            r.Z = ws->act[l + 1];
            r.A = ws->act[l + 1];
            r.rows = rows;
            r.C = a.N;
            r.tid = OPL_TRANSFER_SOFTMAX;
            r.flag = ws->flag;
            r.ld = r.C;
            r.group = row_group(r.C);
            output_rows_kernel<double><<<row_grid(rows, r.group), 256, 0, ws->stream>>>(r);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
        }
        in = ws->act[l + 1];
    }
    return OPL_OK;
}

int output_stage(opl_mlp *ws, const double *y, int64_t rows, bool want_act, bool want_delta, double *obj_out) {
    const int L = ws->layers;
    RowArgs r = {};
    r.Z = ws->act[L];
    r.Y = y;
    r.A = want_act ? ws->out_act : nullptr;
    r.D = want_delta ? ws->delta[L - 1] : nullptr;
    r.rows = rows;
    r.C = (int)ws->width[L];
    r.ld = r.C;
    r.tid = ws->transfer[L - 1];
    r.tparam = ws->tparam[L - 1];
    r.eid = ws->error_id;
    r.ce_div = -(double)ws->norm_rows;
    r.mse_mul = 2.0 / ((double)ws->norm_rows * (double)r.C);
    r.flag = ws->flag;
    r.group = row_group(r.C);
    r.loss_partial = y ? ws->loss_partial : nullptr;
    mark(ws, STAGE_ROWS * 100 + L);
    const int grid = launch_rows(r, ws->loss_blocks, ws->stream);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    if (y && obj_out) {
        mark(ws, STAGE_LOSS * 100);
        loss_final_kernel<<<1, 256, 0, ws->stream>>>(ws->loss_partial, grid, ws->error_id, (double)ws->norm_rows,
                                                      (double)r.C, obj_out, ws->flag);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
    }
    return OPL_OK;
}

int backward(opl_mlp *ws, const double *w, const double *x, int64_t rows, double *grad, int layer_begin = -1) {
    if (layer_begin < 0) layer_begin = ws->layers - 1;
    for (int l = layer_begin; l >= 0; --l) {
        const int din = (int)ws->width[l], dout = (int)ws->width[l + 1];
        // dW_l = A_l^T delta_l  (mlp.py:269-273), reduction over the patterns
        if (int8_dw_ok(ws, l, l == 0 ? x : ws->act[l], rows)) {
            int rc = int8_dw_layer(ws, l, l == 0 ? x : ws->act[l], rows, grad);
            if (rc != OPL_OK) return rc;
        } else {
            SplitPlan plan = plan_split_k(din, dout, (int)rows);
            if ((int64_t)plan.splits * ((int64_t)din * dout + dout) > ws->partials_len)
                return fail(OPL_E_STATE, "split-K scratch too small (layer %d)", l);
            GemmArgs a = {};
            a.A = l == 0 ? x : ws->act[l];
            a.lda = din;
            a.B = ws->delta[l];
            a.ldb = dout;
            a.M = din;
            a.N = dout;
            a.K = (int)rows;
            a.ldc = dout;
            a.k_per_split = plan.k_per_split;
            a.epi = EPI_STORE;
            // layer 0: the first M-tile's CTAs also sum the columns of delta_0 (bias gradient).  The
            // flat layout puts bias right before W_0, so [db ; dW_0] is ONE contiguous reduce.
            const int64_t lead = l == 0 ? dout : 0;
            mark(ws, STAGE_DW * 100 + l);
            if (plan.splits == 1) {
                a.C = grad + ws->w_off[l];
                a.colsum_out = l == 0 ? grad : nullptr;
                int rc = launch_gemm_f64(GEMM_TN, a, 1, ws->stream, &ws->launches);
                if (rc != OPL_OK) return rc;
            } else {
                a.split_stride = lead + (int64_t)din * dout;
                a.C = ws->partials + lead;
                a.colsum_out = l == 0 ? ws->partials : nullptr;
                int rc = launch_gemm_f64(GEMM_TN, a, plan.splits, ws->stream, &ws->launches);
                if (rc != OPL_OK) return rc;
                const int64_t count = lead + (int64_t)din * dout;
                mark(ws, STAGE_DW_REDUCE * 100 + l);
                rc = launch_reduce_partials(ws->partials, plan.splits, count, count, grad + ws->w_off[l] - lead,
                                            ws->stream, &ws->launches);
                if (rc != OPL_OK) return rc;
            }
        }
        if (l == 0) break;
        // delta_{l-1} = (delta_l W_l^T) o f'_{l-1}   (mlp.py:254-260)
        if (dout <= 16 && ws->transfer[l - 1] != OPL_TRANSFER_SOFTMAX &&
            sizeof(double) * ((size_t)din * (dout | 1) + (size_t)64 * dout) <= 48 * 1024) {
            // narrow next layer: K = dout <= 16 makes this a streaming pass, not a GEMM
            const int t = ws->transfer[l - 1];
            const int mode = (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN) ? 1 : (t == OPL_TRANSFER_TANH ? 2 : 0);
            const double *aux = mode == 1 ? ws->deriv[l - 1] : (mode == 2 ? ws->act[l] : nullptr);
            const int rpb = 16;
            const size_t smem = sizeof(double) * ((size_t)din * (dout | 1) + (size_t)rpb * dout);
            const int64_t grid = std::min<int64_t>(ceil_div(rows, rpb), 6 * (int64_t)sm_count());
            double *dmax = int8_arm_delta_max(ws, l - 1, rows);
            mark(ws, STAGE_DA * 100 + l);
            narrow_backprop_kernel<16><<<(unsigned)grid, 256, smem, ws->stream>>>(
                ws->delta[l], w + ws->w_off[l], aux, ws->delta[l - 1], rows, din, dout, rpb, mode, dmax);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
            continue;
        }
        if (int8_da_ok(ws, l, rows)) {
            int rc = int8_da_layer(ws, l, w, rows);
            if (rc != OPL_OK) return rc;
            if (ws->transfer[l - 1] == OPL_TRANSFER_SOFTMAX) {
                mark(ws, STAGE_ROWS * 100 + l - 1);
                const int group = row_group(din);
                softmax_backward_rows_kernel<double><<<row_grid(rows, group), 256, 0, ws->stream>>>(
                    ws->delta[l - 1], ws->act[l], rows, din, din, group);
                OPL_LAUNCH_CHECK();
                ++ws->launches;
            }
            continue;
        }
        GemmArgs a = {};
        a.A = ws->delta[l];
        a.lda = dout;
        a.B = w + ws->w_off[l];   // din x dout row-major == N x K for the NT product
        a.ldb = dout;
        a.C = ws->delta[l - 1];
        a.ldc = din;
        a.M = (int)rows;
        a.N = din;
        a.K = dout;
        a.k_per_split = (int)(ceil_div(a.K, GEMM_BK) * GEMM_BK);
        a.ldaux = din;
        switch (ws->transfer[l - 1]) {
            case OPL_TRANSFER_SOFTPLUS:
            case OPL_TRANSFER_GAUSSIAN:
                a.epi = EPI_MUL_AUX;
                a.aux_in = ws->deriv[l - 1];
                break;
            case OPL_TRANSFER_TANH:
                a.epi = EPI_MUL_DTANH;
                a.aux_in = ws->act[l];
                break;
            default:
                a.epi = EPI_STORE;
        }
        mark(ws, STAGE_DA * 100 + l);
        int rc = launch_gemm_f64(GEMM_NT, a, 1, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        if (ws->transfer[l - 1] == OPL_TRANSFER_SOFTMAX) {
            mark(ws, STAGE_ROWS * 100 + l - 1);
            const int group = row_group(din);
            softmax_backward_rows_kernel<double><<<row_grid(rows, group), 256, 0, ws->stream>>>(
                ws->delta[l - 1], ws->act[l], rows, din, din, group);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
        }
    }
    return OPL_OK;
}

// Can the fused narrow head replace forward(last layer) + output stage + the head's back-propagation?
size_t head_smem(int H, int R, size_t elem = sizeof(double)) {
    return sizeof(double) * ((size_t)H * 20 + (size_t)R * 20) + elem * 2 * (size_t)R * (H + (elem == 8 ? 4 : 8));
}
int head_rows_per_step(const opl_mlp *ws) {
    const int L = ws->layers;
    if (L < 2 || ws->precision != OPL_PRECISION_F64 || !ws->fused_head) return 0;
    const int64_t H = ws->width[L - 1], C = ws->width[L];
    if (C > 16 || H > 256 || (H & 7) || H < 32) return 0;
    if (ws->transfer[L - 2] == OPL_TRANSFER_SOFTMAX) return 0;
    static int r16 = -1;
    if (r16 < 0) {
        const char *env = getenv("OPL_HEAD_R");
        r16 = (env && atoi(env) == 32) ? 0 : 1;
    }
    return r16 ? 16 : 32;
}

// Can the head leave delta_{L-2} to the digit kernel of the weight-gradient GEMM (ozaki_slice_delta_prev)?  Two-layer networks
// whose first layer's weight gradient runs on the INT8 engine: delta_0 is then needed ONLY as that GEMM's digit operand (the
// bias gradient is the ones row of the other operand), and the slope of the hidden transfer is bounded by one.
bool head_skips_delta_prev(const opl_mlp *ws, int64_t rows) {
    if (ws->layers != 2 || ws->masks || !ws->direct_delta_digits) return false;
    const int t = ws->transfer[0];
    if (t != OPL_TRANSFER_SOFTPLUS && t != OPL_TRANSFER_TANH && t != OPL_TRANSFER_LINEAR) return false;
    return int8_dw_ok(ws, 0, ws->x, rows) && rows == ws->rows;
}

// fused head: objective (and, when want_grad, dW_{L-1} and delta_{L-2}); the caller continues back-propagation at L-2
int fused_head(opl_mlp *ws, const double *w, const double *y, int64_t rows, bool want_grad, double *grad, double *obj_out,
               bool late_transfer = false) {
    const int L = ws->layers, R = head_rows_per_step(ws);
    const int H = (int)ws->width[L - 1], C = (int)ws->width[L];
    const size_t smem = head_smem(H, R);
    int grid = (int)std::min<int64_t>(ceil_div(rows, R), (int64_t)sm_count() * (R == 16 ? 2 : 1));
    grid = std::min(grid, ws->loss_blocks);
    if (want_grad) grid = (int)std::min<int64_t>(grid, ws->partials_len / ((int64_t)H * C));
    if (grid < 1) return fail(OPL_E_STATE, "fused head: scratch too small");
    HeadArgs a = {};
    a.A = ws->act[L - 1];
    a.W = w + ws->w_off[L - 1];
    a.Y = y;
    a.rows = rows;
    a.ld = H;
    a.H = H;
    a.C = C;
    a.R = R;
    a.tid = ws->transfer[L - 1];
    a.tparam = ws->tparam[L - 1];
    a.eid = ws->error_id;
    a.ce_div = -(double)ws->norm_rows;
    a.ce_rdiv = 1.0 / a.ce_div;
    a.mse_mul = 2.0 / ((double)ws->norm_rows * (double)C);
    a.flag = ws->flag;
    a.loss_partial = ws->loss_partial;
    const bool no_da = want_grad && R == 16 && head_skips_delta_prev(ws, rows);
    if (late_transfer && (R != 16 || (want_grad && !no_da))) return fail(OPL_E_STATE, "fused head: late transfer without its schedule");
    if (late_transfer && want_grad) a.deriv_out = ws->deriv[L - 2];
    if (want_grad) {
        const int t = ws->transfer[L - 2];
        a.mode = (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN) ? 1 : (t == OPL_TRANSFER_TANH ? 2 : 0);
        a.aux = a.mode == 1 ? ws->deriv[L - 2] : (a.mode == 2 ? ws->act[L - 1] : nullptr);
        a.dw_partial = ws->partials;
        if (no_da) {
            a.delta_out = ws->delta[L - 1];          // rows x C: delta_L
            a.classmax = ws->int8->classmax;
            if (!ws->int8->accumulators_cleared) OPL_CUDA(cudaMemsetAsync(a.classmax, 0, sizeof(double) * 16, ws->stream));
        } else {
            a.delta_prev = ws->delta[L - 2];
            a.colmax = int8_arm_delta_max(ws, L - 2, rows);
        }
    }
    static bool configured_dev[64] = {};     // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_kernel<double, 2, 32, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem(256, 32)));
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_kernel<double, 4, 16, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem(256, 16)));
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_kernel<double, 4, 16, 8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem(256, 16)));
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_kernel<double, 4, 16, 8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem(256, 16)));
        configured = true;
    }
    static long long *head_trace = nullptr;
    const char *trace_path = getenv("OPL_HEAD_TRACE");
#ifndef OPL_HEAD_TRACE_BUILD
    if (trace_path) {
        static bool told = false;
        if (!told) fprintf(stderr, "opl_b200: OPL_HEAD_TRACE needs a library built with -DOPL_HEAD_TRACE_BUILD; ignored\n");
        told = true;
        trace_path = nullptr;
    }
#endif
    if (trace_path && !head_trace) {
        OPL_CUDA(cudaMalloc(&head_trace, sizeof(long long) * 128));
        OPL_CUDA(cudaMemset(head_trace, 0, sizeof(long long) * 128));
    }
    a.trace = trace_path ? head_trace : nullptr;
    mark(ws, STAGE_HEAD * 100 + L - 1);
    if (R == 32)
        narrow_head_kernel<double, 2, 32, 16, false><<<grid, 512, smem, ws->stream>>>(a);
    else if (late_transfer)
        narrow_head_kernel<double, 4, 16, 8, false, false, true><<<grid, 256, smem, ws->stream>>>(a);
    else if (no_da)
        narrow_head_kernel<double, 4, 16, 8, false, false><<<grid, 256, smem, ws->stream>>>(a);
    else
        narrow_head_kernel<double, 4, 16, 8, false><<<grid, 256, smem, ws->stream>>>(a);
    OPL_LAUNCH_CHECK();
    if (trace_path) {     // tools only: per-step phase stamps of CTA 0, thread 0 (synchronises)
        long long host[128];
        OPL_CUDA(cudaStreamSynchronize(ws->stream));
        OPL_CUDA(cudaMemcpy(host, head_trace, sizeof(host), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trace_path, "w")) {
            fprintf(f, "# narrow_head_kernel, CTA 0 thread 0, cycles since the first stamp: step_begin tile_landed logits_done sync row_stage_done sync dA_done dW_done\n");
            for (int i = 0; i < 16 && host[8 * i + 1]; ++i) {
                for (int k = 0; k < 8; ++k) fprintf(f, "%lld ", host[8 * i + k] - host[0]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    ++ws->launches;
    if (want_grad && grid >= 16) {
        // one launch folds the per-CTA dW partials AND the loss partials (objective + domain flag)
        const int64_t count = (int64_t)H * C;
        const int64_t fgrid = std::min<int64_t>(ceil_div(count * 32, 256), 4 * (int64_t)sm_count());
        cudaStream_t fold_stream = ws->stream;
        if (no_da && ws->side_fold && !ws->profiling) {
            // two-layer schedule: the digit kernel and the dW_0 GEMM follow and need nothing from the folds
            if (!ws->side) {
                OPL_CUDA(cudaStreamCreateWithFlags(&ws->side, cudaStreamNonBlocking));
                OPL_CUDA(cudaEventCreateWithFlags(&ws->ev_head, cudaEventDisableTiming));
                OPL_CUDA(cudaEventCreateWithFlags(&ws->ev_fold, cudaEventDisableTiming));
            }
            OPL_CUDA(cudaEventRecord(ws->ev_head, ws->stream));
            OPL_CUDA(cudaStreamWaitEvent(ws->side, ws->ev_head, 0));
            fold_stream = ws->side;
        } else {
            mark(ws, STAGE_DW_REDUCE * 100 + L - 1);
        }
        head_fold_kernel<<<(unsigned)fgrid, 256, 0, fold_stream>>>(ws->partials, grid, count, count, grad + ws->w_off[L - 1],
                                                                ws->loss_partial, ws->error_id, (double)ws->norm_rows,
                                                                (double)C, obj_out, ws->flag);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
        if (fold_stream != ws->stream) {
            OPL_CUDA(cudaEventRecord(ws->ev_fold, fold_stream));
            ws->fold_pending = true;             // evaluate() joins before it returns
        }
        return OPL_OK;
    }
    mark(ws, STAGE_LOSS * 100);
    loss_final_kernel<<<1, 256, 0, ws->stream>>>(ws->loss_partial, grid, ws->error_id, (double)ws->norm_rows, (double)C, obj_out, ws->flag);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    if (want_grad) {
        mark(ws, STAGE_DW_REDUCE * 100 + L - 1);
        const int64_t count = (int64_t)H * C;
        int rc = launch_reduce_partials(ws->partials, grid, count, count, grad + ws->w_off[L - 1], ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
    }
    return OPL_OK;
}

}  // namespace

#include "mlp_fast.inl"
#include "mlp_small.inl"

namespace {

int evaluate(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev, bool want_grad,
             double *grad_out_dev) {
    if (!ws->x || ws->rows <= 0) return fail(OPL_E_STATE, "no training data resident: call opl_mlp_set_data_* first");
    const int64_t n = ws->n;
    {   // small network: the whole evaluation is one launch
        const SmallPlan pl = small_plan(ws, ws->rows);
        if (pl.ok) return small_evaluate(ws, pl, x_dev, alpha, dir_dev, want_grad, grad_out_dev, nullptr);
    }
    int tgrid = (int)std::min<int64_t>(ceil_div(n, 256), 4 * (int64_t)sm_count());
    mark(ws, STAGE_TRIAL * 100);
    int rc;
    if (ws->precision == OPL_PRECISION_TF32) {
        rc = fast_trial_convert(ws, x_dev, alpha, dir_dev);      // trial point + padded TF32 copy of the weights: one launch
        if (rc == OPL_OK) rc = fast_evaluate(ws, want_grad, grad_out_dev);
        if (ws->fold_pending) {                  // the head's folds rejoin the launch stream
            ws->fold_pending = false;
            OPL_CUDA(cudaStreamWaitEvent(ws->stream, ws->ev_fold, 0));
        }
        mark(ws, STAGE_END * 100);
        return rc;
    }
    bool fused_trial = false;
    if (ws->int8) ws->int8->w0_digits_ready = ws->int8->accumulators_cleared = false;
    if (!ws->masks && int8_fwd_ok(ws, 0, ws->x, ws->rows)) {
        // INT8 forward layer 0: the trial point and the digits of W_0^T in one cooperative launch
        Int8State *s8 = ws->int8;
        // (it also clears the class maxima and the delta scales the two-layer schedule accumulates into by atomicMax)
        const bool two_layer = want_grad && head_rows_per_step(ws) == 16 && head_skips_delta_prev(ws, ws->rows);
        rc = ozaki_trial_w0_digits(x_dev, dir_dev, alpha, ws->w_trial, n, ws->w_off[0], ws->width[0], ws->width[1], s8->w_dig,
                                   s8->w_scale, s8->colmax, s8->colmax_len, ozaki_pad(ws->width[1]), ws->flag,
                                   two_layer ? s8->classmax : nullptr, 16, two_layer ? s8->d_scale : nullptr, (int)ws->width[1],
                                   ws->stream, &ws->launches, &fused_trial);
        if (rc != OPL_OK) return rc;
        s8->w0_digits_ready = fused_trial;
        s8->accumulators_cleared = fused_trial && two_layer;
    }
    if (!fused_trial) {
        trial_point_kernel<<<tgrid, 256, 0, ws->stream>>>(n, x_dev, alpha, dir_dev, ws->w_trial, ws->flag);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
    }
    const int64_t w0_count = ws->width[0] * ws->width[1];
    const int mgrid = (int)std::min<int64_t>(ceil_div(w0_count, 256), 4 * (int64_t)sm_count());
    if (ws->masks) {   // (X o m) W0 == X (diag(m) W0) exactly for a 0/1 mask
        scale_rows_kernel<<<mgrid, 256, 0, ws->stream>>>(ws->w_trial + ws->w_off[0], ws->width[0], ws->width[1], ws->masks);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
    }
    if (head_rows_per_step(ws) > 0) {
        // two-layer INT8 network with a softplus hidden layer: the forward GEMM stores the pre-activations and the head applies
        // the transfer (see narrow_head_kernel, ZIN)
        const bool late = ws->late_transfer && ws->layers == 2 && !ws->masks && ws->transfer[0] == OPL_TRANSFER_SOFTPLUS &&
                          head_rows_per_step(ws) == 16 && int8_fwd_ok(ws, 0, ws->x, ws->rows) &&
                          (!want_grad || head_skips_delta_prev(ws, ws->rows));
        if (ws->int8) ws->int8->store_z0 = late;
        rc = forward(ws, ws->w_trial, ws->x, ws->rows, want_grad, true, ws->layers - 1);
        if (ws->int8) ws->int8->store_z0 = false;
        if (rc != OPL_OK) return rc;
        rc = fused_head(ws, ws->w_trial, ws->y, ws->rows, want_grad, grad_out_dev, grad_out_dev + n, late);
        if (rc != OPL_OK) return rc;
        if (want_grad && head_rows_per_step(ws) == 16 && head_skips_delta_prev(ws, ws->rows)) {
            // two-layer network: delta_0 exists only as the digit operand of the dW_0 GEMM, cut from (delta_L, W_1, slope)
            const int t = ws->transfer[0];
            DeltaOnTheFly otf;
            otf.deltaL = ws->delta[1];
            otf.C = (int)ws->width[2];
            otf.W = ws->w_trial + ws->w_off[1];
            otf.mode = t == OPL_TRANSFER_SOFTPLUS ? 1 : (t == OPL_TRANSFER_TANH ? 2 : 0);
            otf.aux = otf.mode == 1 ? ws->deriv[0] : (otf.mode == 2 ? ws->act[1] : nullptr);
            rc = int8_dw_layer(ws, 0, ws->x, ws->rows, grad_out_dev, &otf);
        } else if (want_grad) {
            rc = backward(ws, ws->w_trial, ws->x, ws->rows, grad_out_dev, ws->layers - 2);
        }
    } else {
        rc = forward(ws, ws->w_trial, ws->x, ws->rows, want_grad, true);
        if (rc != OPL_OK) return rc;
        rc = output_stage(ws, ws->y, ws->rows, false, want_grad, grad_out_dev + n);
        if (rc != OPL_OK) return rc;
        if (want_grad) rc = backward(ws, ws->w_trial, ws->x, ws->rows, grad_out_dev);
    }
    if (rc == OPL_OK && want_grad && ws->masks) {   // dW0 = (X o m)^T delta_0 = diag(m) X^T delta_0
        scale_rows_kernel<<<mgrid, 256, 0, ws->stream>>>(grad_out_dev + ws->w_off[0], ws->width[0], ws->width[1], ws->masks);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
    }
    if (ws->fold_pending) {                      // the head's folds rejoin the launch stream
        ws->fold_pending = false;
        OPL_CUDA(cudaStreamWaitEvent(ws->stream, ws->ev_fold, 0));
    }
    mark(ws, STAGE_END * 100);
    return rc;
}

int read_scalars(opl_mlp *ws, const double *grad_dev, const double *dir_dev, bool have_grad, double *host_out) {
    const int64_t n = ws->n;
    if (have_grad) {
        int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ws->probe_blocks);
        probe_partial_kernel<<<grid, 256, 0, ws->stream>>>(n, grad_dev, dir_dev, ws->probe_partial);
        probe_final_kernel<<<1, 256, 0, ws->stream>>>(grid, ws->probe_partial, grad_dev + n, ws->scal);
        OPL_LAUNCH_CHECK();
        ws->launches += 2;
        OPL_CUDA(cudaMemcpyAsync(ws->pinned, ws->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    } else {
        OPL_CUDA(cudaMemcpyAsync(ws->pinned, grad_dev + n, sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
        OPL_CUDA(cudaMemcpyAsync(ws->pinned + 3, grad_dev + n + 1, sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    }
    OPL_CUDA(cudaStreamSynchronize(ws->stream));
    host_out[0] = ws->pinned[0];
    host_out[1] = have_grad ? ws->pinned[1] : 0.0;
    host_out[2] = have_grad ? ws->pinned[2] : 0.0;
    // the flag slot is a sum over ranks of 0/1 values when the batch is sharded: every rank sees the same value
    if (ws->pinned[3] != 0.0)
        return fail(OPL_E_DOMAIN, "invalid value encountered in log (negative prediction in cross entropy)");
    return OPL_OK;
}

}  // namespace

extern "C" {

int opl_mlp_create(opl_mlp **out, int32_t n_widths, const int64_t *widths, const int32_t *transfer_ids,
                   const double *transfer_params, int32_t error_id, int64_t max_rows, int32_t precision,
                   void *cuda_stream) {
    if (!out || !widths || !transfer_ids || n_widths < 2) return fail(OPL_E_ARG, "opl_mlp_create: bad arguments");
    if (error_id != OPL_ERROR_MSE && error_id != OPL_ERROR_CROSS_ENTROPY)
        return fail(OPL_E_ARG, "opl_mlp_create: unknown error id %d", error_id);
    if (precision != OPL_PRECISION_F64 && precision != OPL_PRECISION_TF32)
        return fail(OPL_E_ARG, "opl_mlp_create: unknown precision %d", precision);
    if (max_rows < 1) return fail(OPL_E_ARG, "opl_mlp_create: max_rows must be >= 1");
    for (int i = 0; i < n_widths; ++i)
        if (widths[i] < 1 || widths[i] > (1 << 24)) return fail(OPL_E_ARG, "opl_mlp_create: bad width");
    for (int i = 0; i < n_widths - 1; ++i)
        if (transfer_ids[i] < 0 || transfer_ids[i] > OPL_TRANSFER_LOGIT ||
            (transfer_ids[i] == OPL_TRANSFER_LOGIT && i != n_widths - 2))
            return fail(OPL_E_ARG, "opl_mlp_create: unknown transfer id %d (logit is an output transfer)", transfer_ids[i]);
    if (max_rows > 0x7fffffff) return fail(OPL_E_ARG, "opl_mlp_create: max_rows too large");
    if (opl_device_count() < 1) return fail(OPL_E_CUDA, "no CUDA device: opl_b200 has no CPU fallback");

    opl_mlp *ws = new opl_mlp();
    ws->layers = n_widths - 1;
    ws->width.assign(widths, widths + n_widths);
    ws->transfer.assign(transfer_ids, transfer_ids + ws->layers);
    ws->tparam.resize(ws->layers, 1.0);
    if (transfer_params)
        for (int l = 0; l < ws->layers; ++l) ws->tparam[l] = transfer_params[l];
    ws->error_id = error_id;
    ws->precision = precision;
    ws->max_rows = max_rows;
    ws->stream = (cudaStream_t)cuda_stream;
    ws->w_off.resize(ws->layers);
    int64_t at = widths[1];
    for (int l = 0; l < ws->layers; ++l) {
        ws->w_off[l] = at;
        at += widths[l] * widths[l + 1];
    }
    ws->n = at;
    ws->mask_off.resize(ws->layers);
    for (int l = 0, m = 0; l < ws->layers; ++l) {
        ws->mask_off[l] = m;
        m += (int)widths[l];
    }

    const int L = ws->layers;
    ws->act.assign(L + 1, nullptr);
    ws->deriv.assign(L, nullptr);
    ws->delta.assign(L, nullptr);
    int64_t partials = 0;
#define OPL_TRY(expr)                 \
    do {                              \
        cudaError_t _e = (expr);      \
        if (_e != cudaSuccess) {      \
            free_all(ws);             \
            delete ws;                \
            return fail(OPL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
        }                             \
    } while (0)
    for (int l = 0; l < L && precision == OPL_PRECISION_F64; ++l) {
        const size_t bytes = sizeof(double) * (size_t)max_rows * (size_t)widths[l + 1];
        OPL_TRY(cudaMalloc(&ws->act[l + 1], bytes));
        const int t = ws->transfer[l];
        if (l < L - 1 && (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN)) {
            OPL_TRY(cudaMalloc(&ws->deriv[l], bytes));
            ws->delta[l] = ws->deriv[l];   // delta_l overwrites D_l element-wise in the NT epilogue
        } else {
            OPL_TRY(cudaMalloc(&ws->delta[l], bytes));
        }
        SplitPlan plan = plan_split_k((int)widths[l], (int)widths[l + 1], (int)max_rows);
        partials = std::max<int64_t>(partials, (int64_t)(plan.splits + 1) * (widths[l] * widths[l + 1] + widths[l + 1]));
    }
    partials = std::max<int64_t>(partials, (int64_t)(4 * sm_count() + 8) * widths[1]);
    if (L >= 2)      // per-CTA dW (+ bias-gradient) partials of the fused narrow head, up to 2 CTAs per SM
        partials = std::max<int64_t>(partials, 2 * (int64_t)sm_count() * widths[L - 1] * (widths[L] + 1));
    if (ws->n <= 4096)   // per-CTA gradient partials of the one-launch small-network kernel (32 patterns per CTA at least)
        partials = std::max<int64_t>(partials, std::min<int64_t>((int64_t)1 << 20, ceil_div(max_rows, 32) * (ws->n + 2)));
    if (precision == OPL_PRECISION_F64) {
        const char *env = getenv("OPL_F64_ENGINE");
        if (env && std::string(env) == "dmma") {
            ws->gemm_engine = OPL_ENGINE_DMMA;
            ws->fused_head = false;
        }
        int rc = int8_create(ws, &partials);
        if (rc != OPL_OK) {
            free_all(ws);
            delete ws;
            return rc;
        }
    }
    ws->partials_len = partials;
    OPL_TRY(cudaMalloc(&ws->partials, sizeof(double) * (size_t)partials));
    OPL_TRY(cudaMalloc(&ws->out_act, sizeof(double) * (size_t)max_rows * (size_t)widths[L]));
    OPL_TRY(cudaMalloc(&ws->w_trial, sizeof(double) * (size_t)(ws->n + 2)));
    OPL_TRY(cudaMalloc(&ws->host_grad, sizeof(double) * (size_t)(ws->n + 2)));
    ws->loss_blocks = 8 * sm_count();
    OPL_TRY(cudaMalloc(&ws->loss_partial, sizeof(double) * (size_t)ws->loss_blocks));
    ws->probe_blocks = 2 * sm_count();
    OPL_TRY(cudaMalloc(&ws->probe_partial, sizeof(double) * 2 * (size_t)ws->probe_blocks));
    OPL_TRY(cudaMalloc(&ws->scal, sizeof(double) * 4));
    OPL_TRY(cudaMalloc(&ws->flag, sizeof(int)));
    OPL_TRY(cudaMemset(ws->flag, 0, sizeof(int)));
    OPL_TRY(cudaMallocHost(&ws->pinned, sizeof(double) * 8));
    OPL_TRY(cudaHostAlloc(&ws->mapped_host, sizeof(double) * 8, cudaHostAllocMapped));
    memset(ws->mapped_host, 0, sizeof(double) * 8);
    OPL_TRY(cudaHostGetDevicePointer(&ws->mapped_dev, ws->mapped_host, 0));
    OPL_TRY(cudaMalloc(&ws->ticket, sizeof(unsigned int)));
    OPL_TRY(cudaMemset(ws->ticket, 0, sizeof(unsigned int)));
    if (const char *env = getenv("OPL_SMALL_PATH")) ws->small_path = atoi(env) != 0;
    if (const char *env = getenv("OPL_DIRECT_DELTA")) ws->direct_delta_digits = atoi(env) != 0;
    if (const char *env = getenv("OPL_LATE_TRANSFER")) ws->late_transfer = atoi(env) != 0;
    if (const char *env = getenv("OPL_SIDE_FOLD")) ws->side_fold = atoi(env) != 0;
#undef OPL_TRY
    if (precision == OPL_PRECISION_TF32) {
        int rc = fast_create(ws);
        if (rc != OPL_OK) {
            free_all(ws);
            delete ws;
            return rc;
        }
    }
    *out = ws;
    return OPL_OK;
}

int opl_mlp_destroy(opl_mlp *ws) {
    if (!ws) return OPL_OK;
    cudaStreamSynchronize(ws->stream);
    free_all(ws);
    delete ws;
    return OPL_OK;
}

int opl_mlp_set_engine(opl_mlp *ws, int32_t engine) {
    if (!ws || (engine != OPL_ENGINE_DMMA && engine != OPL_ENGINE_INT8)) return fail(OPL_E_ARG, "opl_mlp_set_engine: bad arguments");
    ws->gemm_engine = engine;
    ws->fused_head = engine == OPL_ENGINE_INT8;   // OPL_ENGINE_DMMA = the plain schedule: one kernel per reference operation
    return OPL_OK;
}

int opl_mlp_set_small_path(opl_mlp *ws, int32_t enable) {
    if (!ws) return fail(OPL_E_ARG, "opl_mlp_set_small_path: null");
    ws->small_path = enable != 0;
    return OPL_OK;
}

int opl_mlp_set_stream(opl_mlp *ws, void *cuda_stream) {
    if (!ws) return fail(OPL_E_ARG, "opl_mlp_set_stream: null");
    if (ws->stream != (cudaStream_t)cuda_stream) {
        OPL_CUDA(cudaStreamSynchronize(ws->stream));
        ws->stream = (cudaStream_t)cuda_stream;
    }
    return OPL_OK;
}

int64_t opl_mlp_param_count(const opl_mlp *ws) { return ws ? ws->n : -1; }
int64_t opl_mlp_launch_count(const opl_mlp *ws) { return ws ? ws->launches : -1; }

int opl_mlp_set_masks_host(opl_mlp *ws, const double *masks, int64_t len) {
    if (!ws) return fail(OPL_E_ARG, "opl_mlp_set_masks_host: null");
    if (!masks) {
        OPL_CUDA(cudaStreamSynchronize(ws->stream));
        cudaFree(ws->masks);
        ws->masks = nullptr;
        return OPL_OK;
    }
    if (ws->precision != OPL_PRECISION_F64) return fail(OPL_E_ARG, "dropout masks are supported in the FP64 parity mode only");
    int64_t want = 0;
    for (int l = 0; l < ws->layers; ++l) want += ws->width[l];
    if (len != want) return fail(OPL_E_ARG, "opl_mlp_set_masks_host: expected %lld mask entries, got %lld", (long long)want, (long long)len);
    for (int l = 0; l + 1 < ws->layers; ++l)
        if (ws->transfer[l] == OPL_TRANSFER_SOFTMAX) return fail(OPL_E_ARG, "dropout on a hidden softmax layer is not supported");
    if (!ws->masks) OPL_CUDA(cudaMalloc(&ws->masks, sizeof(double) * (size_t)want));
    OPL_CUDA(cudaMemcpyAsync(ws->masks, masks, sizeof(double) * (size_t)want, cudaMemcpyHostToDevice, ws->stream));
    OPL_CUDA(cudaStreamSynchronize(ws->stream));
    return OPL_OK;
}

int opl_mlp_profile(opl_mlp *ws, int32_t enable) {
    if (!ws) return fail(OPL_E_ARG, "opl_mlp_profile: null");
    ws->profiling = enable != 0;
    ws->ev_used = 0;
    return OPL_OK;
}

int opl_mlp_profile_read(opl_mlp *ws, int32_t capacity, int32_t *kinds, float *ms, int32_t *count) {
    if (!ws || !kinds || !ms || !count) return fail(OPL_E_ARG, "opl_mlp_profile_read: null pointer");
    OPL_CUDA(cudaStreamSynchronize(ws->stream));
    int32_t out = 0;
    for (size_t i = 0; i + 1 < ws->ev_used && out < capacity; ++i) {
        if (ws->ev_kind[i] == STAGE_END * 100) continue;   // gap between evaluations
        float t = 0.f;
        OPL_CUDA(cudaEventElapsedTime(&t, ws->ev_pool[i], ws->ev_pool[i + 1]));
        kinds[out] = ws->ev_kind[i];
        ms[out] = t;
        ++out;
    }
    *count = out;
    ws->ev_used = 0;
    return OPL_OK;
}

int opl_mlp_set_data_host(opl_mlp *ws, const double *x, const double *y, int64_t rows, int64_t norm_rows) {
    if (!ws || !x || !y) return fail(OPL_E_ARG, "opl_mlp_set_data_host: null pointer");
    if (rows < 1 || rows > ws->max_rows) return fail(OPL_E_ARG, "opl_mlp_set_data_host: rows=%lld outside [1,%lld]",
                                                      (long long)rows, (long long)ws->max_rows);
    const int64_t d0 = ws->width[0], dl = ws->width[ws->layers];
    if (!ws->x_own) OPL_CUDA(cudaMalloc(&ws->x_own, sizeof(double) * (size_t)ws->max_rows * (size_t)d0));
    if (!ws->y_own) OPL_CUDA(cudaMalloc(&ws->y_own, sizeof(double) * (size_t)ws->max_rows * (size_t)dl));
    OPL_CUDA(cudaMemcpyAsync(ws->x_own, x, sizeof(double) * (size_t)rows * (size_t)d0, cudaMemcpyHostToDevice, ws->stream));
    OPL_CUDA(cudaMemcpyAsync(ws->y_own, y, sizeof(double) * (size_t)rows * (size_t)dl, cudaMemcpyHostToDevice, ws->stream));
    OPL_CUDA(cudaStreamSynchronize(ws->stream));
    ws->x = ws->x_own;
    ws->y = ws->y_own;
    ws->rows = rows;
    ws->norm_rows = norm_rows > 0 ? norm_rows : rows;
    if (ws->precision == OPL_PRECISION_TF32) return fast_set_x(ws);
    return int8_set_x(ws);
}

int opl_mlp_set_data_device(opl_mlp *ws, const double *x_dev, const double *y_dev, int64_t rows, int64_t norm_rows) {
    if (!ws || !x_dev || !y_dev) return fail(OPL_E_ARG, "opl_mlp_set_data_device: null pointer");
    if (rows < 1 || rows > ws->max_rows) return fail(OPL_E_ARG, "opl_mlp_set_data_device: rows out of range");
    ws->x = x_dev;   // borrowed: the caller keeps the buffers alive
    ws->y = y_dev;
    ws->rows = rows;
    ws->norm_rows = norm_rows > 0 ? norm_rows : rows;
    if (ws->precision == OPL_PRECISION_TF32) return fast_set_x(ws);
    return int8_set_x(ws);
}

int opl_mlp_eval_device(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev, int32_t want_grad,
                        double *grad_out_dev) {
    if (!ws || !x_dev || !grad_out_dev) return fail(OPL_E_ARG, "opl_mlp_eval_device: null pointer");
    return evaluate(ws, x_dev, alpha, dir_dev, want_grad != 0, grad_out_dev);
}

int opl_mlp_probe_scalars(opl_mlp *ws, const double *grad_dev, const double *dir_dev, int32_t have_grad,
                          double *host_out) {
    if (!ws || !grad_dev || !host_out) return fail(OPL_E_ARG, "opl_mlp_probe_scalars: null pointer");
    return read_scalars(ws, grad_dev, dir_dev, have_grad != 0, host_out);
}

int opl_mlp_probe_device(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev, int32_t want_grad,
                         double *grad_out_dev, double *host_out) {
    if (!ws || !x_dev || !grad_out_dev || !host_out) return fail(OPL_E_ARG, "opl_mlp_probe_device: null pointer");
    if (!ws->x || ws->rows <= 0) return fail(OPL_E_STATE, "no training data resident: call opl_mlp_set_data_* first");
    const SmallPlan pl = small_plan(ws, ws->rows);
    if (pl.ok) {
        double seq = 0.0;
        int rc = small_evaluate(ws, pl, x_dev, alpha, dir_dev, want_grad != 0, grad_out_dev, &seq);
        if (rc != OPL_OK) return rc;
        rc = small_read_scalars(ws, seq, host_out);
        if (!want_grad) host_out[1] = host_out[2] = 0.0;
        return rc;
    }
    int rc = evaluate(ws, x_dev, alpha, dir_dev, want_grad != 0, grad_out_dev);
    if (rc != OPL_OK) return rc;
    return read_scalars(ws, grad_out_dev, dir_dev, want_grad != 0, host_out);
}

int opl_mlp_obj_host(opl_mlp *ws, const double *params, double *obj_out) {
    if (!ws || !params || !obj_out) return fail(OPL_E_ARG, "opl_mlp_obj_host: null pointer");
    // stage the parameters in host_grad (reused as the device-side x)
    OPL_CUDA(cudaMemcpyAsync(ws->host_grad, params, sizeof(double) * (size_t)ws->n, cudaMemcpyHostToDevice, ws->stream));
    // objective lands in w_trial[n]
    int rc = evaluate(ws, ws->host_grad, 0.0, nullptr, false, ws->w_trial);
    if (rc != OPL_OK) return rc;
    double s[3];
    rc = read_scalars(ws, ws->w_trial, nullptr, false, s);
    *obj_out = s[0];
    return rc;
}

int opl_mlp_obj_jac_host(opl_mlp *ws, const double *params, double *obj_out, double *grad_out) {
    if (!ws || !params || !obj_out || !grad_out) return fail(OPL_E_ARG, "opl_mlp_obj_jac_host: null pointer");
    const int64_t n = ws->n;
    // x staged after the trial buffer is formed: use partial scratch-free path -- params go to
    // w_trial directly through the trial kernel reading host_grad, gradient overwrites host_grad.
    OPL_CUDA(cudaMemcpyAsync(ws->host_grad, params, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ws->stream));
    int rc = evaluate(ws, ws->host_grad, 0.0, nullptr, true, ws->host_grad);
    if (rc != OPL_OK) return rc;
    // A page-locked destination (e.g. an array handed out by a pinned allocator) takes the gradient at the full PCIe rate
    // directly; a pageable one is several times slower as a D2H target, so it goes device -> pinned staging -> one host
    // memcpy after the synchronisation that read_scalars performs anyway.
    cudaPointerAttributes attr;
    const bool direct = cudaPointerGetAttributes(&attr, grad_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!direct) {
        cudaGetLastError();   // (older drivers report unregistered host memory as an error)
        if (!ws->pinned_vec) OPL_CUDA(cudaMallocHost(&ws->pinned_vec, sizeof(double) * (size_t)(n + 1)));
    }
    OPL_CUDA(cudaMemcpyAsync(direct ? grad_out : ws->pinned_vec, ws->host_grad, sizeof(double) * (size_t)n,
                             cudaMemcpyDeviceToHost, ws->stream));
    double s[3];
    rc = read_scalars(ws, ws->host_grad, nullptr, false, s);
    if (!direct) memcpy(grad_out, ws->pinned_vec, sizeof(double) * (size_t)n);
    *obj_out = s[0];
    return rc;
}

int opl_mlp_activate_device(opl_mlp *ws, const double *params_dev, const double *x_dev, int64_t rows, double *out_dev) {
    if (!ws || !params_dev || !x_dev || !out_dev) return fail(OPL_E_ARG, "opl_mlp_activate_device: null pointer");
    if (rows < 1 || rows > ws->max_rows) return fail(OPL_E_ARG, "opl_mlp_activate_device: rows out of range");
    if (ws->precision == OPL_PRECISION_TF32) return fast_activate(ws, params_dev, x_dev, rows, out_dev);
    int rc = forward(ws, params_dev, x_dev, rows, false);
    if (rc != OPL_OK) return rc;
    rc = output_stage(ws, nullptr, rows, true, false, nullptr);
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaMemcpyAsync(out_dev, ws->out_act, sizeof(double) * (size_t)rows * (size_t)ws->width[ws->layers],
                             cudaMemcpyDeviceToDevice, ws->stream));
    return OPL_OK;
}

int opl_mlp_activate_host(opl_mlp *ws, const double *params, const double *x, int64_t rows, double *out) {
    if (!ws || !params || !x || !out) return fail(OPL_E_ARG, "opl_mlp_activate_host: null pointer");
    if (rows < 1 || rows > ws->max_rows) return fail(OPL_E_ARG, "opl_mlp_activate_host: rows=%lld outside [1,%lld]",
                                                      (long long)rows, (long long)ws->max_rows);
    const int64_t d0 = ws->width[0], dl = ws->width[ws->layers];
    if (!ws->x_scratch) OPL_CUDA(cudaMalloc(&ws->x_scratch, sizeof(double) * (size_t)ws->max_rows * (size_t)d0));
    OPL_CUDA(cudaMemcpyAsync(ws->w_trial, params, sizeof(double) * (size_t)ws->n, cudaMemcpyHostToDevice, ws->stream));
    OPL_CUDA(cudaMemcpyAsync(ws->x_scratch, x, sizeof(double) * (size_t)rows * (size_t)d0, cudaMemcpyHostToDevice, ws->stream));
    int rc;
    if (ws->precision == OPL_PRECISION_TF32) {
        rc = fast_activate(ws, ws->w_trial, ws->x_scratch, rows, ws->out_act);
    } else {
        rc = forward(ws, ws->w_trial, ws->x_scratch, rows, false);
        if (rc == OPL_OK) rc = output_stage(ws, nullptr, rows, true, false, nullptr);
    }
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaMemcpyAsync(out, ws->out_act, sizeof(double) * (size_t)rows * (size_t)dl, cudaMemcpyDeviceToHost, ws->stream));
    OPL_CUDA(cudaStreamSynchronize(ws->stream));
    return OPL_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Stand-alone host-buffer forms of Transfer.__call__ / derivative and ErrorFunc.__call__ /
// derivative (transfer.py:33-130, error.py:31-116): the same row kernel the MLP uses.
// ---------------------------------------------------------------------------------------
namespace {

struct Scratch {   // RAII device buffers for the one-shot host calls
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    int alloc(void **out, size_t bytes) {
        OPL_CUDA(cudaMalloc(out, bytes ? bytes : 8));
        ptrs.push_back(*out);
        return OPL_OK;
    }
};

int rows_call(int tid, double tparam, int eid, const double *z_host, const double *y_host, const double *g_host,
              int64_t rows, int64_t cols, double *a_out, double *d_out, double *loss_out) {
    if (rows < 1 || cols < 1 || cols > (1 << 24)) return fail(OPL_E_ARG, "bad matrix shape %lld x %lld", (long long)rows, (long long)cols);
    if (opl_device_count() < 1) return fail(OPL_E_CUDA, "no CUDA device: opl_b200 has no CPU fallback");
    Scratch sc;
    const size_t bytes = sizeof(double) * (size_t)rows * (size_t)cols;
    double *z = nullptr, *y = nullptr, *g = nullptr, *a = nullptr, *d = nullptr, *lp = nullptr, *res = nullptr;
    int *flag = nullptr;
    int rc = sc.alloc((void **)&z, bytes);
    if (rc == OPL_OK && y_host) rc = sc.alloc((void **)&y, bytes);
    if (rc == OPL_OK && g_host) rc = sc.alloc((void **)&g, bytes);
    if (rc == OPL_OK && a_out) rc = sc.alloc((void **)&a, bytes);
    if (rc == OPL_OK && d_out) rc = sc.alloc((void **)&d, bytes);
    if (rc == OPL_OK) rc = sc.alloc((void **)&flag, sizeof(int));
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaMemset(flag, 0, sizeof(int)));
    OPL_CUDA(cudaMemcpy(z, z_host, bytes, cudaMemcpyHostToDevice));
    if (y) OPL_CUDA(cudaMemcpy(y, y_host, bytes, cudaMemcpyHostToDevice));
    if (g) OPL_CUDA(cudaMemcpy(g, g_host, bytes, cudaMemcpyHostToDevice));
    RowArgs r = {};
    r.Z = z;
    r.Y = y;
    r.G = g;
    r.A = a;
    r.D = d;
    r.rows = rows;
    r.C = (int)cols;
    r.ld = cols;
    r.tid = tid;
    r.tparam = tparam;
    r.eid = eid;
    r.ce_div = -(double)rows;
    r.mse_mul = 2.0 / ((double)rows * (double)cols);
    r.flag = flag;
    r.group = row_group(r.C);
    const int max_grid = 8 * sm_count();
    if (loss_out) {
        rc = sc.alloc((void **)&lp, sizeof(double) * (size_t)max_grid);
        if (rc == OPL_OK) rc = sc.alloc((void **)&res, sizeof(double));
        if (rc != OPL_OK) return rc;
        r.loss_partial = lp;
    }
    const int grid = launch_rows(r, max_grid, nullptr);
    OPL_LAUNCH_CHECK();
    if (loss_out) {
        loss_final_kernel<<<1, 256>>>(lp, grid, eid, (double)rows, (double)cols, res, nullptr);
        OPL_LAUNCH_CHECK();
        OPL_CUDA(cudaMemcpy(loss_out, res, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (a_out) OPL_CUDA(cudaMemcpy(a_out, a, bytes, cudaMemcpyDeviceToHost));
    if (d_out) OPL_CUDA(cudaMemcpy(d_out, d, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    OPL_CUDA(cudaMemcpy(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail(OPL_E_DOMAIN, "invalid value encountered in log (negative prediction in cross entropy)");
    return OPL_OK;
}

}  // namespace

extern "C" {

int opl_transfer_apply_host(int32_t transfer_id, double param, const double *x, int64_t rows, int64_t cols,
                            double *out) {
    if (!x || !out || transfer_id < 0 || transfer_id > OPL_TRANSFER_LOGIT)
        return fail(OPL_E_ARG, "opl_transfer_apply_host: bad arguments");
    return rows_call(transfer_id, param, OPL_ERROR_MSE, x, nullptr, nullptr, rows, cols, out, nullptr, nullptr);
}

int opl_transfer_backprop_host(int32_t transfer_id, double param, const double *x, const double *upstream,
                               int64_t rows, int64_t cols, double *out) {
    if (!x || !upstream || !out || transfer_id < 0 || transfer_id > OPL_TRANSFER_LOGIT)
        return fail(OPL_E_ARG, "opl_transfer_backprop_host: bad arguments");
    return rows_call(transfer_id, param, OPL_ERROR_MSE, x, nullptr, upstream, rows, cols, nullptr, out, nullptr);
}

int opl_error_host(int32_t error_id, const double *a, const double *b, int64_t rows, int64_t cols,
                   double *err_out, double *deriv_out) {
    if (!a || !b || !err_out || (error_id != OPL_ERROR_MSE && error_id != OPL_ERROR_CROSS_ENTROPY))
        return fail(OPL_E_ARG, "opl_error_host: bad arguments");
    return rows_call(OPL_TRANSFER_LINEAR, 0.0, error_id, a, b, nullptr, rows, cols, nullptr, deriv_out, err_out);
}

int opl_softmax_jacobian_host(const double *y, int64_t rows, int64_t cols, double *jac_out) {
    if (!y || !jac_out || rows < 1 || cols < 1) return fail(OPL_E_ARG, "opl_softmax_jacobian_host: bad arguments");
    if (opl_device_count() < 1) return fail(OPL_E_CUDA, "no CUDA device: opl_b200 has no CPU fallback");
    Scratch sc;
    double *yd = nullptr, *jd = nullptr;
    const size_t yb = sizeof(double) * (size_t)rows * (size_t)cols, jb = yb * (size_t)cols;
    int rc = sc.alloc((void **)&yd, yb);
    if (rc == OPL_OK) rc = sc.alloc((void **)&jd, jb);
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaMemcpy(yd, y, yb, cudaMemcpyHostToDevice));
    const int64_t total = rows * cols * cols;
    softmax_jacobian_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 8 * (int64_t)sm_count()), 256>>>(
        yd, rows, (int)cols, jd);
    OPL_LAUNCH_CHECK();
    OPL_CUDA(cudaMemcpy(jac_out, jd, jb, cudaMemcpyDeviceToHost));
    return OPL_OK;
}

}  // extern "C"
