// FP64 tensor-core GEMM family for the parity mode (sm_100a).
//
// tcgen05.mma has no f64 kind (kinds are tf32/f16/i8/f8f6f4/mx*), so FP64 tensor math on
// Blackwell is the warp-level DMMA: mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4.  This is
// a register-fragment kernel: cp.async (LDGSTS) multi-stage ring -> padded shared tiles ->
// DMMA, with the layer's bias / transfer function / transfer derivative fused into the
// epilogue so activations make exactly one trip to HBM.
//
// Three operand layouts cover the whole MLP (all matrices row-major, as NumPy holds them):
//   NN  C[M,N] = A[M,K] . B[K,N]      forward      Z_l  = A_l . W_l          (mlp.py:150-160)
//   NT  C[M,N] = A[M,K] . B[N,K]^T    backprop     dA   = delta . W^T        (mlp.py:258)
//   TN  C[M,N] = A[K,M]^T . B[K,N]    weight grad  dW_l = A_l^T . delta_l    (mlp.py:270)
//       (reduction over the patterns: split-K over blockIdx.z into partials, reduced in a
//        fixed order by reduce_partials_kernel => deterministic gradients)
#pragma once
#include "common.cuh"

namespace opl {

enum GemmLayout { GEMM_NN = 0, GEMM_NT = 1, GEMM_TN = 2 };

enum GemmEpilogue {
    EPI_STORE = 0,     // C = acc (+ bias[col])
    EPI_TANH = 1,      // C = tanh(z)
    EPI_SOFTPLUS = 2,  // C = log(1+exp(z)) (inf -> z);  aux_out = 1/(1+exp(-z))
    EPI_GAUSSIAN = 3,  // C = exp(-(z*z/v));             aux_out = -2 z C / v
    EPI_MUL_AUX = 4,   // C = acc * aux_in               (delta = upstream o f'(z))
    EPI_MUL_DTANH = 5  // C = acc * (1 - aux_in^2)       (tanh derivative from the OUTPUT)
};

struct GemmArgs {
    const double *A;
    const double *B;
    double *C;
    int64_t lda, ldb, ldc;
    int M, N, K;
    int k_per_split;        // multiple of 16; == K rounded up when not split
    int64_t split_stride;   // elements between the partial C of consecutive splits
    const double *bias;     // length N or null
    const double *aux_in;   // M x N (ld = ldaux) or null
    double *aux_out;        // M x N (ld = ldaux) or null
    int64_t ldaux;
    int epi;
    double tparam;          // gaussian variance
};

constexpr int GEMM_BK = 16;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4) * B(4x8), FP64.  Fragment ownership (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Copy one operand tile global -> shared.
//   KINNER : tile is [MN rows][BK] (k contiguous in global), row stride LD
//   !KINNER: tile is [BK rows][MN] (mn contiguous in global), row stride LD
template <bool KINNER, int MN_TILE, int LD, int VEC, int NTHREADS>
__device__ __forceinline__ void load_tile(double *dst, const double *__restrict__ src, int64_t ld,
                                          int mn0, int mn_end, int k0, int k_end, int tid) {
    if (KINNER) {
        constexpr int CPR = GEMM_BK / VEC;
        constexpr int TOTAL = MN_TILE * CPR;
#pragma unroll
        for (int base = 0; base < TOTAL; base += NTHREADS) {
            int idx = base + tid;
            if (TOTAL % NTHREADS != 0 && idx >= TOTAL) break;
            int r = idx / CPR, c = (idx % CPR) * VEC;
            int gm = mn0 + r, gk = k0 + c;
            bool ok = gm < mn_end && gk < k_end;
            const double *g = ok ? src + (int64_t)gm * ld + gk : src;
            if (VEC == 2) cp_async16(dst + r * LD + c, g, ok);
            else cp_async8(dst + r * LD + c, g, ok);
        }
    } else {
        constexpr int CPR = MN_TILE / VEC;
        constexpr int TOTAL = GEMM_BK * CPR;
#pragma unroll
        for (int base = 0; base < TOTAL; base += NTHREADS) {
            int idx = base + tid;
            if (TOTAL % NTHREADS != 0 && idx >= TOTAL) break;
            int r = idx / CPR, c = (idx % CPR) * VEC;
            int gk = k0 + r, gm = mn0 + c;
            bool ok = gk < k_end && gm < mn_end;
            const double *g = ok ? src + (int64_t)gk * ld + gm : src;
            if (VEC == 2) cp_async16(dst + r * LD + c, g, ok);
            else cp_async8(dst + r * LD + c, g, ok);
        }
    }
}

__device__ __forceinline__ void epilogue_element(const GemmArgs &p, int row, int col, double acc) {
    double z = acc;
    if (p.bias) z += p.bias[col];
    const int64_t ci = (int64_t)row * p.ldc + col;
    const int64_t xi = (int64_t)row * p.ldaux + col;
    switch (p.epi) {
        case EPI_STORE:
            p.C[ci] = z;
            break;
        case EPI_TANH:
            p.C[ci] = tanh(z);
            break;
        case EPI_SOFTPLUS: {   // calculate.py:128-141
            double r = log(1.0 + exp(z));
            if (isinf(r)) r = z;
            p.C[ci] = r;
            if (p.aux_out) p.aux_out[xi] = 1.0 / (1.0 + exp(-z));
            break;
        }
        case EPI_GAUSSIAN: {   // calculate.py:101-106
            double a = exp(-(z * z / p.tparam));
            p.C[ci] = a;
            if (p.aux_out) p.aux_out[xi] = -2.0 * z * a / p.tparam;
            break;
        }
        case EPI_MUL_AUX:
            p.C[ci] = z * p.aux_in[xi];
            break;
        case EPI_MUL_DTANH: {
            double a = p.aux_in[xi];
            p.C[ci] = z * (1.0 - a * a);
            break;
        }
    }
}

template <int LAYOUT, int BM, int BN, int WM, int WN, int VEC, int STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_f64_kernel(const GemmArgs p) {
    constexpr int BK = GEMM_BK;
    constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
    constexpr bool A_KINNER = (LAYOUT != GEMM_TN);
    constexpr bool B_KINNER = (LAYOUT == GEMM_NT);
    constexpr int A_LD = A_KINNER ? BK + 4 : BM + 4;   // == 4 (mod 16): conflict-free 8-byte fragment loads
    constexpr int B_LD = B_KINNER ? BK + 4 : BN + 4;
    constexpr int A_TILE = (A_KINNER ? BM : BK) * A_LD;
    constexpr int B_TILE = (B_KINNER ? BN : BK) * B_LD;
    constexpr int MI = WM / 8, NI = WN / 8;

    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_TILE;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / (BN / WN)) * WM;
    const int wn0 = (warp % (BN / WN)) * WN;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * p.k_per_split;
    const int k_end = min(p.K, k_begin + p.k_per_split);
    const int nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int kt) {
        const int st = kt % STAGES;
        const int k0 = k_begin + kt * BK;
        load_tile<A_KINNER, BM, A_LD, VEC, NTHREADS>(As + st * A_TILE, p.A, p.lda, m0, p.M, k0, k_end, tid);
        load_tile<B_KINNER, BN, B_LD, VEC, NTHREADS>(Bs + st * B_TILE, p.B, p.ldb, n0, p.N, k0, k_end, tid);
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) issue(s);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < nk) issue(kt + STAGES - 1);
        cp_async_commit();

        const double *a = As + (kt % STAGES) * A_TILE;
        const double *b = Bs + (kt % STAGES) * B_TILE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i)
                af[i] = A_KINNER ? a[(wm0 + 8 * i + g) * A_LD + kk * 4 + t]
                                 : a[(kk * 4 + t) * A_LD + wm0 + 8 * i + g];
#pragma unroll
            for (int j = 0; j < NI; ++j)
                bf[j] = B_KINNER ? b[(wn0 + 8 * j + g) * B_LD + kk * 4 + t]
                                 : b[(kk * 4 + t) * B_LD + wn0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    GemmArgs q = p;
    q.C = p.C + (int64_t)blockIdx.z * p.split_stride;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm0 + 8 * i + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = n0 + wn0 + 8 * j + 2 * t;
            if (col < p.N) epilogue_element(q, row, col, acc[i][j][0]);
            if (col + 1 < p.N) epilogue_element(q, row, col + 1, acc[i][j][1]);
        }
    }
}

// out[i] = sum_s partial[s*stride + i], fixed order => bit-reproducible gradients
int launch_reduce_partials(const double *partial, int splits, int64_t stride, int64_t count, double *out,
                           cudaStream_t stream, int64_t *launches);

// Launch helper: picks the tile shape from N, the vector width from alignment.
// `*launches` is incremented by the number of kernels enqueued.
int launch_gemm_f64(int layout, const GemmArgs &args, int splits, cudaStream_t stream, int64_t *launches);

// split-K plan for the TN (dW) GEMM
struct SplitPlan {
    int splits;
    int k_per_split;
};
SplitPlan plan_split_k(int M, int N, int K);

}  // namespace opl
