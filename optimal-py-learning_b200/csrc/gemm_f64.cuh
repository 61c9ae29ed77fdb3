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
    // TN only: CTAs of the first M-tile also accumulate the column sums of their B tiles
    // (db = 1^T delta_0, mlp.py:276) into colsum_out[split * split_stride + col]
    double *colsum_out;
    // forward only: DropoutTransfer (mlp.py:429-445) multiplies the transfer OUTPUT by a 0/1 vector of
    // active neurons; the derivative is evaluated on that masked output (matters for tanh / gaussian)
    const double *colmask;
};

constexpr int GEMM_BK = 16;

// largest power of two <= min(avail, BK): thread groups that split the BK rows of a B tile evenly
__host__ __device__ constexpr int colsum_groups(int avail) {
    int g = 1;
    while (g * 2 <= avail && g * 2 <= GEMM_BK) g *= 2;
    return g;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
// same with the destination as a shared-window address (callers that step through a row keep it in a register)
__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void *gsrc, bool valid) {
    int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gsrc), "r"(n));
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

// Epilogue over one staged accumulator tile (rows x 2^log2_cols, row stride c_ld in shared
// memory).  Deliberately a compact, non-unrolled loop in ONE non-inlined function: the first
// version inlined exp/log per accumulator register (64x per thread) and produced a 511 KB kernel
// whose epilogue stalled on instruction fetch (profiles/r01_*).  Consecutive threads take
// consecutive columns: conflict-free shared reads, fully coalesced global stores / aux loads.
static __device__ __noinline__ void epilogue_tile(const GemmArgs &p, const double *__restrict__ Cs, int c_ld, int row0,
                                           int col0, int rows, int log2_cols, int tid, int nthreads,
                                           int split) {
    const int cols = 1 << log2_cols;
    const int rmax = min(rows, p.M - row0), cmax = min(cols, p.N - col0);
    if (rmax <= 0 || cmax <= 0) return;
    double *C = p.C + (int64_t)split * p.split_stride;   // may alias p.aux_in (in-place delta)
    const int total = rmax << log2_cols;
    const double *bias = p.bias;
    // four elements in flight per thread: the aux loads / stores of the memory-bound epilogues
    // (K <= 16 GEMMs are pure streaming) need memory-level parallelism, not instruction count
#define OPL_EPI_LOOP(PRELOAD, BODY)                                             \
    for (int base = 0; base < total; base += 4 * nthreads) {                    \
        double zz[4], xx[4];                                                    \
        int64_t cc[4], xc[4];                                                   \
        int cg[4];                                                              \
        bool ok[4];                                                             \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                         \
            const int idx = base + u * nthreads + tid;                          \
            const int r = idx >> log2_cols, c = idx & (cols - 1);               \
            ok[u] = idx < total && c < cmax;                                    \
            cg[u] = col0 + c;                                                   \
            cc[u] = (int64_t)(row0 + r) * p.ldc + col0 + c;                     \
            xc[u] = (int64_t)(row0 + r) * p.ldaux + col0 + c;                   \
            zz[u] = ok[u] ? Cs[r * c_ld + c] : 0.0;                             \
            if (ok[u] && bias) zz[u] += bias[col0 + c];                         \
            xx[u] = 0.0;                                                        \
            if (ok[u]) { const int64_t xi = xc[u]; (void)xi; PRELOAD }          \
        }                                                                       \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                         \
            if (ok[u]) {                                                        \
                const double z = zz[u], x = xx[u];                              \
                const int64_t ci = cc[u], xi = xc[u];                           \
                const int gc = cg[u];                                           \
                (void)x; (void)xi; (void)gc;                                    \
                BODY                                                            \
            }                                                                   \
        }                                                                       \
    }
    switch (p.epi) {
        case EPI_STORE:
            if (p.colmask) { OPL_EPI_LOOP(, C[ci] = z * p.colmask[gc];) }
            else { OPL_EPI_LOOP(, C[ci] = z;) }
            break;
        case EPI_TANH:
            if (p.colmask) { OPL_EPI_LOOP(, C[ci] = tanh(z) * p.colmask[gc];) }
            else { OPL_EPI_LOOP(, C[ci] = tanh(z);) }
            break;
        case EPI_SOFTPLUS: {   // calculate.py:128-141
            double *__restrict__ aux = p.aux_out;
            // one exp serves both: sigmoid(z) = e/(1+e) (== 1/(1+exp(-z)) to rounding; 1 once e overflows)
            const double *cm = p.colmask;
            OPL_EPI_LOOP(, const double e = exp(z); const double d_ = 1.0 + e; double r_ = log(d_);
                         if (isinf(r_)) r_ = z; if (cm) r_ *= cm[gc]; C[ci] = r_;
                         if (aux) aux[xi] = isinf(e) ? 1.0 : e / d_;)
            break;
        }
        case EPI_GAUSSIAN: {   // calculate.py:101-106
            double *__restrict__ aux = p.aux_out;
            const double v = p.tparam;
            const double *cm = p.colmask;
            OPL_EPI_LOOP(, double a = exp(-(z * z / v)); if (cm) a *= cm[gc]; C[ci] = a;
                         if (aux) aux[xi] = -2.0 * z * a / v;)
            break;
        }
        case EPI_MUL_AUX: {
            const double *aux = p.aux_in;   // may alias C (delta overwrites D in place)
            OPL_EPI_LOOP(xx[u] = aux[xi];, C[ci] = z * x;)
            break;
        }
        case EPI_MUL_DTANH: {
            const double *aux = p.aux_in;
            OPL_EPI_LOOP(xx[u] = aux[xi];, C[ci] = z * (1.0 - x * x);)
            break;
        }
    }
#undef OPL_EPI_LOOP
}

template <int LAYOUT, int BM, int BN, int WM, int WN, int VEC, int STAGES, int MINBLOCKS>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINBLOCKS)
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
    // accumulator staging for the epilogue reuses the pipeline buffers
    constexpr int C_LD = BN + 8;                        // == 8 (mod 16): conflict-free 16-byte fragment stores
    constexpr int SMEM_DOUBLES = STAGES * (A_TILE + B_TILE);
    constexpr int PASS_ROWS = (BM * C_LD <= SMEM_DOUBLES) ? BM : ((BM / 2) * C_LD <= SMEM_DOUBLES ? BM / 2 : BM / 4);
    static_assert(PASS_ROWS * C_LD <= SMEM_DOUBLES, "accumulator tile does not fit the pipeline buffers");
    static_assert(PASS_ROWS % WM == 0, "a warp's rows must fall into one epilogue pass");
    constexpr int LOG2_BN = BN == 128 ? 7 : (BN == 64 ? 6 : (BN == 32 ? 5 : 4));
    static_assert((1 << LOG2_BN) == BN, "BN must be 16, 32, 64 or 128");

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

    const bool do_colsum = LAYOUT == GEMM_TN && p.colsum_out != nullptr && blockIdx.y == 0;
    double colsum = 0.0;

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

#pragma unroll 1
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < nk) issue(kt + STAGES - 1);
        cp_async_commit();

        const double *a = As + (kt % STAGES) * A_TILE;
        const double *b = Bs + (kt % STAGES) * B_TILE;
        if (LAYOUT == GEMM_TN && do_colsum) {   // zero-filled rows beyond k_end add nothing
            constexpr int GROUPS = colsum_groups(NTHREADS / BN);
            constexpr int PER = BK / GROUPS;
            const int col = tid % BN, grp = tid / BN;
            if (grp < GROUPS) {
#pragma unroll
                for (int k = 0; k < PER; ++k) colsum += b[(grp * PER + k) * B_LD + col];
            }
        }
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
    __syncthreads();   // every warp is done with the operand tiles: reuse them for the accumulators
    if (LAYOUT == GEMM_TN && do_colsum) {   // fold the per-thread column partials in a fixed order
        constexpr int GROUPS = colsum_groups(NTHREADS / BN);
        double *fold = smem;
        if (tid / BN < GROUPS) fold[tid] = colsum;
        __syncthreads();
        if (tid < BN && n0 + tid < p.N) {
            double t_ = 0.0;
#pragma unroll
            for (int gidx = 0; gidx < GROUPS; ++gidx) t_ += fold[gidx * BN + tid];
            p.colsum_out[(int64_t)blockIdx.z * p.split_stride + n0 + tid] = t_;
        }
        __syncthreads();
    }

    double *Cs = smem;
#pragma unroll
    for (int pass = 0; pass < BM / PASS_ROWS; ++pass) {
        if (wm0 / PASS_ROWS == pass) {   // warp-uniform
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j)
                    *reinterpret_cast<double2 *>(Cs + (wm0 - pass * PASS_ROWS + 8 * i + g) * C_LD + wn0 + 8 * j + 2 * t) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
        }
        __syncthreads();
        epilogue_tile(p, Cs, C_LD, m0 + pass * PASS_ROWS, n0, PASS_ROWS, LOG2_BN, tid, NTHREADS, blockIdx.z);
        if (pass + 1 < BM / PASS_ROWS) __syncthreads();
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
// N-tile width / M-tile height the launcher will use
int gemm_tile_n(int N);
int gemm_tile_m(int layout, int M, int N);
SplitPlan plan_split_k(int M, int N, int K);

}  // namespace opl
