// Instantiations + launch logic for the FP64 DMMA GEMM family (see gemm_f64.cuh).
#include "gemm_f64.cuh"

namespace opl {

__global__ void reduce_partials_kernel(const double *__restrict__ partial, int splits, int64_t stride,
                                       int64_t count, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * stride + i];
        out[i] = s;
    }
}

int launch_reduce_partials(const double *partial, int splits, int64_t stride, int64_t count, double *out,
                           cudaStream_t stream, int64_t *launches) {
    if (count <= 0) return OPL_OK;
    int64_t grid = ceil_div(count, 256);
    const int64_t cap = 4 * (int64_t)sm_count();
    if (grid > cap) grid = cap;
    reduce_partials_kernel<<<(unsigned)grid, 256, 0, stream>>>(partial, splits, stride, count, out);
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

SplitPlan plan_split_k(int M, int N, int K) {
    const int bn = N <= 16 ? 16 : (N <= 64 ? 64 : 128);
    const int64_t tiles = ceil_div(M, 128) * ceil_div(N, bn);
    int64_t want = (2 * (int64_t)sm_count()) / (tiles > 0 ? tiles : 1);
    int64_t cap = K / 256;   // keep at least 256 patterns per split
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    int kps = (int)(ceil_div(ceil_div(K, want), GEMM_BK) * GEMM_BK);
    if (kps < GEMM_BK) kps = GEMM_BK;
    SplitPlan plan;
    plan.k_per_split = kps;
    plan.splits = (int)ceil_div(K > 0 ? K : 1, kps);
    return plan;
}

namespace {

constexpr int kStages = 3;

template <int LAYOUT, int BM, int BN, int WM, int WN, int VEC>
int launch_one(const GemmArgs &a, int splits, cudaStream_t stream) {
    constexpr bool A_KINNER = (LAYOUT != GEMM_TN);
    constexpr bool B_KINNER = (LAYOUT == GEMM_NT);
    constexpr int A_TILE = A_KINNER ? BM * (GEMM_BK + 4) : GEMM_BK * (BM + 4);
    constexpr int B_TILE = B_KINNER ? BN * (GEMM_BK + 4) : GEMM_BK * (BN + 4);
    constexpr int SMEM = kStages * (A_TILE + B_TILE) * (int)sizeof(double);
    constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
    auto kern = gemm_f64_kernel<LAYOUT, BM, BN, WM, WN, VEC, kStages>;
    static bool configured = false;
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, BM), (unsigned)splits);
    kern<<<grid, NTHREADS, SMEM, stream>>>(a);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

template <int LAYOUT, int VEC>
int launch_layout(const GemmArgs &a, int splits, cudaStream_t stream) {
    if (a.N <= 16) return launch_one<LAYOUT, 128, 16, 16, 16, VEC>(a, splits, stream);
    if (a.N <= 64) return launch_one<LAYOUT, 128, 64, 32, 32, VEC>(a, splits, stream);
    return launch_one<LAYOUT, 128, 128, 64, 32, VEC>(a, splits, stream);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int launch_gemm_f64(int layout, const GemmArgs &args, int splits, cudaStream_t stream, int64_t *launches) {
    if (args.M <= 0 || args.N <= 0) return OPL_OK;
    // 16-byte cp.async needs: aligned base, even leading dimension, even extent along the
    // contiguous axis (so a pair never straddles the matrix edge).
    const bool a_kinner = layout != GEMM_TN, b_kinner = layout == GEMM_NT;
    const int a_contig = a_kinner ? args.K : args.M;
    const int b_contig = b_kinner ? args.K : args.N;
    const bool vec = aligned16(args.A) && aligned16(args.B) && args.lda % 2 == 0 && args.ldb % 2 == 0 &&
                     a_contig % 2 == 0 && b_contig % 2 == 0;
    int rc;
    if (layout == GEMM_NN)
        rc = vec ? launch_layout<GEMM_NN, 2>(args, splits, stream) : launch_layout<GEMM_NN, 1>(args, splits, stream);
    else if (layout == GEMM_NT)
        rc = vec ? launch_layout<GEMM_NT, 2>(args, splits, stream) : launch_layout<GEMM_NT, 1>(args, splits, stream);
    else
        rc = vec ? launch_layout<GEMM_TN, 2>(args, splits, stream) : launch_layout<GEMM_TN, 1>(args, splits, stream);
    if (rc == OPL_OK && launches) ++*launches;
    return rc;
}

}  // namespace opl
