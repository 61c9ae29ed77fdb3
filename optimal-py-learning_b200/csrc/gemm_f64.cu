// Instantiations + launch logic for the FP64 DMMA GEMM family (see gemm_f64.cuh).
#include <stdlib.h>

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

// Same sum with one warp per output element: lanes stride over the splits, then a shuffle tree.
// Used when there are few outputs and many splits (a serial loop would chain `splits` dependent
// loads).  The order depends only on `splits`, so results stay bit-reproducible.
__global__ void reduce_partials_warp_kernel(const double *__restrict__ partial, int splits, int64_t stride,
                                            int64_t count, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < count; i += nwarps) {
        double s = 0.0;
        for (int k = lane; k < splits; k += 32) s += partial[(int64_t)k * stride + i];
        s = warp_sum(s);
        if (lane == 0) out[i] = s;
    }
}

int launch_reduce_partials(const double *partial, int splits, int64_t stride, int64_t count, double *out,
                           cudaStream_t stream, int64_t *launches) {
    if (count <= 0) return OPL_OK;
    const int64_t cap = 4 * (int64_t)sm_count();
    if (splits >= 16 && count * 32 <= cap * 256 * 4) {
        int64_t grid = ceil_div(count * 32, 256);
        if (grid > cap) grid = cap;
        reduce_partials_warp_kernel<<<(unsigned)grid, 256, 0, stream>>>(partial, splits, stride, count, out);
        OPL_LAUNCH_CHECK();
        if (launches) ++*launches;
        return OPL_OK;
    }
    int64_t grid = ceil_div(count, 256);
    if (grid > cap) grid = cap;
    reduce_partials_kernel<<<(unsigned)grid, 256, 0, stream>>>(partial, splits, stride, count, out);
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

// Tile policy.  Default: 128x64 CTA tiles (8 warps of 32x32, <=128 registers) so TWO CTAs share an
// SM and one CTA's epilogue / pipeline fill overlaps the other's DMMA stream; OPL_GEMM_TILE=128
// selects the 128x128 (64x32 warp tile, one CTA per SM) variant for A/B measurements.
static int wide_tile() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("OPL_GEMM_TILE");
        cached = (e && atoi(e) == 128) ? 1 : 0;
    }
    return cached;
}

int gemm_tile_n(int N) {
    if (N <= 16) return 16;
    if (N <= 64 || !wide_tile()) return 64;
    return 128;
}

// M-tile height.  The weight-gradient GEMM (TN) has M = layer width: 112-row tiles (8 warps of
// 56x16, balanced over the 4 SM sub-partitions) avoid padding when they divide M better than 128 does (784 = 7 x 112 vs 7 x 128 = 896).
int gemm_tile_m(int layout, int M, int N) {
    if (layout != GEMM_TN || gemm_tile_n(N) != 64) return 128;
    const int64_t pad128 = ceil_div(M, 128) * 128, pad112 = ceil_div(M, 112) * 112;
    return pad112 < pad128 ? 112 : 128;
}

SplitPlan plan_split_k(int M, int N, int K) {
    const int bn = gemm_tile_n(N);
    const int64_t tiles = ceil_div(M, gemm_tile_m(GEMM_TN, M, N)) * ceil_div(N, bn);
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

template <int LAYOUT, int BM, int BN, int WM, int WN, int VEC, int MINBLOCKS, int STAGES = kStages>
int launch_one(const GemmArgs &a, int splits, cudaStream_t stream) {
    constexpr bool A_KINNER = (LAYOUT != GEMM_TN);
    constexpr bool B_KINNER = (LAYOUT == GEMM_NT);
    constexpr int A_TILE = A_KINNER ? BM * (GEMM_BK + 4) : GEMM_BK * (BM + 4);
    constexpr int B_TILE = B_KINNER ? BN * (GEMM_BK + 4) : GEMM_BK * (BN + 4);
    constexpr int SMEM = STAGES * (A_TILE + B_TILE) * (int)sizeof(double);
    constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
    auto kern = gemm_f64_kernel<LAYOUT, BM, BN, WM, WN, VEC, STAGES, MINBLOCKS>;
    static bool configured_dev[64] = {};     // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[current_device_slot()];
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
    const int bn = gemm_tile_n(a.N);
    // narrow outputs are pure streaming: 2 stages / 4 CTAs per SM keep every tile of a 60000-row
    // forward in ONE wave (469 tiles on 592 slots) instead of 1.06 waves on 444
    if (bn == 16 && LAYOUT == GEMM_TN) return launch_one<LAYOUT, 128, 16, 16, 16, VEC, 3, 3>(a, splits, stream);
    if (bn == 16) return launch_one<LAYOUT, 128, 16, 16, 16, VEC, 4, 2>(a, splits, stream);
    if (LAYOUT == GEMM_TN && bn == 64 && gemm_tile_m(LAYOUT, a.M, a.N) == 112)
        return launch_one<LAYOUT, 112, 64, 56, 16, VEC, 2>(a, splits, stream);
    if (bn == 64) return launch_one<LAYOUT, 128, 64, 32, 32, VEC, 2>(a, splits, stream);
    return launch_one<LAYOUT, 128, 128, 64, 32, VEC, 1>(a, splits, stream);
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

// Benchmark hook: time `iters` launches of one plain GEMM (no epilogue fusion) on device buffers
// filled by the caller.  Used by tools/ and bench.py for shape sweeps; not part of the training path.
extern "C" int opl_bench_gemm_f64(int32_t layout, int32_t M, int32_t N, int32_t K, const double *a_dev,
                                  const double *b_dev, double *c_dev, double *scratch_dev, int64_t scratch_len,
                                  int32_t iters, float *ms_out) {
    using namespace opl;
    if (!a_dev || !b_dev || !c_dev || !ms_out || layout < 0 || layout > 2 || iters < 1)
        return fail(OPL_E_ARG, "opl_bench_gemm_f64: bad arguments");
    GemmArgs g = {};
    g.A = a_dev;
    g.B = b_dev;
    g.C = c_dev;
    g.M = M;
    g.N = N;
    g.K = K;
    g.lda = layout == GEMM_TN ? M : K;
    g.ldb = layout == GEMM_NT ? K : N;
    g.ldc = N;
    g.ldaux = N;
    g.epi = EPI_STORE;
    g.k_per_split = (int)(ceil_div(K, GEMM_BK) * GEMM_BK);
    int splits = 1;
    if (layout == GEMM_TN && scratch_dev) {
        SplitPlan plan = plan_split_k(M, N, K);
        if ((int64_t)plan.splits * M * N <= scratch_len && plan.splits > 1) {
            splits = plan.splits;
            g.k_per_split = plan.k_per_split;
            g.C = scratch_dev;
            g.split_stride = (int64_t)M * N;
        }
    }
    cudaEvent_t e0, e1;
    OPL_CUDA(cudaEventCreate(&e0));
    OPL_CUDA(cudaEventCreate(&e1));
    int rc = launch_gemm_f64(layout, g, splits, nullptr, nullptr);   // warm-up
    if (rc != OPL_OK) return rc;
    OPL_CUDA(cudaEventRecord(e0, nullptr));
    for (int i = 0; i < iters; ++i) {
        rc = launch_gemm_f64(layout, g, splits, nullptr, nullptr);
        if (rc != OPL_OK) return rc;
        if (splits > 1) rc = launch_reduce_partials(scratch_dev, splits, (int64_t)M * N, (int64_t)M * N, c_dev, nullptr, nullptr);
        if (rc != OPL_OK) return rc;
    }
    OPL_CUDA(cudaEventRecord(e1, nullptr));
    OPL_CUDA(cudaEventSynchronize(e1));
    OPL_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    *ms_out /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return OPL_OK;
}
