// FP64-accurate GEMM on the INT8 tensor cores (Ozaki splitting) -- sm_100a, tcgen05.mma.kind::i8.
//
// Why: the FP64 tensor path of B200 (DMMA) peaks at 37 TFLOP/s; the INT8 path offers ~60x the MAC rate.
// How: C = A.B with every ROW of A and every COLUMN of B scaled by a power of two and cut into S = 7
// signed 7-bit slices (an error-free transformation: slice p is the exact truncation of the scaled
// remainder, so A = 2^ea * sum_p 2^-7(p+1) A_p up to 2^-49 of the row maximum).  The S(S+1)/2 products
// A_p.B_q with p+q < S are exact INT32 GEMMs; products with the same p+q share a scale, so ONE TMEM
// accumulator per diagonal d = p+q (7 x 64 INT32 columns = 448 of the 512 TMEM columns) receives them
// all, and the epilogue recombines  C = 2^(ea_row + eb_col) * sum_d 2^-7(d+2) acc_d  in FP64 and applies
// the same fused epilogues as the DMMA family.  Measured accuracy vs FP64 (tools/ozaki_proto.py):
// 1e-14 on the forward shape, 7e-13 on a heavy-tailed weight-gradient (gate 1e-10).
//
// INT32 exactness: one product contributes at most K*127^2 per element and a diagonal sums up to 7 of them,
// so a CTA never contracts more than OZ_MAX_K = 16384 (7*16384*16129 < 2^31); longer contractions are
// split over blockIdx.z and folded in FP64 by the caller.
//
// Pipeline per CTA (one 128 x 64 output tile, 320 threads):
//   warp 0      TMA producer: per 128-wide k-block first the 7 B tiles (64 x 128 B each, one barrier, double
//               buffered), then the 7 A tiles (128 x 128 B each) through a 4-deep ring
//   warp 1      MMA issuer: when A_p lands, tcgen05.mma (M128,N64,K32) x4 against B_0..B_{S-1-p} into
//               accumulator p+q; commit frees the A stage; after A_6 a commit frees the B buffer
//   warps 2..9  epilogue: warp w owns TMEM lanes 32*(w%4).. and column half (w-2)/4
#include <cudaTypedefs.h>

#include <algorithm>
#include <vector>

#include "ozaki.cuh"

namespace opl {

namespace {

constexpr int OZ_A_BYTES = OZ_BM * 128;        // 16 KB: 128 rows x 128 int8
constexpr int OZ_B_BYTES = OZ_BN * 128;        //  8 KB
constexpr int OZ_A_STAGES = 4;
constexpr int OZ_SMEM = OZ_A_STAGES * OZ_A_BYTES + 2 * OZ_S * OZ_B_BYTES + 1024 + 256;
constexpr int OZ_THREADS = 320;

__device__ __forceinline__ uint32_t oz_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void oz_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool oz_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(oz_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void oz_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !oz_try_wait(bar, parity); ++spin)
        if (spin > (1u << 28)) __trap();      // protocol bug => trap instead of hanging the GPU
}
__device__ __forceinline__ void oz_tma_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(oz_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}
// D[tmem, INT32] (+)= A[smem, INT8] . B[smem, INT8], 128 x 64 x 32
__device__ __forceinline__ void oz_umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (see gemm_tf32.cu)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(16u >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzakiArgs p) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem;                                   // OZ_A_STAGES x 16 KB
    uint8_t *sB = smem + OZ_A_STAGES * OZ_A_BYTES;        // 2 x 7 x 8 KB
    uint64_t *a_full = reinterpret_cast<uint64_t *>(sB + 2 * OZ_S * OZ_B_BYTES);
    uint64_t *a_empty = a_full + OZ_A_STAGES;
    uint64_t *b_full = a_empty + OZ_A_STAGES;
    uint64_t *b_empty = b_full + 2;
    uint64_t *acc_full = b_empty + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * OZ_BM, n0 = blockIdx.x * OZ_BN;
    const int k_begin = blockIdx.z * p.k_per_split;
    const int k_end = min(p.K, k_begin + p.k_per_split);
    const int nkb = k_end > k_begin ? (k_end - k_begin + 127) / 128 : 0;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < OZ_A_STAGES; ++s) {
            oz_mbar_init(&a_full[s], 1);
            oz_mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            oz_mbar_init(&b_full[s], 1);
            oz_mbar_init(&b_empty[s], 1);
        }
        oz_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_ptr)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ============ TMA producer ============
        int a_it = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int k0 = k_begin + kb * 128;
            const int bs = kb & 1;
            oz_wait(&b_empty[bs], ((kb >> 1) & 1) ^ 1);
            oz_expect_tx(&b_full[bs], OZ_S * OZ_B_BYTES);
            for (int q = 0; q < OZ_S; ++q) oz_tma_3d(sB + (bs * OZ_S + q) * OZ_B_BYTES, &tmB, k0, n0, q, &b_full[bs]);
            for (int pp = 0; pp < OZ_S; ++pp, ++a_it) {
                const int s = a_it % OZ_A_STAGES;
                oz_wait(&a_empty[s], ((a_it / OZ_A_STAGES) & 1) ^ 1);
                oz_expect_tx(&a_full[s], OZ_A_BYTES);
                oz_tma_3d(sA + s * OZ_A_BYTES, &tmA, k0, m0, pp, &a_full[s]);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ============ MMA issuer ============
        // instruction descriptor: D = S32 (2), A = B = signed INT8 (1), both K-major, N>>3, M>>4
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
        uint32_t started = 0;   // bit d: accumulator d already holds data
        int a_it = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int bs = kb & 1;
            oz_wait(&b_full[bs], (kb >> 1) & 1);
            for (int pp = 0; pp < OZ_S; ++pp, ++a_it) {
                const int s = a_it % OZ_A_STAGES;
                oz_wait(&a_full[s], (a_it / OZ_A_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = oz_smem_u32(sA + s * OZ_A_BYTES);
                for (int q = 0; q + pp < OZ_S; ++q) {
                    const uint32_t b = oz_smem_u32(sB + (bs * OZ_S + q) * OZ_B_BYTES);
                    const int d = pp + q;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 4 x K32 per 128-byte row, +32 bytes each
                        oz_umma_i8(tmem_base + (uint32_t)(d * OZ_BN), oz_desc(a + k * 32), oz_desc(b + k * 32), idesc,
                                   ((started >> d) & 1u) | (k != 0 ? 1u : 0u));
                    }
                    started |= 1u << d;
                }
                oz_commit(&a_empty[s]);
            }
            oz_commit(&b_empty[bs]);
        }
        if (nkb > 0) oz_commit(acc_full);
    } else if (warp >= 2) {
        // ============ epilogue ============
        const int lane_group = warp & 3;
        const int half = (warp - 2) >> 2;             // 0: columns 0..31, 1: columns 32..63
        const int row = m0 + lane_group * 32 + lane;
        if (nkb > 0) {
            oz_wait(acc_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const double ea = row < p.M ? p.scale_a[row] : 0.0;
        double *C = p.C + (int64_t)blockIdx.z * p.split_stride;
#pragma unroll 1
        for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 8) {
            double acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.0;
            if (nkb > 0) {
#pragma unroll
                for (int d = 0; d < OZ_S; ++d) {
                    int32_t v[8];
                    oz_tmem_ld8(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(d * OZ_BN + c0), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double w = oz_weight(d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += (double)v[j] * w;   // exact: |v| < 2^31, w a power of two
                }
            }
            if (row < p.M) {
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int col = n0 + c0 + j;
                    if (col >= p.N) break;
                    double z = acc[j] * ea * p.scale_b[col];
                    if (p.bias) z += p.bias[col];
                    const int64_t ci = (int64_t)row * p.ldc + col;
                    const int64_t xi = (int64_t)row * p.ldaux + col;
                    switch (p.epi) {
                        case OZ_EPI_SOFTPLUS: {
                            const double e = exp(z), dd = 1.0 + e;
                            double r = log(dd);
                            if (isinf(r)) r = z;
                            if (p.colmask) r *= p.colmask[col];
                            C[ci] = r;
                            if (p.aux_out) p.aux_out[xi] = isinf(e) ? 1.0 : e / dd;
                            break;
                        }
                        case OZ_EPI_TANH: {
                            double r = tanh(z);
                            if (p.colmask) r *= p.colmask[col];
                            C[ci] = r;
                            break;
                        }
                        case OZ_EPI_GAUSSIAN: {
                            double a_ = exp(-(z * z / p.tparam));
                            if (p.colmask) a_ *= p.colmask[col];
                            C[ci] = a_;
                            if (p.aux_out) p.aux_out[xi] = -2.0 * z * a_ / p.tparam;
                            break;
                        }
                        default:
                            C[ci] = p.colmask ? z * p.colmask[col] : z;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// ------------------------------------------------------------------------------------------
// slicing kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double pow2_ceil_of(double amax) {
    // smallest power of two >= amax (amax > 0), so that |x| / scale <= 1
    int e;
    const double m = frexp(amax, &e);       // amax = m * 2^e, m in [0.5, 1)
    return ldexp(1.0, m == 0.5 ? e - 1 : e);
}

// One warp per row of a row-major [rows x cols] FP64 matrix: scale[row] = 2^ceil(log2 max|x|), then S
// exact 7-bit truncation slices written K-major as INT8 [S][rows_pad][cols_pad] (zero padded).
__global__ void __launch_bounds__(256) slice_rows_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                                         int8_t *__restrict__ out, int64_t rows_pad, int64_t cols_pad,
                                                         double *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const double *x = src + row * ld;
    double amax = 0.0;
    for (int64_t c = lane; c < cols; c += 32) amax = fmax(amax, fabs(x[c]));
    amax = warp_max(amax);
    const double sc = amax > 0.0 && isfinite(amax) ? pow2_ceil_of(amax) : 1.0;
    if (lane == 0) scale[row] = sc;
    const double inv = 1.0 / sc;            // exact (power of two)
    for (int64_t c = lane; c < cols_pad; c += 32) {
        double r = c < cols ? x[c] * inv : 0.0;
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) {
            r *= 128.0;
            double t = trunc(r);
            t = fmin(fmax(t, -127.0), 127.0);
            r -= t;
            out[((int64_t)s * rows_pad + row) * cols_pad + c] = (int8_t)(int)t;
        }
    }
}

// column-wise max |x| of a row-major [rows x cols] matrix, two stages (partial[blockIdx.y][col])
__global__ void colabsmax_partial_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                         int64_t rows_per_block, double *__restrict__ partial) {
    __shared__ double tile[8][33];
    const int64_t col = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    double m = 0.0;
    if (col < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) m = fmax(m, fabs(src[r * ld + col]));
    tile[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0 && col < cols) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t = fmax(t, tile[k][threadIdx.x]);
        partial[blockIdx.y * cols + col] = t;
    }
}
__global__ void colabsmax_final_kernel(const double *__restrict__ partial, int nparts, int64_t cols, double *__restrict__ scale) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col >= cols) return;
    double m = 0.0;
    for (int k = 0; k < nparts; ++k) m = fmax(m, partial[k * cols + col]);
    scale[col] = m > 0.0 && isfinite(m) ? pow2_ceil_of(m) : 1.0;
}

// Slices of the TRANSPOSE: src row-major [rows x cols], per-COLUMN scales; out INT8 [S][cols_pad][rows_pad]
// (each output row = one column of src, contiguous along src's rows = the contraction index).
__global__ void __launch_bounds__(256) slice_cols_transposed_kernel(const double *__restrict__ src, int64_t rows, int64_t cols,
                                                                    int64_t ld, const double *__restrict__ scale,
                                                                    int8_t *__restrict__ out, int64_t cols_pad, int64_t rows_pad) {
    __shared__ int8_t tile[OZ_S][32][36];    // [slice][col][row], padded
    const int64_t r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t r = r0 + rr, c = c0 + tx;
        double v = (r < rows && c < cols) ? src[r * ld + c] / scale[c] : 0.0;
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) {
            v *= 128.0;
            double t = trunc(v);
            t = fmin(fmax(t, -127.0), 127.0);
            v -= t;
            tile[s][tx][rr] = (int8_t)(int)t;
        }
    }
    __syncthreads();
    // write: for each slice and column, 32 consecutive bytes along rows
    for (int idx = threadIdx.x; idx < OZ_S * 32 * 8; idx += blockDim.x) {
        const int s = idx / (32 * 8), rem = idx % (32 * 8), cc = rem / 8, seg = rem % 8;   // 8 segments of 4 bytes
        const int64_t c = c0 + cc, r = r0 + seg * 4;
        if (c < cols_pad && r < rows_pad)
            *reinterpret_cast<int32_t *>(out + ((int64_t)s * cols_pad + c) * rows_pad + r) =
                *reinterpret_cast<const int32_t *>(&tile[s][cc][seg * 4]);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 oz_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 3-D INT8 tensor map over slices [S][rows_pad][k_pad] (k contiguous), box {128 k, box_rows, 1 slice}, 128B swizzle
int oz_tensor_map(CUtensorMap *map, const int8_t *base, int64_t rows_pad, int64_t k_pad, uint32_t box_rows) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = oz_encode_fn();
    if (!fn) return fail(OPL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)rows_pad, (cuuint64_t)OZ_S};
    cuuint64_t strides[2] = {(cuuint64_t)k_pad, (cuuint64_t)k_pad * (cuuint64_t)rows_pad};
    cuuint32_t box[3] = {128, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OPL_E_CUDA, "cuTensorMapEncodeTiled (int8 slices) failed (%d)", (int)r);
    return OPL_OK;
}

int oz_grid1(int64_t total, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), 16 * (int64_t)sm_count()));
}

}  // namespace

int64_t ozaki_pad(int64_t v) { return ceil_div(v, 128) * 128; }

int ozaki_slice_rows(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                     cudaStream_t stream, int64_t *launches) {
    const int64_t rp = ozaki_pad(rows), cp = ozaki_pad(cols);
    OPL_CUDA(cudaMemsetAsync(slices, 0, (size_t)OZ_S * rp * cp, stream));     // padding rows must be zero
    slice_rows_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, stream>>>(src, rows, cols, ld, slices, rp, cp, scale);
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

int ozaki_slice_cols_transposed(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                                double *scratch, int64_t scratch_len, cudaStream_t stream, int64_t *launches) {
    const int64_t cp = ozaki_pad(cols), rp = ozaki_pad(rows);
    int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 256), scratch_len / std::max<int64_t>(cols, 1)));
    nblk = std::min<int64_t>(nblk, 4 * (int64_t)sm_count());
    const int64_t rpb = ceil_div(rows, nblk);
    nblk = ceil_div(rows, rpb);
    dim3 g1((unsigned)ceil_div(cols, 32), (unsigned)nblk);
    colabsmax_partial_kernel<<<g1, dim3(32, 8), 0, stream>>>(src, rows, cols, ld, rpb, scratch);
    colabsmax_final_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, stream>>>(scratch, (int)nblk, cols, scale);
    dim3 g2((unsigned)ceil_div(rp, 32), (unsigned)ceil_div(cp, 32));
    slice_cols_transposed_kernel<<<g2, 256, 0, stream>>>(src, rows, cols, ld, scale, slices, cp, rp);
    OPL_LAUNCH_CHECK();
    if (launches) *launches += 3;
    return OPL_OK;
}

int launch_ozaki_gemm(const int8_t *a_slices, int64_t a_rows_pad, const int8_t *b_slices, int64_t b_rows_pad, int64_t k_pad,
                      const OzakiArgs &args, int splits, cudaStream_t stream, int64_t *launches) {
    if (args.M <= 0 || args.N <= 0) return OPL_OK;
    if (args.k_per_split % 128 || args.k_per_split > OZ_MAX_K)
        return fail(OPL_E_ARG, "ozaki GEMM: k_per_split must be a multiple of 128 and <= %d", OZ_MAX_K);
    CUtensorMap ta, tb;
    int rc = oz_tensor_map(&ta, a_slices, a_rows_pad, k_pad, OZ_BM);
    if (rc == OPL_OK) rc = oz_tensor_map(&tb, b_slices, b_rows_pad, k_pad, OZ_BN);
    if (rc != OPL_OK) return rc;
    static bool configured = false;
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(args.N, OZ_BN), (unsigned)ceil_div(args.M, OZ_BM), (unsigned)splits);
    ozaki_gemm_kernel<<<grid, OZ_THREADS, OZ_SMEM, stream>>>(ta, tb, args);
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

__global__ void oz_fold_kernel(const double *__restrict__ partial, int splits, int64_t stride, int64_t count, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * stride + i];
        out[i] = s;
    }
}

}  // namespace opl

// Validation / benchmark hook: C[M,N] = A[M,K] . B[K,N] (row-major FP64 device buffers) through the Ozaki path
// (slicing included in `slice_ms_out`, the INT8 GEMM + fold in `gemm_ms_out`).
extern "C" int opl_bench_gemm_ozaki(int32_t M, int32_t N, int32_t K, const double *a_dev, const double *b_dev, double *c_dev,
                                    int32_t iters, float *slice_ms_out, float *gemm_ms_out) {
    using namespace opl;
    if (!a_dev || !b_dev || !c_dev || !slice_ms_out || !gemm_ms_out || iters < 1 || M < 1 || N < 1 || K < 1)
        return fail(OPL_E_ARG, "opl_bench_gemm_ozaki: bad arguments");
    const int64_t mp = ozaki_pad(M), np = ozaki_pad(N), kp = ozaki_pad(K);
    int8_t *sa = nullptr, *sb = nullptr;
    double *ea = nullptr, *eb = nullptr, *scratch = nullptr, *partial = nullptr;
    const int64_t scratch_len = 4 * (int64_t)sm_count() * std::max<int64_t>(N, 1);
    int splits = (int)ceil_div(kp, OZ_MAX_K);
    const int k_per_split = (int)(ceil_div(ceil_div(kp, splits), 128) * 128);
    splits = (int)ceil_div(kp, k_per_split);
    OPL_CUDA(cudaMalloc(&sa, (size_t)OZ_S * mp * kp));
    OPL_CUDA(cudaMalloc(&sb, (size_t)OZ_S * np * kp));
    OPL_CUDA(cudaMalloc(&ea, sizeof(double) * mp));
    OPL_CUDA(cudaMalloc(&eb, sizeof(double) * np));
    OPL_CUDA(cudaMalloc(&scratch, sizeof(double) * scratch_len));
    if (splits > 1) OPL_CUDA(cudaMalloc(&partial, sizeof(double) * (size_t)splits * M * N));
    OPL_CUDA(cudaMemset(sb, 0, (size_t)OZ_S * np * kp));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    int rc = OPL_OK;
    float slice_ms = 0.f, gemm_ms = 0.f;
    for (int it = 0; it <= iters && rc == OPL_OK; ++it) {
        cudaEventRecord(e0, nullptr);
        rc = ozaki_slice_rows(a_dev, M, K, K, sa, ea, nullptr, nullptr);
        if (rc == OPL_OK) rc = ozaki_slice_cols_transposed(b_dev, K, N, N, sb, eb, scratch, scratch_len, nullptr, nullptr);
        cudaEventRecord(e1, nullptr);
        OzakiArgs g = {};
        g.C = splits > 1 ? partial : c_dev;
        g.ldc = N;
        g.ldaux = N;
        g.M = M;
        g.N = N;
        g.K = (int)kp;
        g.k_per_split = k_per_split;
        g.split_stride = splits > 1 ? (int64_t)M * N : 0;
        g.scale_a = ea;
        g.scale_b = eb;
        g.epi = OZ_EPI_STORE;
        if (rc == OPL_OK) rc = launch_ozaki_gemm(sa, mp, sb, np, kp, g, splits, nullptr, nullptr);
        if (rc == OPL_OK && splits > 1)
            oz_fold_kernel<<<oz_grid1((int64_t)M * N, 256), 256>>>(partial, splits, (int64_t)M * N, (int64_t)M * N, c_dev);
        cudaEventRecord(e2, nullptr);
        cudaEventSynchronize(e2);
        if (it > 0) {
            float t1 = 0.f, t2 = 0.f;
            cudaEventElapsedTime(&t1, e0, e1);
            cudaEventElapsedTime(&t2, e1, e2);
            slice_ms += t1;
            gemm_ms += t2;
        }
    }
    *slice_ms_out = slice_ms / iters;
    *gemm_ms_out = gemm_ms / iters;
    cudaFree(sa);
    cudaFree(sb);
    cudaFree(ea);
    cudaFree(eb);
    cudaFree(scratch);
    cudaFree(partial);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    if (rc == OPL_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(OPL_E_CUDA, "ozaki bench: %s", cudaGetErrorString(e));
    }
    return rc;
}
