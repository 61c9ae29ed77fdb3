// tcgen05 / TMEM / TMA GEMM for the fast mode -- see gemm_tf32.cuh for the design.
#include <cudaTypedefs.h>

#include <algorithm>

#include "gemm_tf32.cuh"

namespace opl {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], 128 x N x 8 in TF32 with FP32 accumulation
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, Blackwell version 1 (cute::UMMA::SmemDescriptor).
//   layout_type 2 = SWIZZLE_128B          (K-major operands: 16-byte chunks XOR row%8, 1024-byte atoms)
//   layout_type 1 = SWIZZLE_128B_BASE32B  (the ONLY layout for MN-major 32-bit operands: 32-byte chunks
//                                          XOR row%4, 512-byte atoms; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
constexpr int tf32_stages(int bn) { return bn >= 128 ? 4 : 8; }

template <int BN>
struct Tf32Smem {
    static constexpr int STAGES = tf32_stages(BN);
    static constexpr int A_BYTES = TF32_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int PATCH_BYTES = 8 * 4096;   // per epilogue warp: 32 rows x 32 floats, 16-byte chunks swizzled
    static constexpr int BYTES = STAGES * (A_BYTES + B_BYTES) + PATCH_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};
constexpr int TF32_THREADS = 320;      // TMA producer, MMA issuer, 8 epilogue warps

// 32 rows x 32 columns held row-per-thread -> the warp's shared patch -> four full 128-byte row segments per store
// instruction (a row-per-thread float4 store touches 32 different lines per instruction and makes the LSU the bottleneck)
__device__ __forceinline__ void tf32_store_patch(float *patch, const float (&v)[32], float *base, int64_t ld, int row0, int col0,
                                                 int M, int lane) {
    float4 *pw = reinterpret_cast<float4 *>(patch + lane * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) pw[c ^ (lane & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    const int chunk = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int prow = (lane >> 3) + 4 * i, grow = row0 + prow;
        const float4 q = *reinterpret_cast<const float4 *>(patch + prow * 32 + 4 * (chunk ^ (prow & 7)));
        if (grow < M) *reinterpret_cast<float4 *>(base + (int64_t)grow * ld + col0 + 4 * chunk) = q;
    }
    __syncwarp();
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// NC columns at a time with the switch OUTSIDE the element loops and every stage of the transcendental chain written as
// its own loop over the NC elements, so NC independent chains sit in one basic block and interleave.  (Round 1 did this
// for four elements inside a loop over groups; the switch between the groups kept the compiler from interleaving more
// than four chains, and with two epilogue warps per scheduler the MUFU -> FMUL -> MUFU chain latency, not the issue rate,
// bounded the forward GEMM's epilogue.)
template <int NC>
__device__ __forceinline__ void epi_apply(const Tf32GemmArgs &p, const float (&z)[NC], const float (&aux)[NC], float (&r)[NC],
                                          float (&ao)[NC]) {
#pragma unroll
    for (int e = 0; e < NC; ++e) ao[e] = 0.f;
    switch (p.epi) {
        case TEPI_TANH:
#pragma unroll
            for (int e = 0; e < NC; ++e) r[e] = tanhf(z[e]);
            break;
        case TEPI_SOFTPLUS: {
            // softplus = max(z, 0) + log(1 + t), logistic = z >= 0 ? 1/(1 + t) : t/(1 + t) with t = exp(-|z|): three MUFU
            // operations (EX2, RCP, LG2) and ~10 FP32 instructions per element instead of ~65 for expf + log1pf + an
            // IEEE division.  The approximations are good to 2^-21 relative, the stored results are rounded to TF32
            // (2^-11).  NaN and +-inf come out as in the library version (softplus(inf) = inf, logistic(-inf) = 0).
            float t[NC], u[NC], ru[NC];
#pragma unroll
            for (int e = 0; e < NC; ++e) t[e] = __expf(-fabsf(z[e]));
#pragma unroll
            for (int e = 0; e < NC; ++e) u[e] = 1.0f + t[e];
#pragma unroll
            for (int e = 0; e < NC; ++e) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ru[e]) : "f"(u[e]));
#pragma unroll
            for (int e = 0; e < NC; ++e) ao[e] = z[e] >= 0.f ? ru[e] : t[e] * ru[e];
#pragma unroll
            for (int e = 0; e < NC; ++e) u[e] = __log2f(u[e]) * 0.6931471805599453f;
#pragma unroll
            for (int e = 0; e < NC; ++e) {
                // log(1 + t): LG2 has an absolute error of 2^-22, so below t = 2^-8 the series t - t^2/2 takes over
                const float l = t[e] < 0.00390625f ? t[e] * (1.0f - 0.5f * t[e]) : u[e];
                r[e] = fmaxf(z[e], 0.f) + l;
            }
            break;
        }
        case TEPI_GAUSSIAN:
#pragma unroll
            for (int e = 0; e < NC; ++e) {
                r[e] = expf(-(z[e] * z[e] / p.tparam));
                ao[e] = -2.0f * z[e] * r[e] / p.tparam;
            }
            break;
        case TEPI_MUL_AUX:
#pragma unroll
            for (int e = 0; e < NC; ++e) r[e] = z[e] * aux[e];
            break;
        case TEPI_MUL_DTANH:
#pragma unroll
            for (int e = 0; e < NC; ++e) r[e] = z[e] * (1.0f - aux[e] * aux[e]);
            break;
        default:
#pragma unroll
            for (int e = 0; e < NC; ++e) r[e] = z[e];
    }
}

struct Tf32Tile {
    int m0, n0, k_begin, nkb, split;
};
__device__ __forceinline__ Tf32Tile tf32_tile(const Tf32GemmArgs &p, int t, int tiles_m, int tiles_n, int bn) {
    Tf32Tile r;
    const int per_split = tiles_m * tiles_n;
    r.split = t / per_split;
    const int u = t - r.split * per_split;
    int m, n;
    if (tiles_n >= 8 && tiles_m >= 8) {
        // grouped raster for square-ish problems: the CTAs running at the same time cover 8 rows of tiles x ~18 columns
        // instead of ~2 rows x all columns, so the operand set a wave touches fits the L2
        const int group = u / (8 * tiles_n), within = u - group * 8 * tiles_n;
        const int gm = min(8, tiles_m - group * 8);
        m = group * 8 + within % gm;
        n = within / gm;
    } else {
        n = u % tiles_n;
        m = u / tiles_n;
    }
    r.m0 = m * TF32_BM;
    r.n0 = n * bn;
    r.k_begin = r.split * p.k_per_split;
    const int k_end = min(p.K, r.k_begin + p.k_per_split);
    r.nkb = k_end > r.k_begin ? (k_end - r.k_begin + TF32_BK - 1) / TF32_BK : 0;
    return r;
}

// Persistent: one CTA per SM walks tiles (n fastest) with a static schedule; the accumulator is double buffered in
// TMEM (2 x BN columns), so the 8 epilogue warps drain / transform / store tile t while the MMA warp is already
// accumulating tile t+1, and the TMA producer runs ahead across tile boundaries through a 6-8 stage ring.
template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(TF32_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Tf32GemmArgs p, int splits) {
    using S = Tf32Smem<BN>;
    constexpr int STAGES = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * S::A_BYTES;
    float *sPatch = reinterpret_cast<float *>(sB + STAGES * S::B_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + STAGES * S::B_BYTES + S::PATCH_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;       // [2]
    uint64_t *acc_empty = acc_full + 2;        // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + TF32_BM - 1) / TF32_BM, tiles_n = (p.N + BN - 1) / BN;
    const int ntiles = tiles_m * tiles_n * splits;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const Tf32Tile tl = tf32_tile(p, t, tiles_m, tiles_n, BN);
            for (int kb = 0; kb < tl.nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t phase = (it / STAGES) & 1;
                mbar_wait(&empty[s], phase ^ 1);
                mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
                const int k0 = tl.k_begin + kb * TF32_BK;
                uint8_t *a = sA + s * S::A_BYTES, *b = sB + s * S::B_BYTES;
                if (!A_MN) tma_load_2d(a, &tmA, k0, tl.m0, &full[s]);                       // box {32 k, 128 m}
                else
                    for (int j = 0; j < TF32_BM / 32; ++j) tma_load_2d(a + j * 4096, &tmA, tl.m0 + 32 * j, k0, &full[s]);   // box {32 m, 32 k}
                if (!B_MN) tma_load_2d(b, &tmB, k0, tl.n0, &full[s]);                       // box {32 k, BN n}
                else
                    for (int j = 0; j < BN / 32; ++j) tma_load_2d(b + j * 4096, &tmB, tl.n0 + 32 * j, k0, &full[s]);        // box {32 n, 32 k}
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, majors, N>>3, M>>4
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TF32_BM >> 4) << 24);
        uint32_t it = 0, local = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const Tf32Tile tl = tf32_tile(p, t, tiles_m, tiles_n, BN);
            if (tl.nkb == 0) continue;                 // (the epilogue writes zeros without a handshake)
            const uint32_t buf = local & 1;
            mbar_wait(&acc_empty[buf], ((local >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BN;
            for (int kb = 0; kb < tl.nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t phase = (it / STAGES) & 1;
                mbar_wait(&full[s], phase);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + s * S::A_BYTES), b = smem_u32(sB + s * S::B_BYTES);
#pragma unroll
                for (int k = 0; k < TF32_BK / 8; ++k) {
                    // K-major: 8 tf32 = 32 bytes further along the swizzled 128-byte row; SBO = 8 rows x 128 B
                    // MN-major: 8 k-rows = 1024 bytes further (two 4-row atoms, SBO = 512 B);
                    //           LBO = next 32-wide MN block (one TMA box of 32 k-rows = 4096 B)
                    const uint64_t ad = A_MN ? smem_desc(a + k * 1024, 4096, 512, 1) : smem_desc(a + k * 32, 16, 1024, 2);
                    const uint64_t bd = B_MN ? smem_desc(b + k * 1024, 4096, 512, 1) : smem_desc(b + k * 32, 16, 1024, 2);
                    umma_tf32(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty[s]);       // smem stage reusable once these MMAs have read it
            }
            umma_commit(&acc_full[buf]);
            ++local;
        }
    } else if (warp >= 2) {
        // ================= epilogue: TMEM -> registers -> global =================
        const int lane_group = warp & 3;                 // tcgen05.ld: warp w may touch lanes 32*(w%4)..+31
        const int half = (warp - 2) >> 2;                // columns [half * BN/2, (half+1) * BN/2) of the tile
        constexpr int CH = BN / 2;                       // columns per thread
        constexpr int CW = CH >= 32 ? 32 : 16;           // columns per TMEM load
        uint32_t local = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const Tf32Tile tl = tf32_tile(p, t, tiles_m, tiles_n, BN);
            const int row = tl.m0 + lane_group * 32 + lane;
            const uint32_t buf = local & 1;
            if (tl.nkb > 0) {
                mbar_wait(&acc_full[buf], (local >> 1) & 1);
                tc_fence_after();
            }
            float *C = p.C + (int64_t)tl.split * p.split_stride;
            const bool need_aux_in = p.epi == TEPI_MUL_AUX || p.epi == TEPI_MUL_DTANH;
            const bool have_aux_out = p.aux_out != nullptr && (p.epi == TEPI_SOFTPLUS || p.epi == TEPI_GAUSSIAN);
#pragma unroll 1
            for (int c0 = half * CH; c0 < half * CH + CH; c0 += CW) {
                uint32_t v[CW];
                if (tl.nkb > 0) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(buf * BN + c0);
                    if constexpr (CW == 32) tmem_ld32(taddr, v);
                    else tmem_ld16(taddr, v);
                } else {
#pragma unroll
                    for (int j = 0; j < CW; ++j) v[j] = 0u;
                }
                if (c0 + CW >= half * CH + CH && tl.nkb > 0) {     // last read of this buffer by this warp: release it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                const int col0 = tl.n0 + c0;
                if (CW == 32 ? col0 < p.N : (row < p.M && col0 < p.N)) {     // (the patch path needs the whole warp)
                    float out[CW], aux_o[CW];
                    constexpr int NC = 16;                       // columns whose transfer-function chains interleave
#pragma unroll
                    for (int g0 = 0; g0 < CW; g0 += NC) {
                        float zq[NC], axv[NC], rq[NC], aq[NC];
#pragma unroll
                        for (int q = 0; q < NC / 4; ++q) {
                            float4 ax = make_float4(0.f, 0.f, 0.f, 0.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (need_aux_in && row < p.M)
                                ax = *reinterpret_cast<const float4 *>(p.aux_in + (int64_t)row * p.ldaux + col0 + g0 + 4 * q);
                            if (p.bias) bs = *reinterpret_cast<const float4 *>(p.bias + col0 + g0 + 4 * q);
                            axv[4 * q] = ax.x;
                            axv[4 * q + 1] = ax.y;
                            axv[4 * q + 2] = ax.z;
                            axv[4 * q + 3] = ax.w;
                            zq[4 * q] = __uint_as_float(v[g0 + 4 * q]) + bs.x;
                            zq[4 * q + 1] = __uint_as_float(v[g0 + 4 * q + 1]) + bs.y;
                            zq[4 * q + 2] = __uint_as_float(v[g0 + 4 * q + 2]) + bs.z;
                            zq[4 * q + 3] = __uint_as_float(v[g0 + 4 * q + 3]) + bs.w;
                        }
                        epi_apply<NC>(p, zq, axv, rq, aq);
#pragma unroll
                        for (int e = 0; e < NC; ++e) {
                            const int j = g0 + e;
                            float r = p.round_out ? round_tf32(rq[e]) : rq[e];
                            float ao = aq[e];
                            if (col0 + j >= p.n_valid) {      // layer padding stays exactly zero
                                r = 0.f;
                                ao = 0.f;
                            }
                            out[j] = r;
                            aux_o[j] = ao;
                        }
                    }
                    if constexpr (CW == 32) {
                        float *patch = sPatch + (warp - 2) * 1024;
                        const int row0 = tl.m0 + lane_group * 32;
                        tf32_store_patch(patch, out, C, p.ldc, row0, col0, p.M, lane);
                        if (have_aux_out) tf32_store_patch(patch, aux_o, p.aux_out, p.ldaux, row0, col0, p.M, lane);
                    } else {
#pragma unroll
                        for (int q = 0; q < CW / 4; ++q) {
                            *reinterpret_cast<float4 *>(C + (int64_t)row * p.ldc + col0 + 4 * q) =
                                make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
                            if (have_aux_out)
                                *reinterpret_cast<float4 *>(p.aux_out + (int64_t)row * p.ldaux + col0 + 4 * q) =
                                    make_float4(aux_o[4 * q], aux_o[4 * q + 1], aux_o[4 * q + 2], aux_o[4 * q + 3]);
                        }
                    }
                }
            }
            if (tl.nkb > 0) ++local;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

int make_tensor_map_f32(CUtensorMap *map, const float *base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                        uint32_t box_inner, uint32_t box_outer, bool mn_major) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
    if (!fn) return fail(OPL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld_elems * sizeof(float)) % 16)
        return fail(OPL_E_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OPL_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return OPL_OK;
}

namespace {

template <bool A_MN, bool B_MN, int BN>
int launch_tf32(const CUtensorMap &ta, const CUtensorMap &tb, const Tf32GemmArgs &a, int splits, cudaStream_t stream) {
    auto kern = gemm_tf32_kernel<A_MN, B_MN, BN>;
    static bool configured_dev[64] = {};     // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tf32Smem<BN>::BYTES));
        configured = true;
    }
    const int64_t tiles = ceil_div(a.N, BN) * ceil_div(a.M, TF32_BM) * splits;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, sm_count());
    kern<<<grid, TF32_THREADS, Tf32Smem<BN>::BYTES, stream>>>(ta, tb, a, splits);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

template <bool A_MN, bool B_MN>
int launch_bn(int bn, const CUtensorMap &ta, const CUtensorMap &tb, const Tf32GemmArgs &a, int splits, cudaStream_t stream) {
    if (bn == 32) return launch_tf32<A_MN, B_MN, 32>(ta, tb, a, splits, stream);
    if (bn == 64) return launch_tf32<A_MN, B_MN, 64>(ta, tb, a, splits, stream);
    if (bn == 256) return launch_tf32<A_MN, B_MN, 256>(ta, tb, a, splits, stream);
    return launch_tf32<A_MN, B_MN, 128>(ta, tb, a, splits, stream);
}

}  // namespace

int launch_gemm_tf32(int layout, const float *A, int64_t lda, const float *B, int64_t ldb, const Tf32GemmArgs &args,
                     int splits, cudaStream_t stream, int64_t *launches) {
    if (args.M <= 0 || args.N <= 0) return OPL_OK;
    // K may be ragged only for the TN layout (K = patterns): TMA zero-fills rows beyond the tensor
    if ((layout != 2 && args.K % TF32_BK) || args.k_per_split % TF32_BK || args.N % 32 || args.ldc % 4)
        return fail(OPL_E_ARG, "tf32 GEMM needs N %% 32 == 0 and K %% 32 == 0 (padded layer widths)");
    // 128 x 256 tiles (the whole 512-column TMEM as two accumulators) for wide outputs: 25 % less operand traffic per flop
    const int bn = (args.N >= 512 && args.N % 256 == 0) ? 256 : (args.N >= 128 ? 128 : (args.N >= 64 ? 64 : 32));
    const bool a_mn = layout == 2, b_mn = layout != 1;
    CUtensorMap ta, tb;
    int rc;
    // A: K-major [M rows][K]  -> box {32 k, 128 m};  MN-major [K rows][M] -> box {32 m, 32 k}
    rc = a_mn ? make_tensor_map_f32(&ta, A, (uint64_t)args.M, (uint64_t)args.K, (uint64_t)lda, 32, 32, true)
              : make_tensor_map_f32(&ta, A, (uint64_t)args.K, (uint64_t)args.M, (uint64_t)lda, 32, TF32_BM, false);
    if (rc != OPL_OK) return rc;
    // B: K-major [N rows][K]  -> box {32 k, BN n};   MN-major [K rows][N] -> box {32 n, 32 k}
    rc = b_mn ? make_tensor_map_f32(&tb, B, (uint64_t)args.N, (uint64_t)args.K, (uint64_t)ldb, 32, 32, true)
              : make_tensor_map_f32(&tb, B, (uint64_t)args.K, (uint64_t)args.N, (uint64_t)ldb, 32, (uint32_t)bn, false);
    if (rc != OPL_OK) return rc;
    if (layout == 0) rc = launch_bn<false, true>(bn, ta, tb, args, splits, stream);
    else if (layout == 1) rc = launch_bn<false, false>(bn, ta, tb, args, splits, stream);
    else rc = launch_bn<true, true>(bn, ta, tb, args, splits, stream);
    if (rc == OPL_OK && launches) ++*launches;
    return rc;
}

__global__ void reduce_partials_f32_kernel(const float *__restrict__ partial, int splits, int64_t stride, int64_t count,
                                           float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < splits; ++k) s += (double)partial[(int64_t)k * stride + i];
        out[i] = (float)s;
    }
}

}  // namespace opl

// Benchmark / validation hook: one plain TF32 GEMM on caller-provided device buffers.
extern "C" int opl_bench_gemm_tf32(int32_t layout, int32_t M, int32_t N, int32_t K, const float *a_dev,
                                   const float *b_dev, float *c_dev, float *scratch_dev, int64_t scratch_len,
                                   int32_t splits, int32_t iters, float *ms_out) {
    using namespace opl;
    if (!a_dev || !b_dev || !c_dev || !ms_out || layout < 0 || layout > 2 || iters < 1 || splits < 1)
        return fail(OPL_E_ARG, "opl_bench_gemm_tf32: bad arguments");
    Tf32GemmArgs g = {};
    g.C = c_dev;
    g.ldc = N;
    g.M = M;
    g.N = N;
    g.K = K;
    g.n_valid = N;
    g.ldaux = N;
    g.epi = TEPI_STORE;
    g.k_per_split = (int)(ceil_div(ceil_div(K, splits), TF32_BK) * TF32_BK);
    const int64_t lda = layout == 2 ? M : K, ldb = layout == 1 ? K : N;
    if (splits > 1) {
        if (!scratch_dev || (int64_t)splits * M * N > scratch_len) return fail(OPL_E_ARG, "opl_bench_gemm_tf32: scratch too small");
        g.C = scratch_dev;
        g.split_stride = (int64_t)M * N;
    }
    cudaEvent_t e0, e1;
    OPL_CUDA(cudaEventCreate(&e0));
    OPL_CUDA(cudaEventCreate(&e1));
    const int64_t count = (int64_t)M * N;
    const int rgrid = (int)std::min<int64_t>(ceil_div(count, 256), 4 * (int64_t)sm_count());
    for (int i = 0; i <= iters; ++i) {
        if (i == 1) OPL_CUDA(cudaEventRecord(e0, nullptr));
        int rc = launch_gemm_tf32(layout, a_dev, lda, b_dev, ldb, g, splits, nullptr, nullptr);
        if (rc != OPL_OK) return rc;
        if (splits > 1) reduce_partials_f32_kernel<<<rgrid, 256>>>(scratch_dev, splits, count, count, c_dev);
    }
    OPL_CUDA(cudaEventRecord(e1, nullptr));
    OPL_CUDA(cudaEventSynchronize(e1));
    OPL_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    *ms_out /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}
