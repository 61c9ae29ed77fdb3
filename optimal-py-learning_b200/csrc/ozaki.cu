// FP64-accurate GEMM on the INT8 tensor cores (Ozaki splitting) -- sm_100a, tcgen05.mma.kind::i8.
//
// Why: the FP64 tensor path of B200 (DMMA) peaks at 37 TFLOP/s; the INT8 path offers ~120x the MAC rate.
// How: C = A.B with every ROW of A and every COLUMN of B scaled by a power of two and cut into S = 6
// signed radix-256 digits (an error-free transformation: digit p is floor(r_p + 128/255) of the scaled
// remainder, which keeps every digit in [-128,127] and centres the truncation error, so
// A = 2^ea * sum_p 256^-(p+1) A_p up to 2^-49 of the row scale).  The S(S+1)/2 = 21 products A_p.B_q with
// p+q < S are exact INT32 GEMMs; products with the same p+q share a weight, so ONE TMEM accumulator per
// diagonal d = p+q (6 x 64 INT32 columns = 384 of the 512 TMEM columns) receives them all, and the
// epilogue recombines  C = 2^(ea_row + eb_col) * sum_d 256^-(d+2) acc_d  in FP64 (Horner, smallest first)
// and applies the same fused epilogues as the DMMA family.
//
// INT32 exactness: one product contributes at most K*128^2 per element and a diagonal sums up to 6 of them,
// so a tile never contracts more than OZ_MAX_K = 16384 (6*16384*16384 < 2^31); longer contractions are
// split (`splits`) and folded in FP64 by the caller.
//
// Kernel: persistent, one CTA per SM, static tile schedule (n fastest).  Warp roles:
//   warps 0..7  epilogue.  Drain: TMEM -> FP64 (magic-number int->double, Horner over the 6 diagonals, row scale)
//               -> padded shared staging tile; the accumulators are then released, so the next tile's MMAs overlap
//               the finish phase: column scale, bias, transfer function (FP64 exp/log), coalesced 16-byte stores.
//   (warps 8..11 A loaders, OZ_A_WARPS only)
//   warp n-2    TMA producer: per 128-wide k-block the 6 B digit tiles (64 x 128 B each, one barrier, double
//               buffered), then the 6 A digit tiles (128 x 128 B each) through a 3-deep ring
//   warp n-1    MMA issuer: A_p meets B_0..B_{5-p}; those B tiles are contiguous in shared memory and their
//               accumulators p..5 are contiguous in TMEM, so each A_p needs one tcgen05.mma of N = 64 (6-p)
//               (two when N > 256) per K = 32 step: 32 MMAs per k-block, not 84.  With N >= 128 the SS form stays
//               inside the SM's shared-memory bandwidth (OZ_A_SMEM, the default).  OZ_A_UTCCP / OZ_A_WARPS move
//               the A tile into the spare 128 TMEM columns first (tcgen05.cp / loader warps + tcgen05.st) and
//               read it from there (TS form); kept for comparison.
#include <cudaTypedefs.h>

#include <algorithm>
#include <vector>

#include "ozaki.cuh"

namespace opl {

namespace {

constexpr int OZ_A_BYTES = OZ_BM * 128;        // 16 KB: 128 rows x 128 int8
constexpr int OZ_B_BYTES = OZ_BN * 128;        //  8 KB
constexpr int OZ_A_STAGES = 3;
constexpr int OZ_TSLOTS = 4;                   // A staging slots in TMEM: 32 columns each at columns 384..511
constexpr int OZ_TSLOT_COL0 = OZ_S * OZ_BN;    // 384
constexpr int OZ_ZLD = 66;                     // doubles per staged row (padded: conflict-free 16-byte row-per-thread stores)
constexpr int OZ_Z_BYTES = OZ_BM * OZ_ZLD * 8;
constexpr int OZ_OFF_A = 2 * OZ_S * OZ_B_BYTES;                 // 96 KB of B first (1024-aligned tiles)
constexpr int OZ_OFF_Z = OZ_OFF_A + OZ_A_STAGES * OZ_A_BYTES;
constexpr int OZ_OFF_BAR = OZ_OFF_Z + OZ_Z_BYTES;
constexpr int OZ_SMEM = OZ_OFF_BAR + 256 + 1024;
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_LOADER_WARPS = 4;                             // OZ_A_WARPS only
constexpr int OZ_THREADS_BASE = 32 * (OZ_EPI_WARPS + 2);       // 320
constexpr int OZ_THREADS_LOADERS = 32 * (OZ_EPI_WARPS + OZ_LOADER_WARPS + 2);   // 448
static_assert(OZ_SMEM <= 232448, "shared memory budget");
static_assert(OZ_TSLOT_COL0 + OZ_TSLOTS * 32 <= 512, "TMEM budget");

__device__ __forceinline__ uint32_t oz_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void oz_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void oz_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem, INT32] (+)= A[smem, INT8] . B[smem, INT8], 128 x 64 x 32
__device__ __forceinline__ void oz_umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// same with A read from TMEM (lane = row, 8 columns = 32 int8 along k)
__device__ __forceinline__ void oz_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// shared -> TMEM: 128 lanes x 32 bytes
__device__ __forceinline__ void oz_utccp(uint32_t tmem_dst, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void oz_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
        "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
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
__device__ __forceinline__ uint64_t oz_desc_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }
// one lane of the (converged) warp
__device__ __forceinline__ bool oz_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void oz_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_WARPS * 32) : "memory"); }

// exact int32 -> double without the conversion pipe: bits(2^52 + 2^31 + v) - (2^52 + 2^31)
__device__ __forceinline__ double oz_i2d(int32_t v) {
    return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;
}

struct OzTile {
    int m0, n0, k_begin, nkb, split;
};
__device__ __forceinline__ OzTile oz_tile(const OzakiArgs &p, int t, int tiles_m, int tiles_n) {
    OzTile r;
    const int n = t % tiles_n, rest = t / tiles_n;
    const int m = rest % tiles_m;
    r.split = rest / tiles_m;
    r.m0 = m * OZ_BM;
    r.n0 = n * OZ_BN;
    r.k_begin = r.split * p.k_per_split;
    const int k_end = min(p.K, r.k_begin + p.k_per_split);
    r.nkb = k_end > r.k_begin ? (k_end - r.k_begin + 127) / 128 : 0;
    return r;
}

template <int VARIANT>
__global__ void __launch_bounds__(VARIANT == OZ_A_WARPS ? OZ_THREADS_LOADERS : OZ_THREADS_BASE, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzakiArgs p) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sB = smem;                                   // 2 x 6 x 8 KB
    uint8_t *sA = smem + OZ_OFF_A;                        // OZ_A_STAGES x 16 KB
    double *sZ = reinterpret_cast<double *>(smem + OZ_OFF_Z);
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + OZ_OFF_BAR);
    uint64_t *a_empty = a_full + OZ_A_STAGES;
    uint64_t *b_full = a_empty + OZ_A_STAGES;
    uint64_t *b_empty = b_full + 2;
    uint64_t *t_full = b_empty + 2;
    uint64_t *t_empty = t_full + OZ_TSLOTS;
    uint64_t *acc_full = t_empty + OZ_TSLOTS;
    uint64_t *acc_empty = acc_full + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 1);

    // warp roles: 0..7 epilogue, (8..15 A loaders), second-to-last TMA producer, last MMA issuer (the arbiter favours
    // the highest warp id of a sub-partition, and the single issuing thread is the latency-critical one)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (VARIANT == OZ_A_WARPS ? OZ_THREADS_LOADERS : OZ_THREADS_BASE) / 32;
    const int producer_warp = nwarps - 2, mma_warp = nwarps - 1;
    const int tiles_m = (p.M + OZ_BM - 1) / OZ_BM, tiles_n = (p.N + OZ_BN - 1) / OZ_BN;
    const int ntiles = tiles_m * tiles_n * p.splits;

    if (warp == producer_warp && lane == 0) {
        for (int s = 0; s < OZ_A_STAGES; ++s) {
            oz_mbar_init(&a_full[s], 1);
            oz_mbar_init(&a_empty[s], VARIANT == OZ_A_WARPS ? 4 : 1);
        }
        for (int s = 0; s < 2; ++s) {
            oz_mbar_init(&b_full[s], 1);
            oz_mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < OZ_TSLOTS; ++s) {
            oz_mbar_init(&t_full[s], 4);
            oz_mbar_init(&t_empty[s], 1);
        }
        oz_mbar_init(acc_full, 1);
        oz_mbar_init(acc_empty, OZ_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == mma_warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_ptr)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tmem_base = 0;                   // the CTA owns all 512 columns, so the allocation starts at column 0
    if (*tmem_ptr != 0) __trap();

    if (warp == producer_warp) {
        // ============ TMA producer (warp-uniform control flow, one elected lane issues) ============
        const bool leader = oz_elect_one();
        uint32_t kbt = 0;                           // k-blocks produced so far (all tiles)
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            for (int kb = 0; kb < tl.nkb; ++kb, ++kbt) {
                const int k0 = tl.k_begin + kb * 128;
                const uint32_t bs = kbt & 1;
                oz_wait(&b_empty[bs], ((kbt >> 1) & 1) ^ 1);
                if (leader) {
                    oz_expect_tx(&b_full[bs], OZ_S * OZ_B_BYTES);
#pragma unroll
                    for (int q = 0; q < OZ_S; ++q) oz_tma_3d(sB + (bs * OZ_S + q) * OZ_B_BYTES, &tmB, k0, tl.n0, q, &b_full[bs]);
                }
#pragma unroll
                for (int pp = 0; pp < OZ_S; ++pp) {
                    // a_it = 6 kbt + pp: stage and phase are compile-time functions of pp (6 % OZ_A_STAGES == 0)
                    constexpr int per = OZ_S / OZ_A_STAGES;
                    const int s = pp % OZ_A_STAGES;
                    oz_wait(&a_empty[s], ((kbt * per + pp / OZ_A_STAGES) & 1) ^ 1);
                    if (leader) {
                        oz_expect_tx(&a_full[s], OZ_A_BYTES);
                        oz_tma_3d(sA + s * OZ_A_BYTES, &tmA, k0, tl.m0, pp, &a_full[s]);
                    }
                }
            }
        }
    } else if (warp == mma_warp) {
        // ============ MMA issuer (warp-uniform control flow, one elected lane issues) ============
        // The B digit tiles of one k-block are contiguous 8 KB blocks of 64 rows, and the accumulators of diagonals
        // p..5 are contiguous TMEM columns, so A_p . [B_0; ..; B_{5-p}]^T is ONE MMA of N = 64 (6 - p) columns
        // (cut at N = 256): 8 instructions per K = 32 step instead of 21.  All 512 TMEM columns are ours, so the
        // allocation starts at column 0 and every TMEM address below is a compile-time constant; shared-memory
        // descriptors differ only in their low word (start address >> 4), by compile-time offsets.
        // instruction descriptor: D = S32 (2), A = B = signed INT8 (1), both K-major, N>>3, M>>4
        constexpr uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
        const bool leader = oz_elect_one();
        const uint32_t desc_hi = (uint32_t)(oz_desc(0) >> 32);
        const uint32_t a_lo0 = (uint32_t)oz_desc(oz_smem_u32(sA)), b_lo0 = (uint32_t)oz_desc(oz_smem_u32(sB));
        uint32_t kbt = 0, acc_it = 0;               // acc_it counts tiles that use the accumulators
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            if (tl.nkb == 0) continue;              // (the epilogue skips the accumulator handshake as well)
            oz_wait(acc_empty, (acc_it & 1) ^ 1);
            oz_fence_after();
            for (int kb = 0; kb < tl.nkb; ++kb, ++kbt) {
                const uint32_t bs = kbt & 1;
                const uint32_t b_lo = b_lo0 + bs * (uint32_t)((OZ_S * OZ_B_BYTES) >> 4);
                const uint32_t first = kb > 0 ? 1u : 0u;
                oz_wait(&b_full[bs], (kbt >> 1) & 1);
#pragma unroll
                for (int pp = 0; pp < OZ_S; ++pp) {
                    constexpr int per = OZ_S / OZ_A_STAGES;
                    const int s = pp % OZ_A_STAGES;
                    const uint32_t a_par = (kbt * per + pp / OZ_A_STAGES) & 1;
                    const uint32_t slot = (kbt * OZ_S + pp) % OZ_TSLOTS;      // TMEM staging slot (TS variants)
                    const uint32_t a_lo = a_lo0 + (uint32_t)((s * OZ_A_BYTES) >> 4);
                    const uint32_t ta = (uint32_t)OZ_TSLOT_COL0 + slot * 32;
                    if (VARIANT == OZ_A_WARPS) {
                        oz_wait(&t_full[slot], ((kbt * OZ_S + pp) / OZ_TSLOTS) & 1);
                    } else {
                        oz_wait(&a_full[s], a_par);
                    }
                    oz_fence_after();
                    if (leader) {
                        if (VARIANT == OZ_A_UTCCP) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) oz_utccp(ta + k * 8, oz_desc_join(desc_hi, a_lo + k * 2));
                            oz_commit(&a_empty[s]);            // smem stage reusable once the copies have read it
                        }
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int nq = OZ_S - pp;              // B digits this A digit meets
#pragma unroll
                        for (int q0 = 0; q0 < nq; q0 += 4) {   // chunks of <= 4 digits = N <= 256
                            const int qn = nq - q0 < 4 ? nq - q0 : 4;
                            const uint32_t idesc = idesc0 | ((uint32_t)((qn * OZ_BN) >> 3) << 17);
                            const uint32_t d_tmem = (uint32_t)((pp + q0) * OZ_BN);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {      // 4 x K32 per 128-byte row, +32 bytes each
                                const uint32_t acc = (pp == 0 && k == 0) ? first : 1u;
                                const uint64_t bd = oz_desc_join(desc_hi, b_lo + (uint32_t)((q0 * OZ_B_BYTES + k * 32) >> 4));
                                if (VARIANT == OZ_A_SMEM)
                                    oz_umma_ss(d_tmem, oz_desc_join(desc_hi, a_lo + k * 2), bd, idesc, acc);
                                else
                                    oz_umma_ts(d_tmem, ta + k * 8, bd, idesc, acc);
                            }
                        }
                        if (VARIANT == OZ_A_SMEM) oz_commit(&a_empty[s]);
                        if (VARIANT == OZ_A_WARPS) oz_commit(&t_empty[slot]);
                        if (pp == OZ_S - 1) oz_commit(&b_empty[bs]);
                    }
                    __syncwarp();
                }
            }
            if (leader) oz_commit(acc_full);
            __syncwarp();
            ++acc_it;
        }
    } else if (warp < OZ_EPI_WARPS) {
        // ============ epilogue ============
        const int ew = warp;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int half = ew >> 2;                     // drain: columns [32*half, 32*half+32)
        const int lrow = quarter * 32 + lane;         // drain: one accumulator row per thread
        uint32_t acc_it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            // ---- drain: TMEM -> FP64 -> staging tile ----
            if (tl.nkb > 0) {
                oz_wait(acc_full, acc_it & 1);
                oz_fence_after();
                ++acc_it;
            }
            {
                const int row = tl.m0 + lrow;
                const double ea = (row < p.M ? p.scale_a[row] : 0.0) * (1.0 / 65536.0);
                double *zrow = sZ + lrow * OZ_ZLD + half * 32;
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 8) {
                    double acc[8];
                    if (tl.nkb > 0) {
                        int32_t v[OZ_S][8];
#pragma unroll
                        for (int d = 0; d < OZ_S; ++d)
                            oz_tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * OZ_BN + half * 32 + c0), v[d]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            double s = oz_i2d(v[OZ_S - 1][j]);
#pragma unroll
                            for (int d = OZ_S - 2; d >= 0; --d) s = fma(s, 1.0 / 256.0, oz_i2d(v[d][j]));
                            acc[j] = s * ea;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2 *>(zrow + c0 + j) = make_double2(acc[j], acc[j + 1]);
                }
            }
            if (tl.nkb > 0) {
                oz_fence_before();
                __syncwarp();
                if (lane == 0) oz_arrive(acc_empty);          // accumulators free: the next tile's MMAs may start
            }
            oz_epi_bar();
            // ---- finish: warp ew owns staged rows [16 ew, 16 ew + 16), lane owns columns 2 lane, 2 lane + 1 ----
            {
                const int col = tl.n0 + 2 * lane;
                const bool c0ok = col < p.N, c1ok = col + 1 < p.N;
                const double eb0 = c0ok ? p.scale_b[col] : 0.0, eb1 = c1ok ? p.scale_b[col + 1] : 0.0;
                const double bi0 = (p.bias && c0ok) ? p.bias[col] : 0.0, bi1 = (p.bias && c1ok) ? p.bias[col + 1] : 0.0;
                const double mk0 = (p.colmask && c0ok) ? p.colmask[col] : 1.0, mk1 = (p.colmask && c1ok) ? p.colmask[col + 1] : 1.0;
                double *C = p.C + (int64_t)tl.split * p.split_stride;
                const bool vec_c = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && c1ok;
                const bool vec_x = p.aux_out && ((p.ldaux & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0) && c1ok;
#pragma unroll 2
                for (int r = 0; r < 16; ++r) {
                    const int lr = ew * 16 + r;
                    const int row = tl.m0 + lr;
                    if (row >= p.M) break;
                    const double2 zz = *reinterpret_cast<const double2 *>(sZ + lr * OZ_ZLD + 2 * lane);
                    const double z0 = zz.x * eb0 + bi0, z1 = zz.y * eb1 + bi1;
                    double r0, r1, x0 = 0.0, x1 = 0.0;
                    switch (p.epi) {
                        case OZ_EPI_SOFTPLUS: {
                            const double e0 = exp(z0), e1 = exp(z1);
                            const double d0 = 1.0 + e0, d1 = 1.0 + e1;
                            r0 = log(d0);
                            r1 = log(d1);
                            if (isinf(r0)) r0 = z0;
                            if (isinf(r1)) r1 = z1;
                            x0 = isinf(e0) ? 1.0 : e0 / d0;
                            x1 = isinf(e1) ? 1.0 : e1 / d1;
                            break;
                        }
                        case OZ_EPI_TANH:
                            r0 = tanh(z0);
                            r1 = tanh(z1);
                            break;
                        case OZ_EPI_GAUSSIAN:
                            r0 = exp(-(z0 * z0 / p.tparam));
                            r1 = exp(-(z1 * z1 / p.tparam));
                            x0 = -2.0 * z0 * r0 / p.tparam;
                            x1 = -2.0 * z1 * r1 / p.tparam;
                            break;
                        default:
                            r0 = z0;
                            r1 = z1;
                    }
                    r0 *= mk0;
                    r1 *= mk1;
                    if (p.epi == OZ_EPI_GAUSSIAN) {   // derivative uses the masked output, as the reference does
                        x0 = -2.0 * z0 * r0 / p.tparam;
                        x1 = -2.0 * z1 * r1 / p.tparam;
                    }
                    double *cp = C + (int64_t)row * p.ldc + col;
                    if (vec_c) {
                        *reinterpret_cast<double2 *>(cp) = make_double2(r0, r1);
                    } else {
                        if (c0ok) cp[0] = r0;
                        if (c1ok) cp[1] = r1;
                    }
                    if (p.aux_out && (p.epi == OZ_EPI_SOFTPLUS || p.epi == OZ_EPI_GAUSSIAN)) {
                        double *xp = p.aux_out + (int64_t)row * p.ldaux + col;
                        if (vec_x) {
                            *reinterpret_cast<double2 *>(xp) = make_double2(x0, x1);
                        } else {
                            if (c0ok) xp[0] = x0;
                            if (c1ok) xp[1] = x1;
                        }
                    }
                }
            }
            oz_epi_bar();                                     // staging tile free for the next drain
        }
    } else if (VARIANT == OZ_A_WARPS) {
        // ============ A loaders: shared (128B-swizzled rows) -> registers -> TMEM ============
        const int quarter = warp & 3;
        const int lrow = quarter * 32 + lane;
        uint32_t a_it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            for (int i = 0; i < tl.nkb * OZ_S; ++i, ++a_it) {
                const uint32_t s = a_it % OZ_A_STAGES, slot = a_it % OZ_TSLOTS;
                oz_wait(&a_full[s], (a_it / OZ_A_STAGES) & 1);
                const uint8_t *rowp = sA + s * OZ_A_BYTES + lrow * 128;
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(rowp + ((j ^ (lrow & 7)) << 4));
                    v[4 * j + 0] = q.x;
                    v[4 * j + 1] = q.y;
                    v[4 * j + 2] = q.z;
                    v[4 * j + 3] = q.w;
                }
                __syncwarp();
                if (lane == 0) oz_arrive(&a_empty[s]);        // this warp's 32 rows are in registers
                oz_wait(&t_empty[slot], ((a_it / OZ_TSLOTS) & 1) ^ 1);
                oz_fence_after();
                oz_tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(OZ_TSLOT_COL0 + slot * 32), v);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                oz_fence_before();
                __syncwarp();
                if (lane == 0) oz_arrive(&t_full[slot]);
            }
        }
    }
    oz_fence_before();
    __syncthreads();
    if (warp == mma_warp) {
        oz_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// ------------------------------------------------------------------------------------------
// digit (slicing) kernels
// ------------------------------------------------------------------------------------------
// power-of-two scale with |x| / scale <= 0.495 < 1 - 128/255 (x != 0); NaN poisons the product when x is not finite
__device__ __forceinline__ double oz_scale_of(double amax) {
    if (!(amax > 0.0)) return amax == 0.0 ? 1.0 : amax;   // 0 -> 1, NaN -> NaN
    if (isinf(amax)) return __longlong_as_double(0x7ff8000000000000LL);
    int e;
    const double m = frexp(amax, &e);       // amax = m * 2^e, m in [0.5, 1)
    e += m <= 0.99 ? 1 : 2;
    e = max(-1000, min(e, 1023));
    return ldexp(1.0, e);
}

// signed radix-256 digits of r (|r| <= 0.498): t_s = floor(256 r_s + 128/255) in [-128,127], r_{s+1} = 256 r_s - t_s (exact)
__device__ __forceinline__ void oz_digits(double r, int (&t)[OZ_S]) {
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) {
        r *= 256.0;
        double f = floor(r + 128.0 / 255.0);
        f = fmin(fmax(f, -128.0), 127.0);
        r -= f;
        t[s] = (int)f;
    }
}

// One warp per row of a row-major [rows x cols] FP64 matrix: scale[row], then S digit planes written K-major as
// INT8 [S][rows_pad][cols_pad]; each lane converts 4 consecutive columns and stores them as one 32-bit word.
__global__ void __launch_bounds__(256) slice_rows_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                                         int8_t *__restrict__ out, int64_t rows_pad, int64_t cols_pad,
                                                         double *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const double *x = src + row * ld;
    double amax = 0.0;
    bool bad = false;
    for (int64_t c = lane; c < cols; c += 32) {
        const double v = fabs(x[c]);
        bad |= !(v <= 1.7976931348623157e308);
        amax = fmax(amax, v);
    }
    amax = warp_max(amax);
    bad = __any_sync(0xffffffffu, bad);
    const double sc = bad ? __longlong_as_double(0x7ff8000000000000LL) : oz_scale_of(amax);
    if (lane == 0) scale[row] = sc;
    const double inv = bad ? 0.0 : 1.0 / sc;            // exact (power of two)
    for (int64_t c4 = (int64_t)lane * 4; c4 < cols_pad; c4 += 128) {   // pad columns (contraction index) := 0
        uint32_t w[OZ_S];
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) w[s] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int t[OZ_S];
            oz_digits(c4 + j < cols ? x[c4 + j] * inv : 0.0, t);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) w[s] |= (uint32_t)(t[s] & 0xff) << (8 * j);
        }
#pragma unroll
        for (int s = 0; s < OZ_S; ++s)
            *reinterpret_cast<uint32_t *>(out + ((int64_t)s * rows_pad + row) * cols_pad + c4) = w[s];
    }
}

// column-wise max |x| of a row-major [rows x cols] matrix, two stages (partial[blockIdx.y][col]); NaN/inf propagate as NaN
__global__ void colabsmax_partial_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                         int64_t rows_per_block, double *__restrict__ partial) {
    __shared__ double tile[8][33];
    const int64_t col = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    double m = 0.0;
    bool bad = false;
    if (col < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            const double v = fabs(src[r * ld + col]);
            bad |= !(v <= 1.7976931348623157e308);
            m = fmax(m, v);
        }
    tile[threadIdx.y][threadIdx.x] = bad ? __longlong_as_double(0x7ff8000000000000LL) : m;
    __syncthreads();
    if (threadIdx.y == 0 && col < cols) {
        double t = 0.0;
        bool nan = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double v = tile[k][threadIdx.x];
            nan |= v != v;
            t = fmax(t, v);
        }
        partial[blockIdx.y * cols + col] = nan ? __longlong_as_double(0x7ff8000000000000LL) : t;
    }
}
__global__ void colabsmax_final_kernel(const double *__restrict__ partial, int nparts, int64_t cols, double *__restrict__ scale) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col >= cols) return;
    double m = 0.0;
    bool nan = false;
    for (int k = 0; k < nparts; ++k) {
        const double v = partial[k * cols + col];
        nan |= v != v;
        m = fmax(m, v);
    }
    scale[col] = nan ? __longlong_as_double(0x7ff8000000000000LL) : oz_scale_of(m);
}

// Digits of the TRANSPOSE: src row-major [rows x cols], per-COLUMN scales; out INT8 [S][cols_pad][rows_pad]
// (each output row = one column of src, contiguous along src's rows = the contraction index).
// Block = 128 src rows x 32 src columns; a thread converts 4 consecutive rows of one column into one word per digit
// plane, the block then writes 128-byte output rows.
__global__ void __launch_bounds__(256) slice_cols_transposed_kernel(const double *__restrict__ src, int64_t rows, int64_t cols,
                                                                    int64_t ld, const double *__restrict__ scale,
                                                                    int8_t *__restrict__ out, int64_t cols_pad, int64_t rows_pad,
                                                                    int64_t row_off) {
    __shared__ uint32_t tile[OZ_S][32][33];    // [digit][col][word along rows], padded
    const int64_t r0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    const int64_t c = c0 + tx;
    double inv = 0.0;
    if (c < cols) {
        const double sc = scale[c];
        inv = sc == sc ? 1.0 / sc : 0.0;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int wq = ty * 4 + g;             // word index along rows: rows r0 + 4 wq ..
        uint32_t w[OZ_S];
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) w[s] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r0 + wq * 4 + j;
            int t[OZ_S];
            oz_digits((r < rows && c < cols) ? src[r * ld + c] * inv : 0.0, t);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) w[s] |= (uint32_t)(t[s] & 0xff) << (8 * j);
        }
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) tile[s][tx][wq] = w[s];
    }
    __syncthreads();
    // write: OZ_S x 32 output rows of 128 bytes = 8 x 16 bytes each
    for (int idx = threadIdx.x; idx < OZ_S * 32 * 8; idx += 256) {
        const int s = idx >> 8, cc = (idx >> 3) & 31, seg = idx & 7;
        const int64_t oc = c0 + cc + row_off, r = r0 + seg * 16;
        if (c0 + cc < cols && r < rows_pad) {
            uint4 v;
            v.x = tile[s][cc][seg * 4 + 0];
            v.y = tile[s][cc][seg * 4 + 1];
            v.z = tile[s][cc][seg * 4 + 2];
            v.w = tile[s][cc][seg * 4 + 3];
            *reinterpret_cast<uint4 *>(out + ((int64_t)s * cols_pad + oc) * rows_pad + r) = v;
        }
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

// 3-D INT8 tensor map over digit planes [S][rows_pad][k_pad] (k contiguous), box {128 k, box_rows, 1 plane}, 128B swizzle
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
    if (r != CUDA_SUCCESS) return fail(OPL_E_CUDA, "cuTensorMapEncodeTiled (int8 digits) failed (%d)", (int)r);
    return OPL_OK;
}

int oz_grid1(int64_t total, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), 16 * (int64_t)sm_count()));
}

}  // namespace

int64_t ozaki_pad(int64_t v) { return ceil_div(v, 128) * 128; }

int ozaki_slice_rows(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                     bool zero_fill, cudaStream_t stream, int64_t *launches) {
    const int64_t rp = ozaki_pad(rows), cp = ozaki_pad(cols);
    if (zero_fill) OPL_CUDA(cudaMemsetAsync(slices, 0, (size_t)OZ_S * rp * cp, stream));     // padding must be zero
    slice_rows_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, stream>>>(src, rows, cols, ld, slices, rp, cp, scale);
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

// out[s][0][k] = digit s of 1.0 / 4 for k < rows (a row of ones with scale 4: digit 0 = 64, the rest 0)
__global__ void ones_row_kernel(int8_t *__restrict__ out, int64_t rows, int64_t rows_pad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < rows_pad) out[k] = k < rows ? 64 : 0;
}

int ozaki_ones_row(int8_t *slices, double *scale, int64_t rows, int64_t m_pad, cudaStream_t stream, int64_t *launches) {
    const int64_t rp = ozaki_pad(rows);
    for (int s = 1; s < OZ_S; ++s) OPL_CUDA(cudaMemsetAsync(slices + (size_t)s * m_pad * rp, 0, (size_t)rp, stream));
    ones_row_kernel<<<(unsigned)ceil_div(rp, 256), 256, 0, stream>>>(slices, rows, rp);
    const double four = 4.0;
    OPL_CUDA(cudaMemcpyAsync(scale, &four, sizeof(double), cudaMemcpyHostToDevice, stream));
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

int ozaki_slice_cols_transposed(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                                double *scratch, int64_t scratch_len, int64_t row_off, int64_t m_pad, bool zero_fill,
                                cudaStream_t stream, int64_t *launches) {
    const int64_t cp = m_pad > 0 ? m_pad : ozaki_pad(cols + row_off), rp = ozaki_pad(rows);
    if (zero_fill) OPL_CUDA(cudaMemsetAsync(slices, 0, (size_t)OZ_S * rp * cp, stream));
    int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 256), scratch_len / std::max<int64_t>(cols, 1)));
    nblk = std::min<int64_t>(nblk, 4 * (int64_t)sm_count());
    const int64_t rpb = ceil_div(rows, nblk);
    nblk = ceil_div(rows, rpb);
    dim3 g1((unsigned)ceil_div(cols, 32), (unsigned)nblk);
    colabsmax_partial_kernel<<<g1, dim3(32, 8), 0, stream>>>(src, rows, cols, ld, rpb, scratch);
    colabsmax_final_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, stream>>>(scratch, (int)nblk, cols, scale + row_off);
    dim3 g2((unsigned)ceil_div(rp, 128), (unsigned)ceil_div(cols, 32));
    slice_cols_transposed_kernel<<<g2, 256, 0, stream>>>(src, rows, cols, ld, scale + row_off, slices, cp, rp, row_off);
    OPL_LAUNCH_CHECK();
    if (launches) *launches += 3;
    return OPL_OK;
}

int launch_ozaki_gemm(const int8_t *a_slices, int64_t a_rows_pad, const int8_t *b_slices, int64_t b_rows_pad, int64_t k_pad,
                      const OzakiArgs &args, int variant, cudaStream_t stream, int64_t *launches) {
    if (args.M <= 0 || args.N <= 0) return OPL_OK;
    if (args.k_per_split % 128 || args.k_per_split > OZ_MAX_K || args.splits < 1)
        return fail(OPL_E_ARG, "ozaki GEMM: k_per_split must be a multiple of 128 and <= %d", OZ_MAX_K);
    CUtensorMap ta, tb;
    int rc = oz_tensor_map(&ta, a_slices, a_rows_pad, k_pad, OZ_BM);
    if (rc == OPL_OK) rc = oz_tensor_map(&tb, b_slices, b_rows_pad, k_pad, OZ_BN);
    if (rc != OPL_OK) return rc;
    static bool configured = false;
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<OZ_A_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        OPL_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<OZ_A_UTCCP>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        OPL_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<OZ_A_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        configured = true;
    }
    const int64_t tiles = ceil_div(args.M, OZ_BM) * ceil_div(args.N, OZ_BN) * args.splits;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, sm_count());
    switch (variant) {
        case OZ_A_SMEM:
            ozaki_gemm_kernel<OZ_A_SMEM><<<grid, OZ_THREADS_BASE, OZ_SMEM, stream>>>(ta, tb, args);
            break;
        case OZ_A_UTCCP:
            ozaki_gemm_kernel<OZ_A_UTCCP><<<grid, OZ_THREADS_BASE, OZ_SMEM, stream>>>(ta, tb, args);
            break;
        case OZ_A_WARPS:
            ozaki_gemm_kernel<OZ_A_WARPS><<<grid, OZ_THREADS_LOADERS, OZ_SMEM, stream>>>(ta, tb, args);
            break;
        default:
            return fail(OPL_E_ARG, "ozaki GEMM: unknown variant %d", variant);
    }
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

// Validation / benchmark hook: C[M,N] = epi(A[M,K] . B[K,N]) (row-major FP64 device buffers) through the Ozaki path
// (digit extraction timed in `slice_ms_out`, the INT8 GEMM + fold in `gemm_ms_out`).
extern "C" int opl_bench_gemm_ozaki(int32_t M, int32_t N, int32_t K, const double *a_dev, const double *b_dev, double *c_dev,
                                    int32_t variant, int32_t epi, int32_t iters, float *slice_ms_out, float *gemm_ms_out) {
    using namespace opl;
    if (!a_dev || !b_dev || !c_dev || !slice_ms_out || !gemm_ms_out || iters < 1 || M < 1 || N < 1 || K < 1)
        return fail(OPL_E_ARG, "opl_bench_gemm_ozaki: bad arguments");
    const int64_t mp = ozaki_pad(M), np = ozaki_pad(N), kp = ozaki_pad(K);
    int8_t *sa = nullptr, *sb = nullptr;
    double *ea = nullptr, *eb = nullptr, *scratch = nullptr, *partial = nullptr;
    const int64_t scratch_len = 4 * (int64_t)sm_count() * std::max<int64_t>(N, 1);
    int splits = (int)ceil_div(kp, OZ_MAX_K);
    if (splits > 1) {     // long contractions: enough splits to fill the SMs
        const int64_t tiles = ceil_div(M, OZ_BM) * ceil_div(N, OZ_BN);
        splits = (int)std::max<int64_t>(splits, sm_count() / std::max<int64_t>(tiles, 1));
    }
    const int k_per_split = (int)(ceil_div(ceil_div(kp, splits), 128) * 128);
    splits = (int)ceil_div(kp, k_per_split);
    if (splits > 1 && epi != OZ_EPI_STORE) return fail(OPL_E_ARG, "opl_bench_gemm_ozaki: fused epilogues need K <= %d", OZ_MAX_K);
    OPL_CUDA(cudaMalloc(&sa, (size_t)OZ_S * mp * kp));
    OPL_CUDA(cudaMalloc(&sb, (size_t)OZ_S * np * kp));
    OPL_CUDA(cudaMalloc(&ea, sizeof(double) * mp));
    OPL_CUDA(cudaMalloc(&eb, sizeof(double) * np));
    OPL_CUDA(cudaMalloc(&scratch, sizeof(double) * scratch_len));
    if (splits > 1) OPL_CUDA(cudaMalloc(&partial, sizeof(double) * (size_t)splits * M * N));
    OPL_CUDA(cudaMemset(sa, 0, (size_t)OZ_S * mp * kp));
    OPL_CUDA(cudaMemset(sb, 0, (size_t)OZ_S * np * kp));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    int rc = OPL_OK;
    float slice_ms = 0.f, gemm_ms = 0.f;
    for (int it = 0; it <= iters && rc == OPL_OK; ++it) {
        cudaEventRecord(e0, nullptr);
        rc = ozaki_slice_rows(a_dev, M, K, K, sa, ea, false, nullptr, nullptr);
        if (rc == OPL_OK) rc = ozaki_slice_cols_transposed(b_dev, K, N, N, sb, eb, scratch, scratch_len, 0, np, false, nullptr, nullptr);
        cudaEventRecord(e1, nullptr);
        OzakiArgs g = {};
        g.C = splits > 1 ? partial : c_dev;
        g.ldc = N;
        g.ldaux = N;
        g.M = M;
        g.N = N;
        g.K = (int)kp;
        g.k_per_split = k_per_split;
        g.splits = splits;
        g.split_stride = splits > 1 ? (int64_t)M * N : 0;
        g.scale_a = ea;
        g.scale_b = eb;
        g.epi = epi;
        g.tparam = 1.0;
        if (rc == OPL_OK) rc = launch_ozaki_gemm(sa, mp, sb, np, kp, g, variant, nullptr, nullptr);
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
