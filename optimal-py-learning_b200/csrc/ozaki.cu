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
//   warps 0..15 epilogue (warp w: TMEM lane quarter w % 4, 16 of the 64 columns).  Drain: TMEM -> integer recombination of
//               the 6 diagonals -> FP64 (two exact magic-number conversions, one rounding), row scale -> staged in the 128
//               TMEM columns the accumulators leave free; the accumulators are released right after, so the next tile's
//               MMAs overlap the finish phase: column scale, bias, transfer function (csrc/oz_transcend.cuh), and stores
//               that go through a per-warp shared patch to become 64-byte row segments.
//   warps 16,17 MMA issuers taking k-blocks in turn: A_p meets B_0..B_{5-p}; those B tiles are contiguous in shared memory
//               and their accumulators p..5 are contiguous in TMEM, so each A_p needs one tcgen05.mma of N = 64 (6-p)
//               (two when N > 256) per K = 32 step: 32 MMAs per k-block, not 84, and with N >= 128 the SS form stays
//               inside the SM's shared-memory bandwidth.
//   warp 18     TMA producer: per 128-wide k-block the 6 B digit tiles (64 x 128 B each, one barrier, double
//               buffered) and the 6 A digit tiles (128 x 128 B each, one stage per digit = a whole k-block in flight)
// What the per-warp clock64 timeline shows (profiles/r01_ozaki_epilogue_timeline.txt): while the tensor queue is busy the
// epilogue's FP64 instructions run 4-6x slower than with the queue idle (in EVERY sub-partition, not only the issuer's),
// i.e. DFMA and tcgen05.mma compete for the same hardware and the MMAs win; the transfer-function epilogue therefore costs
// roughly its stand-alone time on top of the GEMM, and the lever is its FP64 instruction count (71 -> 31 per element).
// Tried and dropped (profiles/r01_ozaki_gemm_variants_v3.txt): staging A in the spare TMEM columns by tcgen05.cp
// (no faster than SS once the MMAs were merged: the copy serialises with the MMAs) or by loader warps + tcgen05.st.
#include <cooperative_groups.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ozaki.cuh"
#include "oz_transcend.cuh"

namespace opl {

namespace {

constexpr int OZ_A_BYTES = OZ_BM * 128;        // 16 KB: 128 rows x 128 int8
constexpr int OZ_B_BYTES = OZ_BN * 128;        //  8 KB
constexpr int OZ_EPI_WARPS = 16;               // 4 per sub-partition: the FP64 chains of the finish phase need the TLP
constexpr int OZ_ZC = OZ_BN / 4;               // columns of a tile row one epilogue thread keeps in registers
constexpr int OZ_OFF_A = 2 * OZ_S * OZ_B_BYTES;                 // 96 KB of B first (1024-aligned tiles)
constexpr int OZ_OFF_PATCH = OZ_OFF_A + OZ_S * OZ_A_BYTES;      // + 96 KB of A: one stage per digit
constexpr int OZ_OFF_BAR = OZ_OFF_PATCH + OZ_EPI_WARPS * 2048;   // per warp: 32 rows x 8 columns, 16-byte chunks swizzled
constexpr int OZ_SMEM = OZ_OFF_BAR + 384 + 1024;
constexpr int OZ_MMA_WARPS = 2;                // issuers in different sub-partitions taking k-blocks in turn (see the MMA issuer below);
                                               // 4 measure the same as 2, but 21 warps are allocated as 24 and leave 80 registers per thread
constexpr int OZ_THREADS = 32 * (OZ_EPI_WARPS + OZ_MMA_WARPS + 1);   // 608
static_assert(OZ_SMEM <= 232448, "shared memory budget");
static_assert(OZ_S * OZ_BN <= 512, "TMEM budget");
static_assert(OZ_S == 6, "the drain recombines exactly six diagonals as two 48-bit integers");

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
// cluster forms: the A tile is loaded once per cluster (every CTA fetches a slab and multicasts it), so a stage is free
// only when every CTA of the cluster has consumed it -> the consumer's commit arrives on the barrier of every CTA
__device__ __forceinline__ void oz_tma_3d_mc(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(oz_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void oz_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(oz_smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t oz_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// 8 doubles (16 columns) of this thread's TMEM lane: the staged FP64 tile lives in the 128 columns the accumulators leave free
__device__ __forceinline__ void oz_tmem_st_f64x8(uint32_t taddr, const double (&z)[8]) {
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[2 * j] = (uint32_t)__double2loint(z[j]);
        v[2 * j + 1] = (uint32_t)__double2hiint(z[j]);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void oz_tmem_ld_f64x8(uint32_t taddr, double (&z)[8]) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = __hiloint2double((int)v[2 * j + 1], (int)v[2 * j]);
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

// exact int64 -> double for |v| < 2^51: bits(1.5 * 2^52) + v, minus 1.5 * 2^52
__device__ __forceinline__ double oz_l2d(long long v) {
    return __longlong_as_double(0x4338000000000000LL + v) - 6755399441055744.0;
}
// a 2^16 + b 2^8 + c as an exact double: three mad.wide.s32 starting from the magic constant (the 64-bit sum never carries
// into the exponent: |a 2^16 + b 2^8 + c| < 2^48), then one FP64 subtraction
__device__ __forceinline__ double oz_join3(int a, int b, int c) {
    long long h;
    asm("mad.wide.s32 %0, %1, 65536, %2;" : "=l"(h) : "r"(a), "l"(0x4338000000000000LL));
    asm("mad.wide.s32 %0, %1, 256, %0;" : "+l"(h) : "r"(b));
    asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(h) : "r"(c));
    return __longlong_as_double(h) - 6755399441055744.0;
}

struct OzTile {
    int m0, n0, k_begin, nkb, split;
};
__device__ __forceinline__ OzTile oz_tile(const OzakiArgs &p, int t, int tiles_m, int tiles_n) {
    OzTile r;
    const int per_split = tiles_m * tiles_n;
    r.split = t / per_split;
    const int u = t - r.split * per_split;
    int m, n;
    if (tiles_n >= 16 && (tiles_n & 1) == 0 && tiles_m >= 8) {
        // grouped raster for square-ish problems (keeps the operand set of one wave inside the L2).  Groups of 8 tile rows;
        // inside a group the tile columns advance in PAIRS so that a 2-CTA cluster still holds neighbouring columns of
        // one row (tiles_n is even whenever clusters are used).
        const int group = u / (8 * tiles_n), within = u - group * 8 * tiles_n;
        const int gm = min(8, tiles_m - group * 8);
        const int pair = within / (2 * gm), rem = within - pair * 2 * gm;
        m = group * 8 + rem / 2;
        n = 2 * pair + (rem & 1);
    } else {
        n = u % tiles_n;
        m = u / tiles_n;
    }
    r.m0 = m * OZ_BM;
    r.n0 = n * OZ_BN;
    r.k_begin = r.split * p.k_per_split;
    const int k_end = min(p.K, r.k_begin + p.k_per_split);
    r.nkb = k_end > r.k_begin ? (k_end - r.k_begin + 127) / 128 : 0;
    return r;
}

// one 8-column chunk of a warp's 32 rows: registers -> per-warp shared patch -> 64-byte row segments in global memory.
// Patch row = 64 bytes = four 16-byte chunks, chunk c of row r stored at c ^ ((r >> 1) & 3): conflict-free both ways.
// MUL: 0 plain store, 1 value x mul[same position], 2 value x (1 - mul^2) -- read where it is written, so mul may alias base
template <int MUL>
__device__ __forceinline__ void oz_store_chunk(double *patch, const double (&v)[8], double *base, int64_t ld, int row0, int col0,
                                               int M, int N, bool vec, int lane, const double *mul = nullptr) {
    double2 *pw = reinterpret_cast<double2 *>(patch + lane * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) pw[j ^ ((lane >> 1) & 3)] = make_double2(v[2 * j], v[2 * j + 1]);
    __syncwarp();
    const int pc = lane & 3, gcol = col0 + 2 * pc;
    if (vec && row0 + 32 <= M && col0 + 8 <= N) {      // interior patch: nothing to check, one address per thread
        const int64_t off = (int64_t)(row0 + (lane >> 2)) * ld + gcol;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int prow = (lane >> 2) + 8 * i;
            double2 q = *reinterpret_cast<const double2 *>(patch + prow * 8 + 2 * (pc ^ ((prow >> 1) & 3)));
            if (MUL != 0) {
                const double2 m = *reinterpret_cast<const double2 *>(mul + off + (int64_t)8 * i * ld);
                q.x *= MUL == 1 ? m.x : 1.0 - m.x * m.x;
                q.y *= MUL == 1 ? m.y : 1.0 - m.y * m.y;
            }
            *reinterpret_cast<double2 *>(base + off + (int64_t)8 * i * ld) = q;
        }
        __syncwarp();
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int prow = (lane >> 2) + 8 * i, grow = row0 + prow;
        double2 q = *reinterpret_cast<const double2 *>(patch + prow * 8 + 2 * (pc ^ ((prow >> 1) & 3)));
        if (grow < M) {
            double *dst = base + (int64_t)grow * ld + gcol;
            if (vec && gcol + 1 < N) {
                if (MUL != 0) {
                    const double2 m = *reinterpret_cast<const double2 *>(mul + (int64_t)grow * ld + gcol);
                    q.x *= MUL == 1 ? m.x : 1.0 - m.x * m.x;
                    q.y *= MUL == 1 ? m.y : 1.0 - m.y * m.y;
                }
                *reinterpret_cast<double2 *>(dst) = q;
            } else {
                if (gcol < N) {
                    if (MUL != 0) {
                        const double m = mul[(int64_t)grow * ld + gcol];
                        q.x *= MUL == 1 ? m : 1.0 - m * m;
                    }
                    dst[0] = q.x;
                }
                if (gcol + 1 < N) {
                    if (MUL != 0) {
                        const double m = mul[(int64_t)grow * ld + gcol + 1];
                        q.y *= MUL == 1 ? m : 1.0 - m * m;
                    }
                    dst[1] = q.y;
                }
            }
        }
    }
    __syncwarp();
}

template <int EPI, int CS>    // CS = CTAs per cluster sharing one A tile (consecutive n tiles of the same rows)
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzakiArgs p) {
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sB = smem;                                   // 2 x 6 x 8 KB
    uint8_t *sA = smem + OZ_OFF_A;                        // 6 x 16 KB: one stage per digit
    double *sPatch = reinterpret_cast<double *>(smem + OZ_OFF_PATCH);
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + OZ_OFF_BAR);
    uint64_t *a_empty = a_full + 4 * OZ_S;           // a_full / b_full: one set per k-block residue mod 4 (a single waiter each)
    uint64_t *b_full = a_empty + OZ_S;
    uint64_t *b_empty = b_full + 4;
    uint64_t *acc_full = b_empty + 2;
    uint64_t *acc_empty = acc_full + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 1);

    // warp roles: 0..15 epilogue (warp w: TMEM lane quarter w % 4, column quarter w / 4), 16 TMA producer, 17 MMA issuer (the
    // arbiter favours the highest warp id of a sub-partition, and the issuing thread is the latency-critical one)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int mma_warp = OZ_EPI_WARPS, producer_warp = OZ_EPI_WARPS + OZ_MMA_WARPS;   // MMA warps 16.., then the producer
    const int tiles_m = (p.M + OZ_BM - 1) / OZ_BM, tiles_n = (p.N + OZ_BN - 1) / OZ_BN;
    const int ntiles = tiles_m * tiles_n * p.splits;
    // tile schedule: cluster c takes tile groups c, c + #clusters, ..; rank r of the cluster takes tile CS*group + r
    const int crank = CS > 1 ? (int)oz_cluster_rank() : 0;
    const int t_first = (int)(blockIdx.x / CS) * CS + crank, t_step = (int)gridDim.x;
    constexpr uint16_t cmask = (uint16_t)((1u << CS) - 1u);

    if (warp == producer_warp && lane == 0) {
        for (int s = 0; s < 4 * OZ_S; ++s) oz_mbar_init(&a_full[s], 1);
        for (int s = 0; s < OZ_S; ++s) oz_mbar_init(&a_empty[s], CS);
        for (int s = 0; s < 4; ++s) oz_mbar_init(&b_full[s], 1);
        for (int s = 0; s < 2; ++s) oz_mbar_init(&b_empty[s], 1);
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
    if (CS > 1) oz_cluster_sync();                  // every CTA's barriers exist before a peer can signal them
    oz_fence_after();
    if (*tmem_ptr != 0) __trap();                   // the CTA owns all 512 columns, so the allocation starts at column 0

    if (warp == producer_warp) {
        // ============ TMA producer (warp-uniform control flow, one elected lane issues) ============
        const bool leader = oz_elect_one();
        uint32_t kbt = 0;                           // k-blocks produced so far (all tiles)
        for (int t = t_first; t < ntiles; t += t_step) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            for (int kb = 0; kb < tl.nkb; ++kb, ++kbt) {
                const int k0 = tl.k_begin + kb * 128;
                const uint32_t bs = kbt & 1, set = kbt & 3;     // smem stage of B; barrier set of this k-block
                uint64_t *a_full_k = a_full + set * OZ_S;
                oz_wait(&b_empty[bs], ((kbt >> 1) & 1) ^ 1);
                if (leader) {
                    oz_expect_tx(&b_full[set], OZ_S * OZ_B_BYTES);
#pragma unroll
                    for (int q = 0; q < OZ_S; ++q) oz_tma_3d(sB + (bs * OZ_S + q) * OZ_B_BYTES, &tmB, k0, tl.n0, q, &b_full[set]);
                }
#pragma unroll
                for (int pp = 0; pp < OZ_S; ++pp) {   // digit pp always travels through stage pp
                    oz_wait(&a_empty[pp], (kbt & 1) ^ 1);
                    if (leader) {
                        oz_expect_tx(&a_full_k[pp], OZ_A_BYTES);
                        if (CS == 1)
                            oz_tma_3d(sA + pp * OZ_A_BYTES, &tmA, k0, tl.m0, pp, &a_full_k[pp]);
                        else      // my slab of the tile, delivered to every CTA of the cluster
                            oz_tma_3d_mc(sA + pp * OZ_A_BYTES + crank * (OZ_A_BYTES / CS), &tmA, k0, tl.m0 + crank * (OZ_BM / CS), pp,
                                         &a_full_k[pp], cmask);
                    }
                }
            }
        }
    } else if (warp >= mma_warp) {
        // ============ MMA issuers (warp-uniform control flow, one elected lane issues) ============
        // p.mma_warps (1 or 2) warps in different sub-partitions take the k-blocks in turn: the plain GEMM gains 6 % over a
        // single issuer (8192^3: 3083 -> 3248 INT8 TOP/s), whose sub-partition's epilogue warps were also the last to finish
        // every tile (per-warp clock64 traces, profiles/r01_ozaki_epilogue_timeline.txt).
        // Ordering needs no extra handshake: INT32 accumulation is exact and commutative, the accumulators are
        // (re)initialised by digit 0 of a tile's first k-block, and digit pp of k-block k+1 can only be issued after
        // digit pp of k-block k has COMPLETED (its stage is reloaded after that commit) -- so the first k-block's
        // initialising MMAs precede every other MMA of the tile, and when the last k-block's last digit is issued all
        // earlier k-blocks are complete: the issuer of the last k-block commits acc_full for the whole tile.
        const int me = warp - mma_warp;
        const uint32_t turn_mask = (uint32_t)p.mma_warps - 1u;
        const int nt_mine = me < p.mma_warps ? ntiles : 0;   // launch_ozaki_gemm keeps mma_warps <= k-blocks per tile (see there)
        // The B digit tiles of one k-block are contiguous 8 KB blocks of 64 rows, and the accumulators of diagonals
        // p..5 are contiguous TMEM columns, so A_p . [B_0; ..; B_{5-p}]^T is ONE MMA of N = 64 (6 - p) columns
        // (cut at N = 256): 8 instructions per K = 32 step instead of 21.  All 512 TMEM columns are ours, so every
        // TMEM address below is a compile-time constant; shared-memory descriptors differ only in their low word
        // (start address >> 4), by compile-time offsets.
        // instruction descriptor: D = S32 (2), A = B = signed INT8 (1), both K-major, N>>3, M>>4
        constexpr uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
        const bool leader = oz_elect_one();
        const uint32_t desc_hi = (uint32_t)(oz_desc(0) >> 32);
        const uint32_t a_lo0 = (uint32_t)oz_desc(oz_smem_u32(sA)), b_lo0 = (uint32_t)oz_desc(oz_smem_u32(sB));
        uint32_t kbt = 0, local = 0;
        for (int t = t_first; t < nt_mine; t += t_step, ++local) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            if (tl.nkb == 0 && me == 0) {            // nothing to contract: hand the (stale) accumulators over as they are
                oz_wait(acc_empty, (local & 1) ^ 1);
                if (leader) oz_commit(acc_full);
                __syncwarp();
            }
            for (int kb = 0; kb < tl.nkb; ++kb, ++kbt) {
                if ((kbt & turn_mask) != (uint32_t)me) continue;
                if (kb == 0) {
                    oz_wait(acc_empty, (local & 1) ^ 1);     // the previous tile has left TMEM
                    oz_fence_after();
                    if (p.trace && blockIdx.x == 0 && leader && local < 32) p.trace[(local * 18 + 17) * 8 + 0] = clock64();
                }
                const uint32_t bs = kbt & 1, set = kbt & 3, par = (kbt >> 2) & 1;
                const uint64_t *a_full_k = a_full + set * OZ_S;
                const uint32_t b_lo = b_lo0 + bs * (uint32_t)((OZ_S * OZ_B_BYTES) >> 4);
                const uint32_t first = kb > 0 ? 1u : 0u;
                // K = 32 steps of this k-block that hold data (the rest is zero padding of the contraction index)
                const int k_left = (p.k_valid > 0 ? p.k_valid : p.K) - (tl.k_begin + kb * 128);
                const int live = k_left >= 128 ? 4 : (k_left + 31) >> 5;
                oz_wait(&b_full[set], par);
#pragma unroll
                for (int pp = 0; pp < OZ_S; ++pp) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)((pp * OZ_A_BYTES) >> 4);
                    oz_wait(const_cast<uint64_t *>(&a_full_k[pp]), par);
                    oz_fence_after();
                    if (leader) {
                        const int nq = OZ_S - pp;              // B digits this A digit meets
#pragma unroll
                        for (int q0 = 0; q0 < nq; q0 += 4) {   // chunks of <= 4 digits = N <= 256
                            const int qn = nq - q0 < 4 ? nq - q0 : 4;
                            const uint32_t idesc = idesc0 | ((uint32_t)((qn * OZ_BN) >> 3) << 17);
                            const uint32_t d_tmem = (uint32_t)((pp + q0) * OZ_BN);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {      // 4 x K32 per 128-byte row, +32 bytes each
                                if (k >= live) break;
                                const uint32_t acc = (pp == 0 && k == 0) ? first : 1u;
                                const uint64_t bd = oz_desc_join(desc_hi, b_lo + (uint32_t)((q0 * OZ_B_BYTES + k * 32) >> 4));
                                oz_umma_ss(d_tmem, oz_desc_join(desc_hi, a_lo + k * 2), bd, idesc, acc);
                            }
                        }
                        if (CS == 1) oz_commit(&a_empty[pp]);
                        else oz_commit_mc(&a_empty[pp], cmask);
                        if (pp == OZ_S - 1) oz_commit(&b_empty[bs]);
                    }
                    __syncwarp();
                }
                if (kb == tl.nkb - 1) {
                    if (leader) oz_commit(acc_full);
                    if (p.trace && blockIdx.x == 0 && leader && local < 32) p.trace[(local * 18 + 17) * 8 + 1] = clock64();
                    __syncwarp();
                }
            }
        }
    } else {
        // ============ epilogue ============
        // Drain: TMEM accumulators -> FP64 (magic-number int->double, Horner over the 6 diagonals, row scale) -> back into
        // the 128 TMEM columns the accumulators leave free (a 128 x 64 FP64 tile is exactly 64 KB); the accumulators are
        // released right after, so the next tile's MMAs overlap the finish phase (column scale, bias, transfer function,
        // stores), which re-reads its values 8 at a time and therefore has the registers to interleave 8 FP64 chains
        // (a dependent DFMA has ~36 cycles of latency on this part).
        const int half = warp >> 2;                   // columns [16 half, 16 half + 16) of the tile
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int lrow = quarter * 32 + lane;
        double *patch = sPatch + warp * 256;
        uint32_t local = 0;
        for (int t = t_first; t < ntiles; t += t_step, ++local) {
            const OzTile tl = oz_tile(p, t, tiles_m, tiles_n);
            const int row = tl.m0 + lrow;
            const double ea = (row < p.M ? p.scale_a[row] : 0.0) * (1.0 / 4294967296.0);
            const uint32_t stage_addr = ((uint32_t)(quarter * 32) << 16) + (uint32_t)(OZ_S * OZ_BN + half * 2 * OZ_ZC);
            const bool tracer = p.trace && blockIdx.x == 0 && lane == 0 && local < 32;
            long long *tr = p.trace + (local * 18 + warp) * 8;
            if (tracer) tr[2] = clock64();
            oz_wait(acc_full, local & 1);
            oz_fence_after();
            if (tracer) tr[3] = clock64();
#pragma unroll
            for (int c0 = 0; c0 < OZ_ZC; c0 += 8) {
                int32_t v[OZ_S][8];
#pragma unroll
                for (int d = 0; d < OZ_S; ++d)
                    oz_tmem_ld8(((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * OZ_BN + half * OZ_ZC + c0), v[d]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // sum_d acc_d 256^-d, recombined in INTEGER arithmetic (the FP64 pipe is the scarce one here): two exact
                // 48-bit integers hi = acc0 2^16 + acc1 2^8 + acc2 and lo = acc3 2^16 + acc4 2^8 + acc5, each converted exactly
                // (magic-number add), joined by ONE rounding: (hi + lo 2^-24) 2^-16.  The 2^-16 rides on the row scale.
                double zc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    zc[j] = fma(oz_join3(v[3][j], v[4][j], v[5][j]), 1.0 / 16777216.0, oz_join3(v[0][j], v[1][j], v[2][j])) * ea;
                oz_tmem_st_f64x8(stage_addr + 2 * c0, zc);     // FP64 tile -> the 128 spare TMEM columns
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            oz_fence_before();
            __syncwarp();
            // (round 1 released the accumulators of the FP64-heavy epilogues only after their finish phase; with the leaner
            //  finish phase the early release measures 0.006-0.010 ms faster, profiles/r02_ozaki_epilogue_decomposition.txt)
            const bool LATE = p.late_release != 0;              // tools only (OPL_OZ_LATE=1)
            if (lane == 0 && !LATE) oz_arrive(acc_empty);      // accumulators free again: the next tile's MMAs overlap the finish phase
            if (tracer) tr[4] = clock64();
            // ---- finish: 2 chunks of 8 columns ----
            double *C = p.C + (int64_t)tl.split * p.split_stride;
            const bool vec_c = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
            const bool want_x = p.aux_out && (EPI == OZ_EPI_SOFTPLUS || EPI == OZ_EPI_GAUSSIAN);
            const bool vec_x = want_x && ((p.ldaux & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0);
            const int row0 = tl.m0 + quarter * 32;
#pragma unroll
            for (int ch = 0; ch < OZ_ZC / 8; ++ch) {
                const int col0 = tl.n0 + half * OZ_ZC + ch * 8;
                double r[8], x[8], zz[8], zc[8];
                oz_tmem_ld_f64x8(stage_addr + 16 * ch, zc);
                if (tracer && ch == 0) tr[6] = clock64();
                // column scale (+ bias).  Interior tiles skip the predicates (the finish phase is bound by its instruction
                // count, see the header); the branches are warp-uniform
                const bool inner = col0 + 8 <= p.N;
                const bool masked = p.colmask != nullptr;
                if (inner) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) zz[j] = fma(zc[j], __ldg(p.scale_b + col0 + j), __ldg(p.bias + col0 + j));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) zz[j] = zc[j] * __ldg(p.scale_b + col0 + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = col0 + j;
                        const bool ok = col < p.N;
                        const double eb = ok ? __ldg(p.scale_b + col) : 0.0;
                        const double bi = (p.bias && ok) ? __ldg(p.bias + col) : 0.0;
                        zz[j] = zc[j] * eb + bi;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.0;
                // the transfer function of all 8 columns in ONE basic block, so the FP64 chains interleave
                if (EPI == OZ_EPI_SOFTPLUS) {
                    bool regular = true;
#pragma unroll
                    for (int j = 0; j < 8; ++j) regular &= ozt_is_regular(zz[j]);
                    if (__all_sync(0xffffffffu, regular)) ozt_softplus_logistic_regular_n<8>(zz, r, x);
                    else ozt_softplus_logistic_n<8>(zz, r, x);      // inf / NaN / |z| >= 708 somewhere in the warp: full form
                } else if (EPI == OZ_EPI_TANH) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = tanh(zz[j]);
                } else if (EPI == OZ_EPI_GAUSSIAN) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = exp(-(zz[j] * zz[j] / p.tparam));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = zz[j];
                }
                if (masked) {     // dropout: the mask multiplies the transfer OUTPUT (mlp.py:438-445)
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] *= (col0 + j < p.N) ? __ldg(p.colmask + col0 + j) : 1.0;
                }
                if (EPI == OZ_EPI_GAUSSIAN) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = -2.0 * zz[j] * r[j] / p.tparam;   // derivative uses the masked output, as the reference does
                }
                constexpr int MUL = EPI == OZ_EPI_MUL_AUX ? 1 : (EPI == OZ_EPI_MUL_DTANH ? 2 : 0);
                if (tracer && ch == 0) tr[7] = clock64();
                oz_store_chunk<MUL>(patch, r, C, p.ldc, row0, col0, p.M, p.N,
                                    vec_c && (MUL == 0 || (reinterpret_cast<uintptr_t>(p.aux_in) & 15) == 0), lane, p.aux_in);
                if (want_x) oz_store_chunk<0>(patch, x, p.aux_out, p.ldaux, row0, col0, p.M, p.N, vec_x, lane);
            }
            if (lane == 0 && LATE) oz_arrive(acc_empty);
            if (tracer) tr[5] = clock64();
        }
    }
    oz_fence_before();
    __syncthreads();
    if (CS > 1) oz_cluster_sync();                  // no CTA leaves while a peer may still write into it / signal it
    if (warp == mma_warp) {
        oz_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u));
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

// Signed radix-256 digits of four scaled values (|r| <= 0.498), packed one 32-bit word per digit plane (byte j = value j).
// fx = rint(r 2^48) as an integer; u = fx + 0x808080808080 (128 in every byte) lies in [0, 2^48) and its bytes, most
// significant first, are the digits + 128: r = sum_s (byte_s - 128) 256^-(s+1) exactly, |r - fx 2^-48| <= 2^-49.
// One FP64 multiply and one conversion per value instead of a floor / clamp / subtract chain per digit.
__device__ __forceinline__ void oz_digits4(const double (&r)[4], uint32_t (&w)[OZ_S]) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long u =
            ((unsigned long long)(__double2ll_rn(r[j] * 281474976710656.0) + 0x808080808080LL)) ^ 0x808080808080ULL;
        lo[j] = (uint32_t)u;
        hi[j] = (uint32_t)(u >> 32);
    }
    // plane s takes byte (5 - s) of every value: bytes 5,4 live in hi (bytes 1,0), bytes 3..0 in lo
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) {
        const int byte = 5 - s;
        const uint32_t *src = byte >= 4 ? hi : lo;
        const uint32_t b = (uint32_t)(byte & 3);
        const uint32_t p01 = __byte_perm(src[0], src[1], b | ((4 + b) << 4));          // bytes: v0, v1
        const uint32_t p23 = __byte_perm(src[2], src[3], b | ((4 + b) << 4));          // bytes: v2, v3
        w[s] = __byte_perm(p01, p23, 0x5410);
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
        double r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = c4 + j < cols ? x[c4 + j] * inv : 0.0;
        oz_digits4(r, w);
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
// Scale source: MODE 0 `scale` holds finished scales; MODE 1 `colmax` holds raw column maxima (bit patterns gathered
// with atomicMax by the producer of src), the scale is derived here and block row 0 publishes it; MODE 2 (small
// matrices) every block first scans its 32 columns over ALL rows itself, then walks all row chunks (grid.x == 1);
// MODE 3 (cooperative launch, matrices whose grid is co-resident: the weight matrices) every block leaves the maxima of
// its own 128 x 32 patch in `partial`, the grid synchronises, and the maxima are folded from there: one launch instead of three.
// Tile order of the streaming transposed-digit passes: a 1-D grid walks 8 x 8 patches of (column tile, row tile), so the CTAs that
// run at the same time read 2 KB runs of the source rows AND write 1 KB runs of the digit rows (row-tile-fast order scatters the
// reads into 256-byte pieces, column-tile-fast order scatters the writes into 128-byte pieces).
__device__ __forceinline__ bool oz_patch_raster(int64_t b, int64_t ctiles, int64_t rtiles, int64_t &ct, int64_t &rt) {
    const int64_t pc = (ctiles + 7) >> 3, p = b >> 6;
    const int w = (int)(b & 63);
    const int64_t pr = p / pc;
    ct = (p - pr * pc) * 8 + (w & 7);
    rt = pr * 8 + (w >> 3);
    return ct < ctiles && rt < rtiles;
}
static int64_t oz_patch_grid(int64_t ctiles, int64_t rtiles) { return ((ctiles + 7) >> 3) * ((rtiles + 7) >> 3) * 64; }

template <int MODE>
__global__ void __launch_bounds__(256) slice_cols_transposed_kernel(const double *__restrict__ src, int64_t rows, int64_t cols,
                                                                    int64_t ld, double *__restrict__ scale,
                                                                    const double *__restrict__ colmax,
                                                                    int8_t *__restrict__ out, int64_t cols_pad, int64_t rows_pad,
                                                                    int64_t row_off, double *partial) {
    __shared__ uint32_t tile[OZ_S][32][33];    // [digit][col][word along rows], padded
    __shared__ double smax[8][33];
    // MODE 0 / 1 (streaming passes over big matrices): 1-D grid in patch order, one tile per CTA
    constexpr bool PATCHES = MODE == 0 || MODE == 1;
    int64_t row_blk = blockIdx.x, row_blks = gridDim.x, col_blk = blockIdx.y;
    if (PATCHES) {
        row_blks = rows_pad >> 7;
        if (!oz_patch_raster(blockIdx.x, (cols + 31) >> 5, row_blks, col_blk, row_blk)) return;
    }
    const int64_t c0 = col_blk * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    const int64_t c = c0 + tx;
    double sc = 1.0;
    if (MODE == 0) {
        if (c < cols) sc = scale[c];
    } else if (MODE == 1) {
        if (c < cols) {
            sc = oz_scale_of(colmax[c]);
            if (row_blk == 0 && ty == 0) scale[c] = sc;
        }
    } else if (MODE == 3) {
        double m = 0.0;
        bool bad = false;
        const int64_t rb = blockIdx.x * 128;
        if (c < cols)
            for (int64_t r = rb + ty; r < min(rows, rb + 128); r += 8) {
                const double v = fabs(src[r * ld + c]);
                bad |= !(v <= 1.7976931348623157e308);
                m = fmax(m, v);
            }
        smax[ty][tx] = bad ? __longlong_as_double(0x7ff8000000000000LL) : m;
        __syncthreads();
        if (ty == 0 && c < cols) {
            double t = 0.0;
            bool nan = false;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double v = smax[k][tx];
                nan |= v != v;
                t = fmax(t, v);
            }
            partial[blockIdx.x * cols + c] = nan ? __longlong_as_double(0x7ff8000000000000LL) : t;
        }
        __threadfence();
        cooperative_groups::this_grid().sync();
        if (c < cols) {
            double t = 0.0;
            bool nan = false;
            for (unsigned k = 0; k < gridDim.x; ++k) {
                const double v = __ldcg(partial + k * cols + c);
                nan |= v != v;
                t = fmax(t, v);
            }
            sc = nan ? __longlong_as_double(0x7ff8000000000000LL) : oz_scale_of(t);
            if (blockIdx.x == 0 && ty == 0) scale[c] = sc;
        }
    } else {
        double m = 0.0;
        bool bad = false;
        if (c < cols)
            for (int64_t r = ty; r < rows; r += 8) {
                const double v = fabs(src[r * ld + c]);
                bad |= !(v <= 1.7976931348623157e308);
                m = fmax(m, v);
            }
        smax[ty][tx] = bad ? __longlong_as_double(0x7ff8000000000000LL) : m;
        __syncthreads();
        double t = 0.0;
        bool nan = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double v = smax[k][tx];
            nan |= v != v;
            t = fmax(t, v);
        }
        sc = nan ? __longlong_as_double(0x7ff8000000000000LL) : oz_scale_of(t);
        if (c < cols && ty == 0) scale[c] = sc;
    }
    const double inv = sc == sc ? 1.0 / sc : 0.0;
    const int64_t plane = cols_pad * rows_pad;
    for (int64_t r0 = (int64_t)row_blk * 128; r0 < rows_pad; r0 += (int64_t)row_blks * 128) {
        if (r0 + 128 <= rows && c0 + 32 <= cols) {
            // interior tile: all 16 values of the thread requested up front through a stepped pointer, no bounds tests
            const double *sp = src + (r0 + ty * 16) * ld + c;
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j, sp += ld) v[j] = *sp;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t w[OZ_S];
                double q4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) q4[j] = v[4 * g + j] * inv;
                oz_digits4(q4, w);
#pragma unroll
                for (int s = 0; s < OZ_S; ++s) tile[s][tx][ty * 4 + g] = w[s];
            }
            __syncthreads();
            const int cc = threadIdx.x >> 3, seg = threadIdx.x & 7;     // one 16-byte segment of one output row per digit plane
            int8_t *op = out + (c0 + cc + row_off) * rows_pad + r0 + seg * 16;
#pragma unroll
            for (int s = 0; s < OZ_S; ++s, op += plane) {
                uint4 q;
                q.x = tile[s][cc][seg * 4 + 0];
                q.y = tile[s][cc][seg * 4 + 1];
                q.z = tile[s][cc][seg * 4 + 2];
                q.w = tile[s][cc][seg * 4 + 3];
                *reinterpret_cast<uint4 *>(op) = q;
            }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int wq = ty * 4 + g;             // word index along rows: rows r0 + 4 wq ..
            uint32_t w[OZ_S];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t r = r0 + wq * 4 + j;
                v[j] = (r < rows && c < cols) ? src[r * ld + c] * inv : 0.0;
            }
            oz_digits4(v, w);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) tile[s][tx][wq] = w[s];
        }
        __syncthreads();
        // write: OZ_S x 32 output rows of 128 bytes = 8 x 16 bytes each
        for (int idx = threadIdx.x; idx < OZ_S * 32 * 8; idx += 256) {
            const int s = idx >> 8, cc = (idx >> 3) & 31, seg = idx & 7;
            const int64_t oc = c0 + cc + row_off, r = r0 + seg * 16;
            if (c0 + cc < cols && r < rows_pad) {
                uint4 q;
                q.x = tile[s][cc][seg * 4 + 0];
                q.y = tile[s][cc][seg * 4 + 1];
                q.z = tile[s][cc][seg * 4 + 2];
                q.w = tile[s][cc][seg * 4 + 3];
                *reinterpret_cast<uint4 *>(out + ((int64_t)s * cols_pad + oc) * rows_pad + r) = q;
            }
        }
        __syncthreads();
    }
}

// Digits of delta_prev^T WITHOUT delta_prev: for a narrow last layer (C <= 16 outputs) the matrix
//     delta_prev[r][j] = (sum_c deltaL[r][c] W[j][c]) x slope(aux[r][j])
// is a rank-C product, so the weight-gradient GEMM's B operand can be cut straight from deltaL (rows x C), W (H x C) and the
// slope matrix: the fused head then neither forms nor stores delta_prev (a third of its DMMA work, 2 x rows x H x 8 bytes of
// traffic), and this kernel reads the slope matrix instead of delta_prev -- the same bytes as the plain digit pass.
// Scales: a column's maximum is not known before its values are, so the scale comes from the rigorous bound
//     |delta_prev[r][j]| <= sum_c max_r|deltaL[r][c]| |W[j][c]|        (|slope| <= 1 for softplus', 1 - tanh^2 and linear)
// with the per-class maxima gathered by the head (atomicMax on bit patterns).  The bound is loose by the cancellation inside
// the sum (measured ~2-4x, i.e. 1-2 of the 48 digit bits): the digits stay an error-free transformation relative to it.
// A non-finite value poisons the column's scale (NaN pattern by atomicMax), as the plain pass does.
// Block = 128 rows x 32 columns (as slice_cols_transposed_kernel); the block's deltaL rows sit in shared memory (broadcast
// reads), a thread keeps its column's row of W in registers.  mode: 0 slope 1, 1 aux = f', 2 aux = tanh activation.
template <int MAXC, int MODE>
__global__ void __launch_bounds__(256, 4) slice_delta_prev_transposed_kernel(const double *__restrict__ deltaL, int C,
                                                                          const double *__restrict__ W, const double *aux,
                                                                          int64_t rows, int64_t cols, int64_t ld,
                                                                          const double *__restrict__ classmax,
                                                                          double *scale, int8_t *__restrict__ out,
                                                                          int64_t cols_pad, int64_t rows_pad, int64_t row_off) {
    __shared__ uint32_t tile[OZ_S][32][33];
    __shared__ __align__(16) double dl[128][MAXC];
    // (row tile = fast grid index here: the patch order of the plain transposed pass measured 0.076 against 0.070 ms at config 2,
    //  whose 256 columns are a single patch across)
    const int64_t row_blk = blockIdx.x;
    const int64_t c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = c0 + tx;
    double w[MAXC];
    double bound = 0.0;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
        w[k] = (c < cols && k < C) ? W[c * C + k] : 0.0;
        if (k < C) bound += classmax[k] * fabs(w[k]);        // NaN pattern in classmax -> NaN bound -> NaN scale
    }
    const double sc = oz_scale_of(bound);
    unsigned long long *scale_bits = reinterpret_cast<unsigned long long *>(scale + row_off);
    if (c < cols && row_blk == 0 && ty == 0) atomicMax(scale_bits + c, (unsigned long long)__double_as_longlong(sc));
    const double inv = sc == sc ? 1.0 / sc : 0.0;
    bool bad = false;
    const int64_t plane = cols_pad * rows_pad;
    for (int64_t r0 = row_blk * 128; r0 < rows_pad; r0 += (int64_t)gridDim.x * 128) {
        if (r0 + 128 <= rows && c0 + 32 <= cols) {
            // ---- interior tile: no bounds tests, pointers stepped instead of re-derived (the kernel was bound by its
            //      integer / predicate instructions, ~60 of 119 per element; profiles/r02_ncu_eval_kernels.txt) ----
            const double *ap = aux + (r0 + ty * 16) * ld + c;     // the slope values are this kernel's HBM stream
            double a[2][4];
            if (MODE != 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j, ap += ld) a[0][j] = *ap;
            }
            __syncthreads();
            if (MAXC == C) {                                       // rows r0 .. r0 + 127 of delta_L are one contiguous block
                const double *src = deltaL + r0 * MAXC;
                double *dst = &dl[0][0];
#pragma unroll
                for (int i = 0; i < (128 * MAXC + 255) / 256; ++i) {
                    const int e = i * 256 + (int)threadIdx.x;
                    if ((128 * MAXC) % 256 == 0 || e < 128 * MAXC) dst[e] = src[e];
                }
            } else {
                for (int i = threadIdx.x; i < 128 * MAXC; i += 256) {
                    const int r = i / MAXC, k = i - r * MAXC;
                    dl[r][k] = k < C ? deltaL[(r0 + r) * C + k] : 0.0;
                }
            }
            __syncthreads();
            double chk = 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t wd[OZ_S];
                double v[4];
                if (MODE != 0 && g < 3) {
#pragma unroll
                    for (int j = 0; j < 4; ++j, ap += ld) a[(g + 1) & 1][j] = *ap;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 *dp = reinterpret_cast<const double2 *>(&dl[ty * 16 + g * 4 + j][0]);   // 16-byte broadcasts
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int k2 = 0; k2 < MAXC / 2; ++k2) {
                        const double2 d = dp[k2];
                        s0 = fma(d.x, w[2 * k2], s0);
                        s1 = fma(d.y, w[2 * k2 + 1], s1);
                    }
                    const double s_ = s0 + s1;
                    const double val = MODE == 2 ? s_ * (1.0 - a[g & 1][j] * a[g & 1][j]) : (MODE == 1 ? s_ * a[g & 1][j] : s_);
                    v[j] = val * inv;
                }
                chk = fma((v[0] + v[1]) + (v[2] + v[3]), 0.0, chk);   // inf or NaN anywhere -> NaN
                oz_digits4(v, wd);
#pragma unroll
                for (int s = 0; s < OZ_S; ++s) tile[s][tx][ty * 4 + g] = wd[s];
            }
            bad |= chk != chk;
            __syncthreads();
            // every thread owns one 16-byte segment of one output row, in each of the digit planes
            const int cc = threadIdx.x >> 3, seg = threadIdx.x & 7;
            int8_t *op = out + (c0 + cc + row_off) * rows_pad + r0 + seg * 16;
#pragma unroll
            for (int s = 0; s < OZ_S; ++s, op += plane) {
                uint4 q;
                q.x = tile[s][cc][seg * 4 + 0];
                q.y = tile[s][cc][seg * 4 + 1];
                q.z = tile[s][cc][seg * 4 + 2];
                q.w = tile[s][cc][seg * 4 + 3];
                *reinterpret_cast<uint4 *>(op) = q;
            }
            continue;
        }
        // ---- edge tile ----
        __syncthreads();
        for (int i = threadIdx.x; i < 128 * MAXC; i += 256) {
            const int r = i / MAXC, k = i - r * MAXC;
            dl[r][k] = (r0 + r < rows && k < C) ? deltaL[(r0 + r) * C + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const int wq = ty * 4 + g;             // word index along rows: rows r0 + 4 wq ..
            uint32_t wd[OZ_S];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int rl = wq * 4 + j;
                const int64_t r = r0 + rl;
                double val = 0.0;
                if (r < rows && c < cols) {
                    double s0 = 0.0, s1 = 0.0;             // same summation order as the interior tiles
#pragma unroll
                    for (int k2 = 0; k2 < MAXC / 2; ++k2) {
                        s0 = fma(dl[rl][2 * k2], w[2 * k2], s0);
                        s1 = fma(dl[rl][2 * k2 + 1], w[2 * k2 + 1], s1);
                    }
                    const double s_ = s0 + s1;
                    const double av = MODE != 0 ? aux[r * ld + c] : 1.0;
                    val = MODE == 2 ? s_ * (1.0 - av * av) : (MODE == 1 ? s_ * av : s_);
                    bad |= !(fabs(val) <= 1.7976931348623157e308);
                }
                v[j] = val * inv;
            }
            oz_digits4(v, wd);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) tile[s][tx][wq] = wd[s];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < OZ_S * 32 * 8; idx += 256) {
            const int s = idx >> 8, cc = (idx >> 3) & 31, seg = idx & 7;
            const int64_t oc = c0 + cc + row_off, r = r0 + seg * 16;
            if (c0 + cc < cols && r < rows_pad) {
                uint4 q;
                q.x = tile[s][cc][seg * 4 + 0];
                q.y = tile[s][cc][seg * 4 + 1];
                q.z = tile[s][cc][seg * 4 + 2];
                q.w = tile[s][cc][seg * 4 + 3];
                *reinterpret_cast<uint4 *>(out + ((int64_t)s * cols_pad + oc) * rows_pad + r) = q;
            }
        }
    }
    if (bad && c < cols) atomicMax(scale_bits + c, 0x7ff8000000000000ULL);      // poison: the product becomes NaN
}

// Trial point AND the digits of W_0^T in one cooperative launch (an evaluation used to start with two: the trial-point kernel,
// then the cooperative max + digit kernel of W_0): every block forms w = x + fl(alpha dir) for its 128 x 32 patch of W_0, writes
// it to the trial vector and takes the patch's column maxima; the blocks also share the rest of the parameter vector (bias, the
// other layers); after the grid barrier the column maxima are folded and the patch is cut into digits (MODE 3 of
// slice_cols_transposed_kernel, same arithmetic and layout).
struct TrialW0Args {
    const double *x, *dir;     // parameters, direction (may be null)
    double alpha;
    double *w;                 // trial vector out (n entries)
    int64_t n, off;            // W_0 starts at `off` of the flat vector
    int64_t rows, cols;        // W_0 is rows x cols row-major (d0 x d1)
    double *scale;             // cols scales out
    int8_t *out;               // digits of W_0^T: [S][cols_pad][rows_pad]
    int64_t cols_pad, rows_pad;
    double *partial;           // gridDim.x x cols column maxima
    int *flag;                 // domain flag of the evaluation, reset here
    double *zero_a, *zero_b;   // two small arrays the evaluation accumulates into by atomicMax later (class maxima, delta scales),
    int zero_a_len, zero_b_len;   // cleared here instead of by two memset launches; may be null
};

__global__ void __launch_bounds__(256) trial_w0_digits_kernel(const TrialW0Args a) {
    __shared__ uint32_t tile[OZ_S][32][33];
    __shared__ double smax[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c0 = blockIdx.y * 32, c = c0 + tx;
    const int64_t rb = blockIdx.x * 128;
    const double *x0 = a.x + a.off, *d0 = a.dir ? a.dir + a.off : nullptr;
    double *w0 = a.w + a.off;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        if (threadIdx.x == 0) *a.flag = 0;
        for (int i = threadIdx.x; i < a.zero_a_len; i += 256) a.zero_a[i] = 0.0;
        for (int i = threadIdx.x; i < a.zero_b_len; i += 256) a.zero_b[i] = 0.0;
    }
    // ---- the rest of the vector (everything outside W_0), shared by all blocks ----
    {
        const int64_t nb = (int64_t)gridDim.x * gridDim.y, b = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t w0_len = a.rows * a.cols, rest = a.n - w0_len;
        for (int64_t i = b * 256 + threadIdx.x; i < rest; i += nb * 256) {
            const int64_t e = i < a.off ? i : i + w0_len;
            a.w[e] = a.dir ? __dadd_rn(a.x[e], __dmul_rn(a.alpha, a.dir[e])) : a.x[e];
        }
    }
    // ---- pass 1: trial values of this patch, written out, and their column maxima ----
    double m = 0.0;
    bool bad = false;
    if (c < a.cols) {
        // all 16 rows of the thread requested back to back (a loop with a run-time bound serialises the DRAM latencies)
        double xv[16], dv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int64_t r = rb + ty + 8 * i;
            const bool ok = r < a.rows;
            xv[i] = ok ? x0[r * a.cols + c] : 0.0;
            dv[i] = (ok && d0) ? d0[r * a.cols + c] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int64_t r = rb + ty + 8 * i;
            if (r < a.rows) {
                const double v = d0 ? __dadd_rn(xv[i], __dmul_rn(a.alpha, dv[i])) : xv[i];
                w0[r * a.cols + c] = v;
                const double av = fabs(v);
                bad |= !(av <= 1.7976931348623157e308);
                m = fmax(m, av);
            }
        }
    }
    smax[ty][tx] = bad ? __longlong_as_double(0x7ff8000000000000LL) : m;
    __syncthreads();
    if (ty == 0 && c < a.cols) {
        double t = 0.0;
        bool nan = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double v = smax[k][tx];
            nan |= v != v;
            t = fmax(t, v);
        }
        a.partial[blockIdx.x * a.cols + c] = nan ? __longlong_as_double(0x7ff8000000000000LL) : t;
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    double sc = 1.0;
    if (c < a.cols) {
        double t = 0.0;
        bool nan = false;
        for (unsigned k = 0; k < gridDim.x; ++k) {
            const double v = __ldcg(a.partial + k * a.cols + c);
            nan |= v != v;
            t = fmax(t, v);
        }
        sc = nan ? __longlong_as_double(0x7ff8000000000000LL) : oz_scale_of(t);
        if (blockIdx.x == 0 && ty == 0) a.scale[c] = sc;
    }
    const double inv = sc == sc ? 1.0 / sc : 0.0;
    // ---- pass 2: digits of the patch (it was written by this very block: visible after the barrier) ----
    const int64_t r0 = rb;
    if (r0 < a.rows_pad) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int wq = ty * 4 + g;
            uint32_t wd[OZ_S];
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t r = r0 + wq * 4 + j;
                v[j] = (r < a.rows && c < a.cols) ? __ldcg(w0 + r * a.cols + c) * inv : 0.0;
            }
            oz_digits4(v, wd);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) tile[s][tx][wq] = wd[s];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < OZ_S * 32 * 8; idx += 256) {
            const int s = idx >> 8, cc = (idx >> 3) & 31, seg = idx & 7;
            const int64_t oc = c0 + cc, r = r0 + seg * 16;
            if (c0 + cc < a.cols && r < a.rows_pad) {
                uint4 q;
                q.x = tile[s][cc][seg * 4 + 0];
                q.y = tile[s][cc][seg * 4 + 1];
                q.z = tile[s][cc][seg * 4 + 2];
                q.w = tile[s][cc][seg * 4 + 3];
                *reinterpret_cast<uint4 *>(a.out + ((int64_t)s * a.cols_pad + oc) * a.rows_pad + r) = q;
            }
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
                                double *scratch, int64_t scratch_len, int64_t row_off, int64_t m_pad, const double *colmax_raw,
                                cudaStream_t stream, int64_t *launches) {
    const int64_t cp = m_pad > 0 ? m_pad : ozaki_pad(cols + row_off), rp = ozaki_pad(rows);
    const unsigned gy = (unsigned)ceil_div(cols, 32);
    if (colmax_raw) {            // maxima gathered by the kernel that produced src
        slice_cols_transposed_kernel<1><<<(unsigned)oz_patch_grid(gy, rp / 128), 256, 0, stream>>>(
            src, rows, cols, ld, scale + row_off, colmax_raw, slices, cp, rp, row_off, nullptr);
        OPL_LAUNCH_CHECK();
        if (launches) *launches += 1;
        return OPL_OK;
    }
    if (rows <= 256) {           // tiny matrix: one launch, every block scans its own columns
        slice_cols_transposed_kernel<2><<<dim3(1, gy), 256, 0, stream>>>(src, rows, cols, ld, scale + row_off, nullptr, slices, cp, rp,
                                                                         row_off, nullptr);
        OPL_LAUNCH_CHECK();
        if (launches) *launches += 1;
        return OPL_OK;
    }
    {   // the whole grid fits on the device at once (weight matrices): cooperative single launch
        const int64_t gx = ceil_div(rp, 128);
        static int coop_blocks_per_sm = -1;
        if (coop_blocks_per_sm < 0) {
            int v = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, slice_cols_transposed_kernel<3>, 256, 0) != cudaSuccess) v = 0;
            coop_blocks_per_sm = v;
        }
        if (gx * gy <= (int64_t)coop_blocks_per_sm * sm_count() && gx * cols <= scratch_len) {
            double *scale_p = scale + row_off;
            const double *no_colmax = nullptr;
            void *params[] = {(void *)&src, (void *)&rows, (void *)&cols, (void *)&ld, (void *)&scale_p, (void *)&no_colmax,
                              (void *)&slices, (void *)&cp, (void *)&rp, (void *)&row_off, (void *)&scratch};
            OPL_CUDA(cudaLaunchCooperativeKernel((const void *)slice_cols_transposed_kernel<3>, dim3((unsigned)gx, gy), dim3(256),
                                                 params, 0, stream));
            if (launches) *launches += 1;
            return OPL_OK;
        }
    }
    int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 256), scratch_len / std::max<int64_t>(cols, 1)));
    nblk = std::min<int64_t>(nblk, 4 * (int64_t)sm_count());
    const int64_t rpb = ceil_div(rows, nblk);
    nblk = ceil_div(rows, rpb);
    dim3 g1((unsigned)ceil_div(cols, 32), (unsigned)nblk);
    colabsmax_partial_kernel<<<g1, dim3(32, 8), 0, stream>>>(src, rows, cols, ld, rpb, scratch);
    colabsmax_final_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, stream>>>(scratch, (int)nblk, cols, scale + row_off);
    slice_cols_transposed_kernel<0><<<(unsigned)oz_patch_grid(gy, rp / 128), 256, 0, stream>>>(
        src, rows, cols, ld, scale + row_off, nullptr, slices, cp, rp, row_off, nullptr);
    OPL_LAUNCH_CHECK();
    if (launches) *launches += 3;
    return OPL_OK;
}

// *done = true when the fused launch ran (the grid must be co-resident); otherwise nothing was enqueued and the caller runs the
// trial-point kernel and the plain digit pass
int ozaki_trial_w0_digits(const double *x, const double *dir, double alpha, double *w, int64_t n, int64_t off, int64_t rows,
                          int64_t cols, int8_t *slices, double *scale, double *scratch, int64_t scratch_len, int64_t m_pad,
                          int *flag, double *zero_a, int zero_a_len, double *zero_b, int zero_b_len, cudaStream_t stream,
                          int64_t *launches, bool *done) {
    *done = false;
    const int64_t rp = ozaki_pad(rows), cp = m_pad > 0 ? m_pad : ozaki_pad(cols);
    const int64_t gx = ceil_div(rp, 128), gy = ceil_div(cols, 32);
    static int per_sm = -1;
    if (per_sm < 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, trial_w0_digits_kernel, 256, 0) != cudaSuccess) v = 0;
        per_sm = v;
    }
    if (rows <= 256 || gx * gy > (int64_t)per_sm * sm_count() || gx * cols > scratch_len) return OPL_OK;
    TrialW0Args a = {};
    a.x = x;
    a.dir = dir;
    a.alpha = alpha;
    a.w = w;
    a.n = n;
    a.off = off;
    a.rows = rows;
    a.cols = cols;
    a.scale = scale;
    a.out = slices;
    a.cols_pad = cp;
    a.rows_pad = rp;
    a.partial = scratch;
    a.flag = flag;
    a.zero_a = zero_a;
    a.zero_a_len = zero_a ? zero_a_len : 0;
    a.zero_b = zero_b;
    a.zero_b_len = zero_b ? zero_b_len : 0;
    void *params[] = {(void *)&a};
    OPL_CUDA(cudaLaunchCooperativeKernel((const void *)trial_w0_digits_kernel, dim3((unsigned)gx, (unsigned)gy), dim3(256), params, 0,
                                         stream));
    if (launches) ++*launches;
    *done = true;
    return OPL_OK;
}

int ozaki_slice_delta_prev(const double *deltaL, int C, const double *W, const double *aux, int mode, int64_t rows, int64_t cols,
                           int64_t ld, const double *classmax, int8_t *slices, double *scale, int64_t m_pad,
                           cudaStream_t stream, int64_t *launches, bool scale_cleared) {
    if (C < 1 || C > 16) return fail(OPL_E_ARG, "ozaki_slice_delta_prev: C must be in [1, 16]");
    const int64_t cp = m_pad > 0 ? m_pad : ozaki_pad(cols), rp = ozaki_pad(rows);
    if (!scale_cleared)      // scales arrive by atomicMax on their bit patterns
        OPL_CUDA(cudaMemsetAsync(scale, 0, sizeof(double) * (size_t)cols, stream));
    const dim3 grid((unsigned)ceil_div(rp, 128), (unsigned)ceil_div(cols, 32));
#define OPL_DP_LAUNCH(CE, MD)                                                                                               \
    slice_delta_prev_transposed_kernel<CE, MD><<<grid, 256, 0, stream>>>(deltaL, C, W, aux, rows, cols, ld, classmax, scale, \
                                                                         slices, cp, rp, 0)
#define OPL_DP_MODES(CE)                                                       \
    do {                                                                       \
        if (mode == 1) OPL_DP_LAUNCH(CE, 1);                                   \
        else if (mode == 2) OPL_DP_LAUNCH(CE, 2);                              \
        else OPL_DP_LAUNCH(CE, 0);                                             \
    } while (0)
    // class count rounded up to an even template width; exact widths stage delta_L as one contiguous copy
    if (C <= 4) OPL_DP_MODES(4);
    else if (C <= 8) OPL_DP_MODES(8);
    else if (C <= 10) OPL_DP_MODES(10);
    else if (C <= 12) OPL_DP_MODES(12);
    else OPL_DP_MODES(16);
#undef OPL_DP_MODES
#undef OPL_DP_LAUNCH
    OPL_LAUNCH_CHECK();
    if (launches) ++*launches;
    return OPL_OK;
}

int launch_ozaki_gemm(const int8_t *a_slices, int64_t a_rows_pad, const int8_t *b_slices, int64_t b_rows_pad, int64_t k_pad,
                      const OzakiArgs &args, cudaStream_t stream, int64_t *launches) {
    if (args.M <= 0 || args.N <= 0) return OPL_OK;
    if (args.k_per_split % 128 || args.k_per_split > OZ_MAX_K || args.splits < 1 ||
        (int64_t)args.k_per_split * (args.splits - 1) >= args.K)
        return fail(OPL_E_ARG, "ozaki GEMM: k_per_split must be a multiple of 128, <= %d, and every split non-empty", OZ_MAX_K);
    const int64_t tiles_n = ceil_div(args.N, OZ_BN);
    const int64_t tiles = ceil_div(args.M, OZ_BM) * tiles_n * args.splits;
    // cluster size: consecutive n tiles share their A tile (loaded once, multicast)
    // (pairs by default: clusters of 4 halve the traffic again but, at one 225 KB CTA per SM, leave SMs of odd-sized GPCs idle:
    //  measured 1.75x slower; OPL_OZ_CLUSTER=4 selects them)
    int cs = tiles_n % 2 == 0 ? 2 : 1;
    if (const char *env = getenv("OPL_OZ_CLUSTER")) {
        const int want = atoi(env);
        cs = (want >= 4 && tiles_n % 4 == 0) ? 4 : ((want >= 2 && tiles_n % 2 == 0) ? 2 : 1);
    }
    CUtensorMap ta, tb;
    int rc = oz_tensor_map(&ta, a_slices, a_rows_pad, k_pad, OZ_BM / cs);
    if (rc == OPL_OK) rc = oz_tensor_map(&tb, b_slices, b_rows_pad, k_pad, OZ_BN);
    if (rc != OPL_OK) return rc;
    typedef void (*kernel_t)(const CUtensorMap, const CUtensorMap, const OzakiArgs);
    static const kernel_t table[6][3] = {
        {ozaki_gemm_kernel<OZ_EPI_STORE, 1>, ozaki_gemm_kernel<OZ_EPI_STORE, 2>, ozaki_gemm_kernel<OZ_EPI_STORE, 4>},
        {ozaki_gemm_kernel<OZ_EPI_TANH, 1>, ozaki_gemm_kernel<OZ_EPI_TANH, 2>, ozaki_gemm_kernel<OZ_EPI_TANH, 4>},
        {ozaki_gemm_kernel<OZ_EPI_SOFTPLUS, 1>, ozaki_gemm_kernel<OZ_EPI_SOFTPLUS, 2>, ozaki_gemm_kernel<OZ_EPI_SOFTPLUS, 4>},
        {ozaki_gemm_kernel<OZ_EPI_GAUSSIAN, 1>, ozaki_gemm_kernel<OZ_EPI_GAUSSIAN, 2>, ozaki_gemm_kernel<OZ_EPI_GAUSSIAN, 4>},
        {ozaki_gemm_kernel<OZ_EPI_MUL_AUX, 1>, ozaki_gemm_kernel<OZ_EPI_MUL_AUX, 2>, ozaki_gemm_kernel<OZ_EPI_MUL_AUX, 4>},
        {ozaki_gemm_kernel<OZ_EPI_MUL_DTANH, 1>, ozaki_gemm_kernel<OZ_EPI_MUL_DTANH, 2>, ozaki_gemm_kernel<OZ_EPI_MUL_DTANH, 4>}};
    if (args.epi < 0 || args.epi > 5) return fail(OPL_E_ARG, "ozaki GEMM: unknown epilogue %d", args.epi);
    if (args.epi >= OZ_EPI_MUL_AUX && !args.aux_in) return fail(OPL_E_ARG, "ozaki GEMM: epilogue %d needs aux_in", args.epi);
    static bool configured_dev[64] = {};     // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        for (int e = 0; e < 6; ++e)
            for (int c = 0; c < 3; ++c)
                OPL_CUDA(cudaFuncSetAttribute(table[e][c], cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        configured = true;
    }
    static int mma_warps = 0;
    if (mma_warps == 0) {
        const char *env = getenv("OPL_OZ_MMA_WARPS");
        const int want = env ? atoi(env) : OZ_MMA_WARPS;
        mma_warps = std::min(OZ_MMA_WARPS, want >= 4 ? 4 : (want >= 2 ? 2 : 1));
    }
    OzakiArgs largs = args;
    static int late_release = -1;
    if (late_release < 0) {
        const char *env = getenv("OPL_OZ_LATE");
        late_release = (env && atoi(env) != 0) ? 1 : 0;
    }
    largs.late_release = late_release;
    // every issuing warp must own a k-block in EVERY tile: a warp reaches a tile's first k-block having issued into the
    // previous tile, which pins the phase of the accumulator barrier it then waits on (parity waits cannot tell phases
    // two apart)
    const int nkb_regular = args.k_per_split / 128;
    const int nkb_last = (int)ceil_div((int64_t)args.K - (int64_t)(args.splits - 1) * args.k_per_split, 128);
    largs.mma_warps = mma_warps;
    while (largs.mma_warps > 1 && largs.mma_warps > std::min(nkb_regular, nkb_last)) largs.mma_warps >>= 1;
    const kernel_t kernel = table[args.epi][cs == 4 ? 2 : (cs == 2 ? 1 : 0)];
    const int64_t groups = tiles / cs;
    const int64_t nclusters = std::max<int64_t>(1, std::min<int64_t>(groups, sm_count() / cs));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * cs));
    cfg.blockDim = dim3(OZ_THREADS);
    cfg.dynamicSmemBytes = OZ_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cs > 1 ? 1 : 0;
    OPL_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, largs));
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
                                    int32_t epi, int32_t iters, float *slice_ms_out, float *gemm_ms_out) {
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
    // OPL_OZ_TRACE=<file>: per-tile clock64() stamps of CTA 0 in the last iteration (MMA warp: first issue, last commit;
    // epilogue warp 0: wait begin, accumulators ready, accumulators released, finish done), one text line per tile
    long long *trace = nullptr;
    const char *trace_path = getenv("OPL_OZ_TRACE");
    if (trace_path) {
        OPL_CUDA(cudaMalloc(&trace, sizeof(long long) * 32 * 18 * 8));
        OPL_CUDA(cudaMemset(trace, 0, sizeof(long long) * 32 * 18 * 8));
    }
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    int rc = OPL_OK;
    float slice_ms = 0.f, gemm_ms = 0.f;
    for (int it = 0; it <= iters && rc == OPL_OK; ++it) {
        cudaEventRecord(e0, nullptr);
        rc = ozaki_slice_rows(a_dev, M, K, K, sa, ea, false, nullptr, nullptr);
        if (rc == OPL_OK) rc = ozaki_slice_cols_transposed(b_dev, K, N, N, sb, eb, scratch, scratch_len, 0, np, nullptr, nullptr, nullptr);
        cudaEventRecord(e1, nullptr);
        OzakiArgs g = {};
        g.C = splits > 1 ? partial : c_dev;
        g.ldc = N;
        g.ldaux = N;
        g.M = M;
        g.N = N;
        g.K = (int)kp;
        g.k_valid = K;
        g.k_per_split = k_per_split;
        g.splits = splits;
        g.split_stride = splits > 1 ? (int64_t)M * N : 0;
        g.scale_a = ea;
        g.scale_b = eb;
        g.epi = epi;
        g.tparam = 1.0;
        g.trace = (it == iters) ? trace : nullptr;
        if (rc == OPL_OK) rc = launch_ozaki_gemm(sa, mp, sb, np, kp, g, nullptr, nullptr);
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
    if (trace) {
        static long long host[32 * 18 * 8];
        cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(trace_path, "a")) {
            const long long t0 = host[17 * 8];
            fprintf(f, "# M=%d N=%d K=%d epi=%d; cycles since the first MMA stamp.  'mma' tile begin issued | per epilogue warp: wait_begin acc_ready released chunk0_loaded chunk0_transfer_done finish_done\n", M, N, K, epi);
            for (int t = 0; t < 32 && host[(t * 18 + 17) * 8 + 1]; ++t) {
                fprintf(f, "mma %d %lld %lld\n", t, host[(t * 18 + 17) * 8] - t0, host[(t * 18 + 17) * 8 + 1] - t0);
                for (int w = 0; w < 16; ++w) {
                    const long long *e = host + (t * 18 + w) * 8;
                    fprintf(f, "  w%02d %lld %lld %lld %lld %lld %lld\n", w, e[2] - t0, e[3] - t0, e[4] - t0, e[6] - t0, e[7] - t0, e[5] - t0);
                }
            }
            fclose(f);
        }
        cudaFree(trace);
    }
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
