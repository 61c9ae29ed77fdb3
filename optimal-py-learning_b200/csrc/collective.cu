// Gradient all-reduce of the batch-sharded evaluation as ONE kernel over NVLink peer memory.
//
// Each rank's evaluation writes its partial [gradient ; objective] vector straight into a symmetric (peer-mapped)
// buffer; this kernel then (1) announces "my vector of evaluation `seq` is complete" by a release store into every
// peer's flag slot, (2) waits until every rank has announced `seq`, and (3) pulls all `world` vectors over NVLink and
// adds them in RANK ORDER, so every rank obtains the bit-identical sum (the replicated optimizer states must not drift)
// and the result does not depend on arrival order.  One launch, no host round trip, no staging copy: 1.6 MB per peer at
// config 2, i.e. ~11 MB of NVLink reads per GPU at 8 ranks.
//
// On an NVSwitch system with multicast (NVLS) the two-shot form is used instead (allreduce_nvls_kernel): rank r reduces its
// 1/world slice with multimem.ld_reduce (the addition happens in the switch), broadcasts the sums with multimem.st into every
// rank's result buffer, and after a second flag round every rank copies the complete result out -- 2 x 1/world of the vector
// crosses each GPU's links instead of (world-1) x the vector, which is what makes it faster than the pull at 8 ranks.
//
// Buffers are double-buffered by the parity of `seq` by the caller: a rank overwrites buffer (seq+2)&1 only after its
// own all-reduce of seq+1 has seen every peer announce seq+1, which those peers do after finishing their reads of seq.
// Flags are monotonic counters (one slot per writer), never reset.
#include <algorithm>

#include "common.cuh"

namespace opl {

namespace {

constexpr int AR_MAX_WORLD = 16;

struct AllReduceArgs {
    const double *buf[AR_MAX_WORLD];      // peer-mapped address of every rank's vector (own rank: local address)
    uint32_t *flag[AR_MAX_WORLD];         // peer-mapped address of every rank's flag array [AR_MAX_WORLD]
    int world, rank;
    int64_t count;
    uint32_t seq;
    double *out;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_peer2(const double *p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_peer1(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long ar_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until the flag reaches `seq` (monotonic counters; the signed difference handles wrap-around).  Ranks may be far apart
// in host time (one of them logging, checkpointing, timing a CPU baseline), so the wait is patient: fast polls first, then
// sleeping polls, and a trap only after ten minutes -- a dead peer must not hang the GPU for good.
__device__ __forceinline__ void wait_flags(const uint32_t *mine, uint32_t seq) {
    unsigned int spins = 0;
    unsigned long long start = 0;
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
        if (++spins < 4096) continue;
        if (start == 0) start = ar_now_ns();
        __nanosleep(1000);
        if ((spins & 4095u) == 0 && ar_now_ns() - start > 600ull * 1000000000ull) __trap();
    }
}

__global__ void __launch_bounds__(256) allreduce_pull_kernel(const AllReduceArgs p) {
    // (1) my vector was written by earlier kernels of this stream: complete and visible at kernel start
    if (blockIdx.x == 0 && threadIdx.x < p.world) {
        __threadfence_system();
        st_release_sys(p.flag[threadIdx.x] + p.rank, p.seq);
    }
    // (2) every CTA waits for every rank's announcement (flags are monotonic; the difference handles wrap-around)
    if (threadIdx.x < p.world) wait_flags(p.flag[p.rank] + threadIdx.x, p.seq);
    __syncthreads();
    // (3) rank-ordered sum.  All peers' loads of an element pair are issued before the first addition (NVLink latency is
    // ~2 us: dependent load -> add -> load chains made this kernel slower than NCCL); two pairs per thread and step.
    const int64_t pairs = p.count >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool second = j < pairs;
        double2 v[AR_MAX_WORLD], u[AR_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < AR_MAX_WORLD; ++r) {
            if (r < p.world) {
                v[r] = ld_peer2(p.buf[r] + 2 * i);
                if (second) u[r] = ld_peer2(p.buf[r] + 2 * j);
            }
        }
        double2 s = v[0], q = second ? u[0] : make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 1; r < AR_MAX_WORLD; ++r) {
            if (r < p.world) {
                s.x += v[r].x;
                s.y += v[r].y;
                if (second) {
                    q.x += u[r].x;
                    q.y += u[r].y;
                }
            }
        }
        *reinterpret_cast<double2 *>(p.out + 2 * i) = s;
        if (second) *reinterpret_cast<double2 *>(p.out + 2 * j) = q;
    }
    if ((p.count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = ld_peer1(p.buf[0] + p.count - 1);
        for (int r = 1; r < p.world; ++r) s += ld_peer1(p.buf[r] + p.count - 1);
        p.out[p.count - 1] = s;
    }
}

struct NvlsArgs {
    double *mc_in;                        // multicast address of the ranks' input vectors (this parity)
    double *mc_res;                       // multicast address of the ranks' result buffers (this parity)
    const double *res_local;              // this rank's result buffer
    uint32_t *flag[AR_MAX_WORLD];         // peer-mapped flag arrays: [0, 16) "input ready", [16, 32) "slice broadcast"
    int world, rank;
    int64_t count;
    uint32_t seq;
    double *out;
    unsigned int *counter;                // local, zero between launches: CTAs done with their part of the slice
};

__global__ void __launch_bounds__(256) allreduce_nvls_kernel(const NvlsArgs p) {
    // (A) inputs ready everywhere
    if (blockIdx.x == 0 && threadIdx.x < p.world) {
        __threadfence_system();
        st_release_sys(p.flag[threadIdx.x] + p.rank, p.seq);
    }
    if (threadIdx.x < p.world) wait_flags(p.flag[p.rank] + threadIdx.x, p.seq);
    __syncthreads();
    // (B) my slice: sum in the switch, broadcast to every rank's result buffer
    const int64_t per = (p.count + p.world - 1) / p.world;
    const int64_t lo = per * p.rank, hi = lo + per < p.count ? lo + per : p.count;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        double v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(p.mc_in + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p.mc_res + i), "d"(v) : "memory");
    }
    // (C) the last CTA of this rank announces the slice; everybody waits for every rank's slice
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(p.counter, 1u) == gridDim.x - 1) {
            *p.counter = 0;
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(p.flag[r] + AR_MAX_WORLD + p.rank, p.seq);
        }
    }
    if (threadIdx.x < p.world) wait_flags(p.flag[p.rank] + AR_MAX_WORLD + threadIdx.x, p.seq);
    __syncthreads();
    // (D) the complete sum leaves the recycled buffer
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += (int64_t)gridDim.x * blockDim.x)
        p.out[i] = ld_peer1(p.res_local + i);
}

}  // namespace

}  // namespace opl

extern "C" int opl_allreduce_sum_nvls_device(int32_t world, int32_t rank, uint64_t mc_in, uint64_t mc_res, const double *res_local,
                                             const uint64_t *flag_ptrs, int64_t count, uint32_t seq, double *out_dev,
                                             uint32_t *counter_dev, void *cuda_stream) {
    using namespace opl;
    if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world || !mc_in || !mc_res || !res_local || !flag_ptrs ||
        !counter_dev || count < 0 || (count > 0 && !out_dev) || (mc_in & 7) || (mc_res & 7))
        return fail(OPL_E_ARG, "opl_allreduce_sum_nvls_device: bad arguments (world <= %d)", AR_MAX_WORLD);
    if (count == 0) return OPL_OK;
    NvlsArgs a = {};
    for (int r = 0; r < world; ++r) {
        if (!flag_ptrs[r] || (flag_ptrs[r] & 3)) return fail(OPL_E_ARG, "opl_allreduce_sum_nvls_device: flag pointer %d is null or misaligned", r);
        a.flag[r] = reinterpret_cast<uint32_t *>(flag_ptrs[r]);
    }
    a.mc_in = reinterpret_cast<double *>(mc_in);
    a.mc_res = reinterpret_cast<double *>(mc_res);
    a.res_local = res_local;
    a.world = world;
    a.rank = rank;
    a.count = count;
    a.seq = seq;
    a.out = out_dev;
    a.counter = counter_dev;
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(sm_count(), ceil_div(count, 256)));
    allreduce_nvls_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(a);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}

extern "C" int opl_allreduce_sum_device(int32_t world, int32_t rank, const uint64_t *buffer_ptrs, const uint64_t *flag_ptrs,
                                        int64_t count, uint32_t seq, double *out_dev, void *cuda_stream) {
    using namespace opl;
    if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world || !buffer_ptrs || !flag_ptrs || count < 0 ||
        (count > 0 && !out_dev))
        return fail(OPL_E_ARG, "opl_allreduce_sum_device: bad arguments (world <= %d)", AR_MAX_WORLD);
    if (count == 0) return OPL_OK;
    AllReduceArgs a = {};
    for (int r = 0; r < world; ++r) {
        if (!buffer_ptrs[r] || !flag_ptrs[r] || (buffer_ptrs[r] & 15) || (flag_ptrs[r] & 3))
            return fail(OPL_E_ARG, "opl_allreduce_sum_device: peer pointer %d is null or misaligned", r);
        a.buf[r] = reinterpret_cast<const double *>(buffer_ptrs[r]);
        a.flag[r] = reinterpret_cast<uint32_t *>(flag_ptrs[r]);
    }
    if (reinterpret_cast<uintptr_t>(out_dev) & 15) return fail(OPL_E_ARG, "opl_allreduce_sum_device: out must be 16-byte aligned");
    a.world = world;
    a.rank = rank;
    a.count = count;
    a.seq = seq;
    a.out = out_dev;
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(sm_count(), ceil_div(count / 2 + 1, 256)));
    allreduce_pull_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(a);
    OPL_LAUNCH_CHECK();
    return OPL_OK;
}
