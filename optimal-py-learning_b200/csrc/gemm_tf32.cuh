// TF32 tensor-core GEMM family for the fast mode (sm_100a): tcgen05.mma with TMEM accumulators,
// operands staged by TMA into 128B-swizzled shared memory, mbarrier producer/consumer ring.
//
//   warp 0 (one lane)  TMA producer      cp.async.bulk.tensor.2d -> smem stage, complete_tx on full[s]
//   warp 1 (one lane)  MMA issuer        tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) x4 per
//                                        k-block, tcgen05.commit -> empty[s]; last commit -> tmem_full
//   warps 2..5         epilogue          tcgen05.ld 32x32b (warp w owns TMEM lanes 32*(w%4)..), fused
//                                        bias / transfer / derivative, stores
//
// The same three operand layouts as the FP64 family, all on row-major FP32 matrices whose leading
// dimensions are multiples of 32 (the fast-mode workspace pads every layer width):
//   NN  C[M,N] = A[M,K] . B[K,N]      A K-major,  B MN-major   (forward)
//   NT  C[M,N] = A[M,K] . B[N,K]^T    A K-major,  B K-major    (delta . W^T)
//   TN  C[M,N] = A[K,M]^T . B[K,N]    A MN-major, B MN-major   (weight gradient, split-K over rows)
// MN-major operands are legal for kind::tf32 (instruction-descriptor bits 15/16), so no transposed
// copies of activations are ever made.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace opl {

enum Tf32Epilogue {
    TEPI_STORE = 0,     // C = acc (+ bias)
    TEPI_TANH = 1,
    TEPI_SOFTPLUS = 2,  // C = softplus(z), aux_out = sigmoid(z)
    TEPI_GAUSSIAN = 3,  // C = exp(-z^2/v), aux_out = -2 z C / v
    TEPI_MUL_AUX = 4,   // C = acc * aux_in
    TEPI_MUL_DTANH = 5  // C = acc * (1 - aux_in^2)
};

struct Tf32GemmArgs {
    float *C;
    int64_t ldc;
    int M, N, K;             // extents of the (padded) problem; K multiple of 32
    int n_valid;             // columns >= n_valid are written as 0 (layer padding)
    int k_per_split;         // multiple of 32
    int64_t split_stride;    // elements between split partials
    const float *bias;       // length N (padded) or null
    const float *aux_in;
    float *aux_out;
    int64_t ldaux;
    int epi;
    float tparam;
    int round_out;           // store C rounded to TF32 (it is the operand of a later GEMM)
};

constexpr int TF32_BM = 128;
constexpr int TF32_BK = 32;      // 32 tf32 = 128 bytes = one swizzle row
constexpr int TF32_STAGES = 4;

// Encode a 2-D FP32 tensor map; box = box_inner x box_outer elements.  K-major operands use the plain
// 128B swizzle, MN-major operands the 128B swizzle with 32-byte atoms (the only MN-major layout UMMA
// accepts for 32-bit data).
int make_tensor_map_f32(CUtensorMap *map, const float *base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                        uint32_t box_inner, uint32_t box_outer, bool mn_major);

// layout: 0 NN, 1 NT, 2 TN.  A/B leading dimensions in elements.
int launch_gemm_tf32(int layout, const float *A, int64_t lda, const float *B, int64_t ldb, const Tf32GemmArgs &args,
                     int splits, cudaStream_t stream, int64_t *launches);

}  // namespace opl
