// FP64-accurate GEMM on the INT8 tensor cores (Ozaki splitting) -- interface.  See ozaki.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace opl {

constexpr int OZ_S = 7;          // 7 slices x 7 bits = 49 bits below each row / column maximum
constexpr int OZ_BM = 128;
constexpr int OZ_BN = 64;        // 7 diagonal accumulators x 64 INT32 columns = 448 of 512 TMEM columns
constexpr int OZ_MAX_K = 16384;  // 7 * 16384 * 127^2 < 2^31: INT32 accumulation stays exact

enum OzakiEpilogue { OZ_EPI_STORE = 0, OZ_EPI_TANH = 1, OZ_EPI_SOFTPLUS = 2, OZ_EPI_GAUSSIAN = 3 };

struct OzakiArgs {
    double *C;
    int64_t ldc;
    int M, N, K;              // K = padded contraction length (multiple of 128)
    int k_per_split;          // multiple of 128, <= OZ_MAX_K
    int64_t split_stride;
    const double *scale_a;    // per row of A: power of two >= max|row|
    const double *scale_b;    // per column of B
    const double *bias;
    double *aux_out;
    int64_t ldaux;
    int epi;
    double tparam;
    const double *colmask;
};

// weight of diagonal d = p + q in the recombination: 2^-7(d+2)
__host__ __device__ inline double oz_weight(int d) {
    double w = 1.0 / 16384.0;                 // 2^-14
    for (int i = 0; i < d; ++i) w *= 1.0 / 128.0;
    return w;
}

int64_t ozaki_pad(int64_t v);   // round up to a multiple of 128

// slices of a row-major FP64 [rows x cols] matrix, scaled per ROW: INT8 [S][pad(rows)][pad(cols)], scale[rows]
int ozaki_slice_rows(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                     cudaStream_t stream, int64_t *launches);
// slices of the TRANSPOSE, scaled per COLUMN of src: INT8 [S][pad(cols)][pad(rows)], scale[cols];
// scratch: >= min(ceil(rows/256), 4*SMs) * cols doubles
int ozaki_slice_cols_transposed(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                                double *scratch, int64_t scratch_len, cudaStream_t stream, int64_t *launches);
// C = (A slices) . (B slices)^T with both operands K-major; rows_pad = padded row counts of the slice arrays
int launch_ozaki_gemm(const int8_t *a_slices, int64_t a_rows_pad, const int8_t *b_slices, int64_t b_rows_pad, int64_t k_pad,
                      const OzakiArgs &args, int splits, cudaStream_t stream, int64_t *launches);

}  // namespace opl
