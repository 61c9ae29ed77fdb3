// FP64-accurate GEMM on the INT8 tensor cores (Ozaki splitting) -- interface.  See ozaki.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace opl {

constexpr int OZ_S = 6;          // 6 signed radix-256 digits = 48 bits below each row / column scale
constexpr int OZ_BM = 128;
constexpr int OZ_BN = 64;        // 6 diagonal accumulators x 64 INT32 columns = 384 of 512 TMEM columns
constexpr int OZ_MAX_K = 16384;  // 6 * 16384 * 128^2 < 2^31: INT32 accumulation stays exact

enum OzakiEpilogue {
    OZ_EPI_STORE = 0, OZ_EPI_TANH = 1, OZ_EPI_SOFTPLUS = 2, OZ_EPI_GAUSSIAN = 3,   // forward: transfer (+ derivative into aux_out)
    OZ_EPI_MUL_AUX = 4,      // back-propagation: C = (A.B) o aux_in            (aux_in = f' stored by the forward pass; may alias C)
    OZ_EPI_MUL_DTANH = 5     // back-propagation: C = (A.B) o (1 - aux_in^2)    (aux_in = the tanh activation)
};

struct OzakiArgs {
    double *C;
    int64_t ldc;
    int M, N, K;              // K = padded contraction length (multiple of 128)
    int k_valid;              // true contraction length (0: K).  The zero padding behind it is not multiplied: the last
                              // k-block issues only the K = 32 steps that hold data (784 -> 25 instead of 28 steps)
    int k_per_split;          // multiple of 128, <= OZ_MAX_K
    int splits;
    int64_t split_stride;
    const double *scale_a;    // per row of A: power of two, |a| <= 0.498 * scale
    const double *scale_b;    // per column of B
    const double *bias;
    double *aux_out;
    int64_t ldaux;
    const double *aux_in;     // OZ_EPI_MUL_*: same leading dimension as C
    int epi;
    double tparam;
    const double *colmask;
    int mma_warps;            // 1, 2 or 4 MMA-issuing warps taking k-blocks in turn; set by launch_ozaki_gemm
    long long *trace;         // tools only (OPL_OZ_TRACE): CTA 0 writes 8 clock64() stamps per tile, see ozaki.cu
    int late_release;         // tools only (OPL_OZ_LATE=1): release the accumulators after the finish phase instead of after the drain
};

int64_t ozaki_pad(int64_t v);   // round up to a multiple of 128

// digits of a row-major FP64 [rows x cols] matrix, scaled per ROW: INT8 [S][pad(rows)][pad(cols)], scale[rows].
// Padding must already be zero (ozaki_slice_* never write it); pass zero_fill to have it cleared here.
int ozaki_slice_rows(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                     bool zero_fill, cudaStream_t stream, int64_t *launches);
// digits of the TRANSPOSE, scaled per COLUMN of src: INT8 [S][m_pad][pad(rows)], scale[cols];
// scratch: >= min(ceil(rows/256), 4*SMs) * cols doubles.
// Output row of src column c is c + row_off (scale[c + row_off]); m_pad = padded row count of the digit array
// (0: pad(cols + row_off)).  colmax_raw (optional): max |src[:, c]| already gathered by the producer of src
// (atomicMax on the bit patterns, NaN pattern for non-finite entries); saves the column-maximum pass.
int ozaki_slice_cols_transposed(const double *src, int64_t rows, int64_t cols, int64_t ld, int8_t *slices, double *scale,
                                double *scratch, int64_t scratch_len, int64_t row_off, int64_t m_pad, const double *colmax_raw,
                                cudaStream_t stream, int64_t *launches);
// trial point w = x + fl(alpha dir) (whole flat vector, n entries) AND the digits of W_0^T (W_0 = w[off : off + rows cols],
// rows x cols row-major; output as ozaki_slice_cols_transposed) in ONE cooperative launch; also resets the evaluation's
// domain flag and clears up to two small accumulator arrays (zero_a, zero_b; may be null).  *done = false: the grid does not fit, nothing was enqueued.
int ozaki_trial_w0_digits(const double *x, const double *dir, double alpha, double *w, int64_t n, int64_t off, int64_t rows,
                          int64_t cols, int8_t *slices, double *scale, double *scratch, int64_t scratch_len, int64_t m_pad,
                          int *flag, double *zero_a, int zero_a_len, double *zero_b, int zero_b_len, cudaStream_t stream,
                          int64_t *launches, bool *done);
// digits of delta_prev^T = ((deltaL W^T) o slope)^T formed on the fly from deltaL (rows x C, C <= 16), W (cols x C) and the slope
// matrix aux (rows x cols, pitch ld; mode 0 none, 1 aux = f', 2 aux = tanh activation); classmax[c] = max_r |deltaL[r][c]| (raw
// bit patterns as gathered by atomicMax).  Same output layout as ozaki_slice_cols_transposed: INT8 [S][m_pad][pad(rows)], scale[cols].
int ozaki_slice_delta_prev(const double *deltaL, int C, const double *W, const double *aux, int mode, int64_t rows, int64_t cols,
                           int64_t ld, const double *classmax, int8_t *slices, double *scale, int64_t m_pad,
                           cudaStream_t stream, int64_t *launches, bool scale_cleared = false);
// row 0 of a transposed digit array [S][m_pad][pad(rows)] := a row of ones (scale[0] = 4): the bias-gradient row
int ozaki_ones_row(int8_t *slices, double *scale, int64_t rows, int64_t m_pad, cudaStream_t stream, int64_t *launches);
// FP64 fold of split-K partials is done by the caller (launch_reduce_partials)
// C = (A digits) . (B digits)^T with both operands K-major; rows_pad = padded row counts of the digit arrays
int launch_ozaki_gemm(const int8_t *a_slices, int64_t a_rows_pad, const int8_t *b_slices, int64_t b_rows_pad, int64_t k_pad,
                      const OzakiArgs &args, cudaStream_t stream, int64_t *launches);

}  // namespace opl
