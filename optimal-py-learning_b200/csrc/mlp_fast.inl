// Fast mode (OPL_PRECISION_TF32) of the MLP workspace -- included by mlp.cu.
//
// Same schedule as the parity mode, but every matrix lives in FP32 with its column count padded to
// a multiple of 32 (TMA needs 16-byte row pitches; one UMMA k-block is 32 TF32 values) and every
// GEMM is the tcgen05/TMEM/TMA kernel of gemm_tf32.cu.  GEMM operands (X, W, activations, deltas)
// are stored pre-rounded to TF32 with round-to-nearest, so the tensor core's operand truncation is
// exact and the only error left is the unbiased 2^-11 operand rounding (objective / gradient agree
// with the FP64 reference to ~1e-4, gate 1e-3).  Transfer derivatives stay full FP32; row-wise
// softmax / cross entropy / loss reduction run in FP64 on the FP32 logits.  The external interface
// is unchanged: FP64 flat parameters in, FP64 flat gradient + objective out.

struct FastState {
    std::vector<int64_t> pw;         // padded layer widths
    std::vector<int64_t> w_off;      // offset of W_l inside w32 (bias32 occupies [0, pw[1]))
    int64_t w_len = 0;
    float *x32 = nullptr;            // rows x pw[0]
    float *x32_scratch = nullptr;    // activate() on fresh patterns
    float *w32 = nullptr;
    std::vector<float *> act;        // act[l] rows x pw[l], l = 1..L (act[L] = logits)
    std::vector<float *> deriv;
    std::vector<float *> delta;
    float *out_act = nullptr;        // rows x pw[L]
    float *partial = nullptr;
    int64_t partial_len = 0;
};

namespace {

int64_t pad32(int64_t v) { return ceil_div(v, 32) * 32; }

__global__ void pad_convert_kernel(const double *__restrict__ src, int64_t rows, int64_t cols, int64_t ld_dst,
                                   float *__restrict__ dst) {
    const int64_t total = rows * ld_dst;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld_dst, c = i % ld_dst;
        dst[i] = c < cols ? round_tf32((float)src[r * cols + c]) : 0.f;
    }
}

__global__ void unpad_convert_kernel(const float *__restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                                     double *__restrict__ dst) {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (double)src[(i / cols) * ld_src + (i % cols)];
}

// out[i*dout + j] = sum_s partial[s*stride + i*pcols + j]  (FP64 fold of the FP32 split-K partials)
__global__ void reduce_unpad_kernel(const float *__restrict__ partial, int splits, int64_t stride, int64_t pcols,
                                    int64_t din, int64_t dout, double *__restrict__ out) {
    const int64_t total = din * dout;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / dout, j = e % dout;
        double s = 0.0;
        for (int k = 0; k < splits; ++k) s += (double)partial[(int64_t)k * stride + i * pcols + j];
        out[e] = s;
    }
}

// column sums of a rows x cols FP32 matrix (row pitch ld, a multiple of 32 with zero padding), stage 1:
// partial[blockIdx.y][col].  HBM-bound: a thread owns four adjacent columns (one 16-byte load per row) and four rows are in
// flight per thread; block = 32 x 8 threads = 128 columns x 8 row phases.
__global__ void __launch_bounds__(256) colsum32_partial_kernel(const float *__restrict__ D, int64_t rows, int64_t ld, int cols,
                                                               int64_t rows_per_block, double *__restrict__ partial) {
    __shared__ double tile[8][32][5];
    const int c4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int64_t r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (c4 < ld) {
        int64_t r = r0 + threadIdx.y;
        for (; r + 24 < r1; r += 32) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4 *>(D + (r + 8 * u) * ld + c4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s[0] += (double)v[u].x;
                s[1] += (double)v[u].y;
                s[2] += (double)v[u].z;
                s[3] += (double)v[u].w;
            }
        }
        for (; r < r1; r += 8) {
            const float4 v = *reinterpret_cast<const float4 *>(D + r * ld + c4);
            s[0] += (double)v.x;
            s[1] += (double)v.y;
            s[2] += (double)v.z;
            s[3] += (double)v.w;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[threadIdx.y][threadIdx.x][e] = s[e];
    __syncthreads();
    if (threadIdx.y < 4) {                       // row phase y folds component y of every thread column (fixed order)
        const int col = c4 + threadIdx.y;
        if (col < cols) {
            double tot = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tot += tile[k][threadIdx.x][threadIdx.y];
            partial[(int64_t)blockIdx.y * cols + col] = tot;
        }
    }
}

// delta_prev[r][j] = (sum_c delta[r][c] W[j][c]) x slope[r][j] for a NARROW next layer (C <= 16 outputs) in the fast mode: the
// "GEMM" has K = C (padded to 32 for the tensor path: 70 % of the MMA work multiplies zeros and the kernel is bound by its
// 128 x N epilogue), so it is a streaming pass over the rows x H slope / output matrices instead.  One CTA takes a block of
// rows; W (H x C) and the block's delta rows sit in shared memory; thread j walks the rows with 8 loads in flight: global
// traffic is coalesced along j.  FP32 products and sums (the tensor path would round the operands to TF32 first), the output
// is stored TF32-rounded because it is an operand of the next GEMMs; columns past H (layer padding) are written as zeros.
// mode: 0 out = s, 1 out = s * aux (f' stored by the forward GEMM), 2 out = s * (1 - aux^2) (tanh from the activation)
template <int MAXC>      // MAXC = classes padded to a multiple of 4 (12 or 16)
__global__ void __launch_bounds__(256) narrow_backprop32_kernel(const float *__restrict__ delta, int64_t ldd,
                                                                const float *__restrict__ W, int64_t ldw, const float *aux,
                                                                float *out, int64_t ld, int64_t rows, int H, int Hpad, int C,
                                                                int rows_per_block, int mode) {
    extern __shared__ __align__(16) float nsm[];
    constexpr int JT = 4;                   // ADJACENT columns per thread: one 16-byte load / store per pattern, and the block's
                                            // delta row is read once for four outputs
    constexpr int RU = 4;                   // patterns in flight per thread (4 x 16 bytes of loads outstanding)
    float *Ds = nsm;                        // rows_per_block x MAXC (classes >= C are zero), read as float4 broadcasts
    const int64_t nblocks = (rows + rows_per_block - 1) / rows_per_block;
    for (int j0 = 0; j0 < Hpad; j0 += 256 * JT) {
        const int c0 = j0 + JT * threadIdx.x;           // this thread's four columns (Hpad is a multiple of 32)
        const bool live = c0 < Hpad;
        float w[JT][MAXC];                               // their rows of W (registers, reused for every pattern)
#pragma unroll
        for (int q = 0; q < JT; ++q)
#pragma unroll
            for (int c = 0; c < MAXC; ++c) w[q][c] = (c0 + q < H && c < C) ? W[(int64_t)(c0 + q) * ldw + c] : 0.f;
        for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const int64_t r0 = blk * rows_per_block;
            const int nr = (int)min((int64_t)rows_per_block, rows - r0);
            __syncthreads();
            for (int i = threadIdx.x; i < nr * MAXC; i += blockDim.x) {
                const int r = i / MAXC, c = i - r * MAXC;
                Ds[i] = c < C ? delta[(r0 + r) * ldd + c] : 0.f;
            }
            __syncthreads();
            if (!live) continue;
            for (int rb = 0; rb < nr; rb += RU) {
                float4 x[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u)
                    x[u] = (mode != 0 && rb + u < nr) ? *reinterpret_cast<const float4 *>(aux + (r0 + rb + u) * ld + c0)
                                                       : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    if (rb + u >= nr) break;
                    float s_[JT];
#pragma unroll
                    for (int q = 0; q < JT; ++q) s_[q] = 0.f;
                    const float4 *dp = reinterpret_cast<const float4 *>(Ds + (rb + u) * MAXC);
#pragma unroll
                    for (int c4 = 0; c4 < MAXC / 4; ++c4) {
                        const float4 d = dp[c4];
#pragma unroll
                        for (int q = 0; q < JT; ++q) {
                            s_[q] = fmaf(d.x, w[q][4 * c4], s_[q]);
                            s_[q] = fmaf(d.y, w[q][4 * c4 + 1], s_[q]);
                            s_[q] = fmaf(d.z, w[q][4 * c4 + 2], s_[q]);
                            s_[q] = fmaf(d.w, w[q][4 * c4 + 3], s_[q]);
                        }
                    }
                    const float xv[JT] = {x[u].x, x[u].y, x[u].z, x[u].w};
                    float o[JT];
#pragma unroll
                    for (int q = 0; q < JT; ++q) {
                        const float v = mode == 2 ? s_[q] * (1.f - xv[q] * xv[q]) : s_[q] * xv[q];
                        o[q] = c0 + q < H ? round_tf32(v) : 0.f;        // layer padding stays exactly zero
                    }
                    *reinterpret_cast<float4 *>(out + (r0 + rb + u) * ld + c0) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
}

int egrid(int64_t total) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 8 * (int64_t)sm_count()));
}

void fast_free(FastState *f) {
    if (!f) return;
    cudaFree(f->x32);
    cudaFree(f->x32_scratch);
    cudaFree(f->w32);
    for (size_t l = 0; l < f->act.size(); ++l) cudaFree(f->act[l]);
    for (size_t l = 0; l < f->deriv.size(); ++l) cudaFree(f->deriv[l]);
    for (size_t l = 0; l < f->delta.size(); ++l)
        if (l >= f->deriv.size() || f->delta[l] != f->deriv[l]) cudaFree(f->delta[l]);
    cudaFree(f->out_act);
    cudaFree(f->partial);
    delete f;
}

SplitPlan plan_split_tf32(int64_t M, int64_t N, int64_t K) {
    const int64_t bn = (N >= 512 && N % 256 == 0) ? 256 : (N >= 128 ? 128 : (N >= 64 ? 64 : 32));
    const int64_t tiles = ceil_div(M, TF32_BM) * ceil_div(N, bn);
    int64_t want = (2 * (int64_t)sm_count()) / std::max<int64_t>(tiles, 1);
    want = std::max<int64_t>(1, std::min<int64_t>(want, K / 1024));
    SplitPlan plan;
    plan.k_per_split = (int)(ceil_div(ceil_div(K, want), TF32_BK) * TF32_BK);
    plan.splits = (int)ceil_div(std::max<int64_t>(K, 1), plan.k_per_split);
    return plan;
}

int fast_create(opl_mlp *ws) {
    FastState *f = new FastState();
    ws->fast = f;
    const int L = ws->layers;
    f->pw.resize(L + 1);
    for (int l = 0; l <= L; ++l) f->pw[l] = pad32(ws->width[l]);
    f->w_off.resize(L);
    int64_t at = f->pw[1];
    for (int l = 0; l < L; ++l) {
        f->w_off[l] = at;
        at += f->pw[l] * f->pw[l + 1];
    }
    f->w_len = at;
    f->act.assign(L + 1, nullptr);
    f->deriv.assign(L, nullptr);
    f->delta.assign(L, nullptr);
    const int64_t rows = ws->max_rows;
#define OPL_FTRY(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return fail(OPL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
    OPL_FTRY(cudaMalloc(&f->x32, sizeof(float) * (size_t)rows * (size_t)f->pw[0]));
    OPL_FTRY(cudaMalloc(&f->w32, sizeof(float) * (size_t)f->w_len));
    OPL_FTRY(cudaMemset(f->w32, 0, sizeof(float) * (size_t)f->w_len));   // padding rows of every W_l stay zero
    int64_t partial = 0;
    for (int l = 0; l < L; ++l) {
        const size_t bytes = sizeof(float) * (size_t)rows * (size_t)f->pw[l + 1];
        OPL_FTRY(cudaMalloc(&f->act[l + 1], bytes));
        const int t = ws->transfer[l];
        if (l < L - 1 && (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN)) {
            OPL_FTRY(cudaMalloc(&f->deriv[l], bytes));
            f->delta[l] = f->deriv[l];
        } else {
            OPL_FTRY(cudaMalloc(&f->delta[l], bytes));
            OPL_FTRY(cudaMemset(f->delta[l], 0, bytes));   // padded columns of delta_L stay zero
        }
        SplitPlan plan = plan_split_tf32(f->pw[l], f->pw[l + 1], rows);
        partial = std::max<int64_t>(partial, (int64_t)(plan.splits + 1) * f->pw[l] * f->pw[l + 1]);
    }
    f->partial_len = partial;
    OPL_FTRY(cudaMalloc(&f->partial, sizeof(float) * (size_t)partial));
    OPL_FTRY(cudaMalloc(&f->out_act, sizeof(float) * (size_t)rows * (size_t)f->pw[L]));
#undef OPL_FTRY
    return OPL_OK;
}

int fast_convert_x(opl_mlp *ws, const double *x_dev, int64_t rows, float *dst) {
    FastState *f = ws->fast;
    pad_convert_kernel<<<egrid(rows * f->pw[0]), 256, 0, ws->stream>>>(x_dev, rows, ws->width[0], f->pw[0], dst);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    return OPL_OK;
}

int fast_set_x(opl_mlp *ws) { return fast_convert_x(ws, ws->x, ws->rows, ws->fast->x32); }

// Trial point AND its padded FP32 / TF32-rounded copy in one launch (the fast mode's evaluation started with four: the trial
// point, then one pad-convert per weight matrix and the bias): w = x + fl(alpha dir) into the FP64 trial vector (the head and
// the bias read it), every element also rounded to TF32 into its padded slot of w32.
struct FastTrialArgs {
    int layers;
    int64_t bias_len;                 // d1
    int64_t src_off[9], cols[9], dst_off[9], dst_pitch[9];
};
__global__ void __launch_bounds__(256) fast_trial_convert_kernel(int64_t n, const double *__restrict__ x, double alpha,
                                                                 const double *__restrict__ dir, double *__restrict__ w,
                                                                 float *__restrict__ w32, const FastTrialArgs a, int *flag) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = dir ? __dadd_rn(x[i], __dmul_rn(alpha, dir[i])) : x[i];
        w[i] = v;
        const float r = round_tf32((float)v);
        if (i < a.bias_len) {
            w32[i] = r;
        } else {
            int l = 0;
            while (l + 1 < a.layers && i >= a.src_off[l + 1]) ++l;
            const int64_t k = i - a.src_off[l];
            const int64_t row = k / a.cols[l], col = k - row * a.cols[l];
            w32[a.dst_off[l] + row * a.dst_pitch[l] + col] = r;
        }
    }
}

// one launch for the evaluation path (padding rows / columns of w32 were cleared at creation and are never written)
int fast_trial_convert(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev) {
    FastState *f = ws->fast;
    if (ws->layers > 9) return fail(OPL_E_ARG, "fast mode: at most 9 layers");
    FastTrialArgs a = {};
    a.layers = ws->layers;
    a.bias_len = ws->width[1];
    for (int l = 0; l < ws->layers; ++l) {
        a.src_off[l] = ws->w_off[l];
        a.cols[l] = ws->width[l + 1];
        a.dst_off[l] = f->w_off[l];
        a.dst_pitch[l] = f->pw[l + 1];
    }
    const int grid = (int)std::min<int64_t>(ceil_div(ws->n, 256), 4 * (int64_t)sm_count());
    fast_trial_convert_kernel<<<grid, 256, 0, ws->stream>>>(ws->n, x_dev, alpha, dir_dev, ws->w_trial, f->w32, a, ws->flag);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    return OPL_OK;
}

int fast_convert_w(opl_mlp *ws, const double *w) {
    FastState *f = ws->fast;
    // bias as a 1 x d1 matrix, then every weight matrix into its padded slot
    pad_convert_kernel<<<1, 256, 0, ws->stream>>>(w, 1, ws->width[1], f->pw[1], f->w32);
    for (int l = 0; l < ws->layers; ++l)
        pad_convert_kernel<<<egrid(f->pw[l] * f->pw[l + 1]), 256, 0, ws->stream>>>(
            w + ws->w_off[l], ws->width[l], ws->width[l + 1], f->pw[l + 1], f->w32 + f->w_off[l]);
    // rows beyond width[l] of each padded W_l were cleared once at creation and are never written
    OPL_LAUNCH_CHECK();
    ws->launches += ws->layers + 1;
    return OPL_OK;
}

int tepi_for(int transfer) {
    switch (transfer) {
        case OPL_TRANSFER_TANH: return TEPI_TANH;
        case OPL_TRANSFER_SOFTPLUS: return TEPI_SOFTPLUS;
        case OPL_TRANSFER_GAUSSIAN: return TEPI_GAUSSIAN;
        default: return TEPI_STORE;
    }
}

int fast_forward(opl_mlp *ws, const float *x32, int64_t rows, bool keep_deriv, int layer_end = -1) {
    FastState *f = ws->fast;
    const float *in = x32;
    if (layer_end < 0) layer_end = ws->layers;
    for (int l = 0; l < layer_end; ++l) {
        Tf32GemmArgs a = {};
        a.C = f->act[l + 1];
        a.ldc = f->pw[l + 1];
        a.M = (int)rows;
        a.N = (int)f->pw[l + 1];
        a.K = (int)f->pw[l];
        a.n_valid = (int)ws->width[l + 1];
        a.k_per_split = a.K;
        a.bias = l == 0 ? f->w32 : nullptr;
        a.ldaux = f->pw[l + 1];
        a.tparam = (float)ws->tparam[l];
        const bool last = l == ws->layers - 1;
        a.epi = last ? TEPI_STORE : tepi_for(ws->transfer[l]);
        a.aux_out = (!last && keep_deriv) ? f->deriv[l] : nullptr;
        a.round_out = last ? 0 : 1;   // hidden activations feed the next GEMM
        mark(ws, STAGE_FWD * 100 + l);
        int rc = launch_gemm_tf32(0, in, f->pw[l], f->w32 + f->w_off[l], f->pw[l + 1], a, 1, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        if (!last && ws->transfer[l] == OPL_TRANSFER_SOFTMAX) {
            RowArgsT<float> r = {};
            r.Z = f->act[l + 1];
            r.A = f->act[l + 1];
            r.rows = rows;
            r.C = (int)ws->width[l + 1];
            r.ld = f->pw[l + 1];
            r.tid = OPL_TRANSFER_SOFTMAX;
            r.flag = ws->flag;
            r.group = row_group(r.C);
            mark(ws, STAGE_ROWS * 100 + l);
            output_rows_kernel<float><<<row_grid(rows, r.group), 256, 0, ws->stream>>>(r);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
        }
        in = f->act[l + 1];
    }
    return OPL_OK;
}

int fast_output(opl_mlp *ws, const double *y, int64_t rows, bool want_act, bool want_delta, double *obj_out) {
    FastState *f = ws->fast;
    const int L = ws->layers;
    RowArgsT<float> r = {};
    r.Z = f->act[L];
    r.Y = y;
    r.A = want_act ? f->out_act : nullptr;
    r.D = want_delta ? f->delta[L - 1] : nullptr;
    r.rows = rows;
    r.C = (int)ws->width[L];
    r.ld = f->pw[L];
    r.tid = ws->transfer[L - 1];
    r.tparam = ws->tparam[L - 1];
    r.eid = ws->error_id;
    r.ce_div = -(double)ws->norm_rows;
    r.mse_mul = 2.0 / ((double)ws->norm_rows * (double)r.C);
    r.flag = ws->flag;
    r.group = row_group(r.C);
    r.loss_partial = y ? ws->loss_partial : nullptr;
    mark(ws, STAGE_ROWS * 100 + L);
    const int grid = launch_rows(r, ws->loss_blocks, ws->stream);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    if (y && obj_out) {
        mark(ws, STAGE_LOSS * 100);
        loss_final_kernel<<<1, 256, 0, ws->stream>>>(ws->loss_partial, grid, ws->error_id, (double)ws->norm_rows,
                                                      (double)r.C, obj_out, ws->flag);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
    }
    return OPL_OK;
}

int fast_backward(opl_mlp *ws, const float *x32, int64_t rows, double *grad, int layer_begin = -1, bool have_bias_grad = false) {
    FastState *f = ws->fast;
    if (layer_begin < 0) layer_begin = ws->layers - 1;
    for (int l = layer_begin; l >= 0; --l) {
        const int64_t pin = f->pw[l], pout = f->pw[l + 1];
        {   // dW_l = A_l^T delta_l : both operands MN-major, split-K over the patterns
            SplitPlan plan = plan_split_tf32(pin, pout, rows);
            if ((int64_t)plan.splits * pin * pout > f->partial_len) return fail(OPL_E_STATE, "fast split-K scratch too small");
            Tf32GemmArgs a = {};
            a.C = f->partial;
            a.ldc = pout;
            a.M = (int)pin;
            a.N = (int)pout;
            a.K = (int)rows;
            a.n_valid = (int)pout;
            a.k_per_split = plan.k_per_split;
            a.split_stride = pin * pout;
            a.epi = TEPI_STORE;
            mark(ws, STAGE_DW * 100 + l);
            int rc = launch_gemm_tf32(2, l == 0 ? x32 : f->act[l], pin, f->delta[l], pout, a, plan.splits, ws->stream,
                                      &ws->launches);
            if (rc != OPL_OK) return rc;
            mark(ws, STAGE_DW_REDUCE * 100 + l);
            reduce_unpad_kernel<<<egrid(ws->width[l] * ws->width[l + 1]), 256, 0, ws->stream>>>(
                f->partial, plan.splits, pin * pout, pout, ws->width[l], ws->width[l + 1], grad + ws->w_off[l]);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
        }
        if (l == 0) break;
        // delta_{l-1} = (delta_l W_l^T) o f'_{l-1}
        if (ws->width[l + 1] <= 16 && ws->transfer[l - 1] != OPL_TRANSFER_SOFTMAX) {
            // narrow next layer: a streaming pass, not a GEMM (see narrow_backprop32_kernel)
            const int t = ws->transfer[l - 1];
            const int mode = (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN) ? 1 : (t == OPL_TRANSFER_TANH ? 2 : 0);
            const float *aux = mode == 1 ? f->deriv[l - 1] : (mode == 2 ? f->act[l] : nullptr);
            const int H = (int)ws->width[l], C = (int)ws->width[l + 1], rpb = 64;
            const int64_t grid = std::min<int64_t>(ceil_div(rows, rpb), 4 * (int64_t)sm_count());
            mark(ws, STAGE_DA * 100 + l);
            if (C <= 12)
                narrow_backprop32_kernel<12><<<(unsigned)grid, 256, sizeof(float) * rpb * 12, ws->stream>>>(
                    f->delta[l], pout, f->w32 + f->w_off[l], pout, aux, f->delta[l - 1], pin, rows, H, (int)pin, C, rpb, mode);
            else
                narrow_backprop32_kernel<16><<<(unsigned)grid, 256, sizeof(float) * rpb * 16, ws->stream>>>(
                    f->delta[l], pout, f->w32 + f->w_off[l], pout, aux, f->delta[l - 1], pin, rows, H, (int)pin, C, rpb, mode);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
            continue;
        }
        Tf32GemmArgs a = {};
        a.C = f->delta[l - 1];
        a.ldc = pin;
        a.M = (int)rows;
        a.N = (int)pin;
        a.K = (int)pout;
        a.n_valid = (int)ws->width[l];
        a.k_per_split = a.K;
        a.ldaux = pin;
        a.round_out = 1;
        switch (ws->transfer[l - 1]) {
            case OPL_TRANSFER_SOFTPLUS:
            case OPL_TRANSFER_GAUSSIAN:
                a.epi = TEPI_MUL_AUX;
                a.aux_in = f->deriv[l - 1];
                break;
            case OPL_TRANSFER_TANH:
                a.epi = TEPI_MUL_DTANH;
                a.aux_in = f->act[l];
                break;
            default:
                a.epi = TEPI_STORE;
        }
        mark(ws, STAGE_DA * 100 + l);
        int rc = launch_gemm_tf32(1, f->delta[l], pout, f->w32 + f->w_off[l], pout, a, 1, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        if (ws->transfer[l - 1] == OPL_TRANSFER_SOFTMAX) {
            const int group = row_group((int)ws->width[l]);
            mark(ws, STAGE_ROWS * 100 + l - 1);
            softmax_backward_rows_kernel<float><<<row_grid(rows, group), 256, 0, ws->stream>>>(
                f->delta[l - 1], f->act[l], rows, pin, (int)ws->width[l], group);
            OPL_LAUNCH_CHECK();
            ++ws->launches;
        }
    }
    if (!have_bias_grad) {   // db = column sums of delta_0
        const int cols = (int)ws->width[1];
        const int64_t gx = ceil_div(f->pw[1], 128);
        int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 64), (8 * (int64_t)sm_count()) / gx + 1));
        if (nblk * cols > ws->partials_len) nblk = std::max<int64_t>(1, ws->partials_len / cols);
        const int64_t rpb = ceil_div(rows, nblk);
        nblk = ceil_div(rows, rpb);
        dim3 grid((unsigned)gx, (unsigned)nblk);
        mark(ws, STAGE_BIAS * 100);
        colsum32_partial_kernel<<<grid, dim3(32, 8), 0, ws->stream>>>(f->delta[0], rows, f->pw[1], cols, rpb, ws->partials);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
        int rc = launch_reduce_partials(ws->partials, (int)nblk, cols, cols, grad, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
    }
    return OPL_OK;
}

// The fused narrow head on FP32 (TF32-rounded) activations: same kernel as the parity mode, FP64 arithmetic inside (the
// head is streaming, not tensor work), FP32 loads / TF32-rounded stores.  When the head's delta_prev is delta_0 (two-layer
// networks) its per-CTA column sums are the bias gradient and the separate column-sum pass disappears.
// ---------------------------------------------------------------------------------------
// The fused narrow head of the fast mode on the TF32 tensor path.  Same streaming structure as narrow_head_kernel (one
// pass over the hidden activations: logits, output transfer + error, delta_L, dW += A^T delta_L, delta_prev =
// (delta_L W^T) o slope), but the three small products are mma.sync.m16n8k8 TF32 with FP32 accumulation instead of FP64
// DMMA: the operands are TF32-rounded FP32 already, a 16-pattern step is exactly one M = 16 tile, and the FP64 pipe (one
// DMMA per 3.9 cycles per SM, shared with every FP64 instruction) that bounds the FP64 head is not touched.  R = 16
// patterns per step, 8 warps, 51-56 KB of shared memory: three to four CTAs per SM.
//   logits  Z[16 x 16]   = As[16 x H] . W[H x 16]          warp -> (class tile, K quarter); quarters folded in fixed order
//   dA      D[16 x H]    = dL[16 x 16] . W^T[16 x H]        warp -> hidden tiles w, w+8, ..; epilogue x slope, TF32 round, store
//   dW      G[H x 16]   += As^T[H x 16] . dL[16 x 16]       warp -> 16-row hidden tiles w, w+8; FP32 accumulators in registers
// W is kept once, transposed (Wt[class][hidden], pitch H+4); As pitch H+4; both conflict-free for the logits fragments.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int FH_R = 16, FH_THREADS = 256, FH_LDZ = 24;
size_t fast_head_smem(int H) {
    return sizeof(float) * ((size_t)16 * (H + 4) + 16 * FH_LDZ + 4 * 256 + 2 * (size_t)FH_R * (H + 4));
}

__global__ void __launch_bounds__(FH_THREADS, 3) narrow_head_tf32_kernel(const HeadArgsT<float> p) {
    extern __shared__ float fsm[];
    __shared__ double scratch[33];
    const int H = p.H, C = p.C;
    const int ldT = H + 4, ldA = H + 4;
    float *Wt = fsm;                          // 16 x ldT, classes >= C are zero rows
    float *Zs = Wt + 16 * ldT;                // 16 x 24: delta_L (TF32-rounded)
    float *Zp = Zs + 16 * FH_LDZ;             // 4 x 16 x 16: partial logits of the four K quarters
    float *As0 = Zp + 4 * 256;                // 2 x 16 x ldA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bool softmax = p.tid == OPL_TRANSFER_SOFTMAX;
    const bool backward = p.delta_prev != nullptr;
    const int h4 = H >> 2;                    // 16-byte chunks per row
    const int64_t nblocks = (p.rows + FH_R - 1) / FH_R;
    auto fetch = [&](int64_t blk, int buf) {   // a warp per row, lanes along the row (as narrow_head_kernel)
        const int64_t r0 = blk * FH_R;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(As0) + (unsigned)(buf * FH_R * ldA) * 4u;
        for (int r = warp; r < FH_R; r += FH_THREADS / 32) {
            const bool ok = r0 + r < p.rows;
            const float *src = p.A + (ok ? (r0 + r) * p.ld : 0) + 4 * lane;
            unsigned d = dst + (unsigned)(r * ldA + 4 * lane) * 4u;
            for (int c4 = lane; c4 < h4; c4 += 32, src += 128, d += 512) cp_async16_s(d, ok ? src : p.A, ok);
        }
        cp_async_commit();
    };
    if ((int64_t)blockIdx.x < nblocks) fetch(blockIdx.x, 0);
    for (int i = threadIdx.x; i < H * 16; i += FH_THREADS) {
        const int h = i >> 4, c = i & 15;
        Wt[c * ldT + h] = c < C ? round_tf32((float)p.W[h * C + c]) : 0.f;
    }
    RowArgs ra = {};
    ra.eid = p.eid;
    ra.ce_div = p.ce_div;
    ra.mse_mul = p.mse_mul;
    double loss = 0.0;
    bool bad = false;
    float dw[2][2][4];                        // [hidden 16-tile j][class tile][fragment]
    float csum[4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) dw[j][n][e] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) csum[i][0] = csum[i][1] = 0.f;
    const int ksteps = H >> 3;
    const bool two_class_steps = C > 8;
    int buf = 0;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, buf ^= 1) {
        const int64_t r0 = blk * FH_R;
        const int nr = (int)min((int64_t)FH_R, p.rows - r0);
        const float *As = As0 + (size_t)buf * FH_R * ldA;
        // slope values of this step's delta_prev fragments and this thread's target value: requested now
        float2 x[4][2];
        if (backward && p.mode != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int col = 8 * (warp + 8 * i) + 2 * t;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int row = g + 8 * hf;
                    x[i][hf] = (col < H && row < nr) ? *reinterpret_cast<const float2 *>(p.aux + (r0 + row) * p.ld + col)
                                                     : make_float2(1.f, 1.f);
                }
            }
        }
        double y_mine = 0.0;
        {
            const int r = warp * 2 + (lane >> 4), cl = lane & 15;
            if (r < nr && cl < C) y_mine = p.Y[(r0 + r) * C + cl];
        }
        cp_async_wait<0>();
        __syncthreads();                       // tile `buf` landed; the previous step is done with the other buffer, Zs and Zp
        if (blk + gridDim.x < nblocks) fetch(blk + gridDim.x, buf ^ 1);
        // ---- logits: warp -> (class tile nt, K quarter kq); k-steps kq, kq + 4, .. ----
        {
            const int nt = warp & 1, kq = warp >> 1;
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            const float *a_lo = As + (size_t)g * ldA + t, *a_hi = As + (size_t)(g + 8) * ldA + t;
            const float *b = Wt + (size_t)(8 * nt + g) * ldT + t;
            int ks = kq;
            for (; ks + 4 < ksteps; ks += 8) {
                mma_tf32_16x8x8(c0, __float_as_uint(a_lo[8 * ks]), __float_as_uint(a_hi[8 * ks]), __float_as_uint(a_lo[8 * ks + 4]),
                                __float_as_uint(a_hi[8 * ks + 4]), __float_as_uint(b[8 * ks]), __float_as_uint(b[8 * ks + 4]));
                const int k2 = ks + 4;
                mma_tf32_16x8x8(c1, __float_as_uint(a_lo[8 * k2]), __float_as_uint(a_hi[8 * k2]), __float_as_uint(a_lo[8 * k2 + 4]),
                                __float_as_uint(a_hi[8 * k2 + 4]), __float_as_uint(b[8 * k2]), __float_as_uint(b[8 * k2 + 4]));
            }
            if (ks < ksteps)
                mma_tf32_16x8x8(c0, __float_as_uint(a_lo[8 * ks]), __float_as_uint(a_hi[8 * ks]), __float_as_uint(a_lo[8 * ks + 4]),
                                __float_as_uint(a_hi[8 * ks + 4]), __float_as_uint(b[8 * ks]), __float_as_uint(b[8 * ks + 4]));
            float *zp = Zp + kq * 256 + g * 16 + 8 * nt + 2 * t;
            *reinterpret_cast<float2 *>(zp) = make_float2(c0[0] + c1[0], c0[1] + c1[1]);
            *reinterpret_cast<float2 *>(zp + 8 * 16) = make_float2(c0[2] + c1[2], c0[3] + c1[3]);
        }
        __syncthreads();
        // ---- output transfer + error + delta_L: one pattern per half-warp, lane = class ----
        {
            const int hw = lane >> 4, cl = lane & 15;
            const int r = warp * 2 + hw;
            const bool live = r < nr && cl < C;
            const float zf = (Zp[r * 16 + cl] + Zp[256 + r * 16 + cl]) + (Zp[512 + r * 16 + cl] + Zp[768 + r * 16 + cl]);
            float d_out = 0.f;
            if (softmax && p.eid == OPL_ERROR_CROSS_ENTROPY) {
                // the classification head in FP32: same formulas, FLT_MAX where the FP64 path has DBL_MAX
                float mx = live ? zf : -INFINITY;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float e = live ? expf(zf - mx) : 0.f;
                float den = e;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
                const float a = e / den;
                float gq = 0.f;
                if (live) {
                    const float y = (float)y_mine;
                    float lg = logf(a);
                    lg = isnan(lg) ? 0.f : (isinf(lg) ? (lg > 0 ? 3.4028234664e38f : -3.4028234664e38f) : lg);
                    loss += (double)(lg * y);
                    float q = y / a;
                    q = isnan(q) ? 0.f : (isinf(q) ? (q > 0 ? 3.4028234664e38f : -3.4028234664e38f) : q);
                    gq = q / (float)p.ce_div;
                }
                float tt = live ? gq * a : 0.f;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
                d_out = live ? softmax_delta(a, gq, tt) : 0.f;
            } else {
                const double z = (double)zf;
                double a_out;
                if (softmax) {
                    const double mx = group_max(live ? z : -INFINITY, 16);
                    const double e = live ? exp(z - mx) : 0.0;
                    a_out = e / group_sum(e, 16);
                } else {
                    a_out = live ? transfer_value(p.tid, p.tparam, z) : 0.0;
                }
                double gq = 0.0;
                if (live) gq = error_term(ra, a_out, y_mine, loss, bad);
                double d;
                if (softmax) {
                    const double tt = group_sum(live ? gq * a_out : 0.0, 16);
                    d = softmax_delta(a_out, gq, tt);
                } else {
                    d = live ? gq * transfer_slope(p.tid, p.tparam, z, a_out) : 0.0;
                }
                d_out = live ? (float)d : 0.f;
            }
            if (backward) Zs[r * FH_LDZ + cl] = round_tf32(d_out);
        }
        __syncthreads();
        if (!backward) continue;
        // ---- delta_prev = (delta_L W^T) o slope: warp -> hidden tiles w, w + 8, .. ----
        {
            const uint32_t a0 = __float_as_uint(Zs[g * FH_LDZ + t]), a1 = __float_as_uint(Zs[(g + 8) * FH_LDZ + t]);
            const uint32_t a2 = __float_as_uint(Zs[g * FH_LDZ + t + 4]), a3 = __float_as_uint(Zs[(g + 8) * FH_LDZ + t + 4]);
            uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
            if (two_class_steps) {
                e0 = __float_as_uint(Zs[g * FH_LDZ + 8 + t]);
                e1 = __float_as_uint(Zs[(g + 8) * FH_LDZ + 8 + t]);
                e2 = __float_as_uint(Zs[g * FH_LDZ + 12 + t]);
                e3 = __float_as_uint(Zs[(g + 8) * FH_LDZ + 12 + t]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nt = warp + 8 * i;
                if (8 * nt >= H) break;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                const float *b = Wt + (size_t)t * ldT + 8 * nt + g;
                mma_tf32_16x8x8(c, a0, a1, a2, a3, __float_as_uint(b[0]), __float_as_uint(b[4 * ldT]));
                if (two_class_steps) mma_tf32_16x8x8(c, e0, e1, e2, e3, __float_as_uint(b[8 * ldT]), __float_as_uint(b[12 * ldT]));
                const int col = 8 * nt + 2 * t;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int row = g + 8 * hf;
                    if (row < nr) {
                        float o0 = c[2 * hf], o1 = c[2 * hf + 1];
                        if (p.mode == 1) {
                            o0 *= x[i][hf].x;
                            o1 *= x[i][hf].y;
                        } else if (p.mode == 2) {
                            o0 *= 1.f - x[i][hf].x * x[i][hf].x;
                            o1 *= 1.f - x[i][hf].y * x[i][hf].y;
                        }
                        const float2 st = make_float2(round_tf32(o0), round_tf32(o1));
                        *reinterpret_cast<float2 *>(p.delta_prev + (r0 + row) * p.ld + col) = st;
                        csum[i][0] += st.x;
                        csum[i][1] += st.y;
                    }
                }
            }
        }
        // ---- dW += A^T delta_L: warp -> hidden 16-tiles w, w + 8 ----
        {
            uint32_t b[2][2][2];              // [row step][class tile][b0, b1]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    b[ks][n][0] = __float_as_uint(Zs[(8 * ks + t) * FH_LDZ + 8 * n + g]);
                    b[ks][n][1] = __float_as_uint(Zs[(8 * ks + t + 4) * FH_LDZ + 8 * n + g]);
                }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int m0 = 16 * (warp + 8 * j);
                if (m0 >= H) break;
                const bool hi_ok = m0 + 8 < H;           // H is a multiple of 8: the upper half of the last tile may be absent
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const float *a = As + (size_t)(8 * ks + t) * ldA + m0 + g;
                    const uint32_t a0 = __float_as_uint(a[0]), a1 = hi_ok ? __float_as_uint(a[8]) : 0u;
                    const uint32_t a2 = __float_as_uint(a[4 * ldA]), a3 = hi_ok ? __float_as_uint(a[4 * ldA + 8]) : 0u;
                    mma_tf32_16x8x8(dw[j][0], a0, a1, a2, a3, b[ks][0][0], b[ks][0][1]);
                    if (two_class_steps) mma_tf32_16x8x8(dw[j][1], a0, a1, a2, a3, b[ks][1][0], b[ks][1][1]);
                }
            }
        }
        __syncthreads();                       // Zs / Zp / this A buffer are free again
    }
    cp_async_wait<0>();
    if (bad) *p.flag = 1;
    if (backward) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m0 = 16 * (warp + 8 * j);
            if (m0 >= H) break;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int m = m0 + g + 8 * (e >> 1), c = 8 * n + 2 * t + (e & 1);
                    if (m < H && c < C) p.dw_partial[(size_t)blockIdx.x * H * C + (size_t)m * C + c] = (double)dw[j][n][e];
                }
        }
        if (p.colsum_partial) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nt = warp + 8 * i;
                if (8 * nt >= H) break;
                float s0 = csum[i][0], s1 = csum[i][1];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (g == 0) {
                    p.colsum_partial[(size_t)blockIdx.x * H + 8 * nt + 2 * t] = (double)s0;
                    p.colsum_partial[(size_t)blockIdx.x * H + 8 * nt + 2 * t + 1] = (double)s1;
                }
            }
        }
    }
    loss = block_sum(loss, scratch);
    if (threadIdx.x == 0) p.loss_partial[blockIdx.x] = loss;
}

bool fast_head_ok(const opl_mlp *ws) {
    const int L = ws->layers;
    if (L < 2 || !ws->fused_head) return false;
    const int64_t H = ws->width[L - 1], C = ws->width[L];
    if (C > 16 || H > 256 || (H & 7) || H < 32) return false;
    return ws->transfer[L - 2] != OPL_TRANSFER_SOFTMAX;
}

int fast_fused_head(opl_mlp *ws, bool want_grad, double *grad, double *obj_out) {
    FastState *f = ws->fast;
    const int L = ws->layers;
    const int H = (int)ws->width[L - 1], C = (int)ws->width[L];
    constexpr int R = 16;
    const size_t smem = fast_head_smem(H);
    int grid = (int)std::min<int64_t>(ceil_div(ws->rows, R), 3 * (int64_t)sm_count());
    grid = std::min(grid, ws->loss_blocks);
    if (want_grad) grid = (int)std::min<int64_t>(grid, ws->partials_len / ((int64_t)H * (C + 1)));
    if (grid < 1) return fail(OPL_E_STATE, "fused head (fast mode): scratch too small");
    HeadArgsT<float> a = {};
    a.A = f->act[L - 1];
    a.ld = f->pw[L - 1];
    a.W = ws->w_trial + ws->w_off[L - 1];
    a.Y = ws->y;
    a.rows = ws->rows;
    a.H = H;
    a.C = C;
    a.R = R;
    a.tid = ws->transfer[L - 1];
    a.tparam = ws->tparam[L - 1];
    a.eid = ws->error_id;
    a.ce_div = -(double)ws->norm_rows;
    a.mse_mul = 2.0 / ((double)ws->norm_rows * (double)C);
    a.flag = ws->flag;
    a.loss_partial = ws->loss_partial;
    const bool bias_here = want_grad && L == 2;
    if (want_grad) {
        const int t = ws->transfer[L - 2];
        a.mode = (t == OPL_TRANSFER_SOFTPLUS || t == OPL_TRANSFER_GAUSSIAN) ? 1 : (t == OPL_TRANSFER_TANH ? 2 : 0);
        a.aux = a.mode == 1 ? f->deriv[L - 2] : (a.mode == 2 ? f->act[L - 1] : nullptr);
        a.delta_prev = f->delta[L - 2];
        a.dw_partial = ws->partials;
        a.colsum_partial = bias_here ? ws->partials + (size_t)grid * H * C : nullptr;
    }
    static bool configured_dev[64] = {};     // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)fast_head_smem(256)));
        configured = true;
    }
    mark(ws, STAGE_HEAD * 100 + L - 1);
    narrow_head_tf32_kernel<<<grid, FH_THREADS, smem, ws->stream>>>(a);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    if (want_grad && grid >= 16) {
        // one launch folds the per-CTA dW partials, the bias-gradient column sums and the loss partials
        const int64_t count = (int64_t)H * C, count2 = bias_here ? H : 0;
        const int64_t fgrid = std::min<int64_t>(ceil_div((count + count2) * 32, 256), 4 * (int64_t)sm_count());
        cudaStream_t fold_stream = ws->stream;
        if (L == 2 && bias_here && ws->side_fold && !ws->profiling) {
            // two-layer network: only the dW_0 GEMM follows, with FP32 split-K partials of its own (f->partial): the folds run
            // beside it on the workspace's side stream and rejoin before the evaluation returns (as in the parity mode)
            if (!ws->side) {
                OPL_CUDA(cudaStreamCreateWithFlags(&ws->side, cudaStreamNonBlocking));
                OPL_CUDA(cudaEventCreateWithFlags(&ws->ev_head, cudaEventDisableTiming));
                OPL_CUDA(cudaEventCreateWithFlags(&ws->ev_fold, cudaEventDisableTiming));
            }
            OPL_CUDA(cudaEventRecord(ws->ev_head, ws->stream));
            OPL_CUDA(cudaStreamWaitEvent(ws->side, ws->ev_head, 0));
            fold_stream = ws->side;
        } else {
            mark(ws, STAGE_DW_REDUCE * 100 + L - 1);
        }
        head_fold_kernel<<<(unsigned)fgrid, 256, 0, fold_stream>>>(ws->partials, grid, count, count, grad + ws->w_off[L - 1],
                                                                ws->loss_partial, ws->error_id, (double)ws->norm_rows,
                                                                (double)C, obj_out, ws->flag, a.colsum_partial, (int64_t)H,
                                                                count2, grad);
        OPL_LAUNCH_CHECK();
        ++ws->launches;
        if (fold_stream != ws->stream) {
            OPL_CUDA(cudaEventRecord(ws->ev_fold, fold_stream));
            ws->fold_pending = true;
        }
        return OPL_OK;
    }
    mark(ws, STAGE_LOSS * 100);
    loss_final_kernel<<<1, 256, 0, ws->stream>>>(ws->loss_partial, grid, ws->error_id, (double)ws->norm_rows, (double)C, obj_out, ws->flag);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    if (want_grad) {
        mark(ws, STAGE_DW_REDUCE * 100 + L - 1);
        const int64_t count = (int64_t)H * C;
        int rc = launch_reduce_partials(ws->partials, grid, count, count, grad + ws->w_off[L - 1], ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        if (bias_here) {
            mark(ws, STAGE_BIAS * 100);
            rc = launch_reduce_partials(a.colsum_partial, grid, H, H, grad, ws->stream, &ws->launches);
            if (rc != OPL_OK) return rc;
        }
    }
    return OPL_OK;
}

int fast_evaluate(opl_mlp *ws, bool want_grad, double *grad_out_dev) {
    FastState *f = ws->fast;
    int rc = OPL_OK;
    if (fast_head_ok(ws)) {
        rc = fast_forward(ws, f->x32, ws->rows, want_grad, ws->layers - 1);
        if (rc != OPL_OK) return rc;
        rc = fast_fused_head(ws, want_grad, grad_out_dev, grad_out_dev + ws->n);
        if (rc != OPL_OK) return rc;
        if (want_grad) rc = fast_backward(ws, f->x32, ws->rows, grad_out_dev, ws->layers - 2, ws->layers == 2);
        return rc;
    }
    rc = fast_forward(ws, f->x32, ws->rows, want_grad);
    if (rc != OPL_OK) return rc;
    rc = fast_output(ws, ws->y, ws->rows, false, want_grad, grad_out_dev + ws->n);
    if (rc != OPL_OK) return rc;
    if (want_grad) rc = fast_backward(ws, f->x32, ws->rows, grad_out_dev);
    return rc;
}

// activate: x (FP64 device, rows x d0) -> out (FP64 device, rows x dL)
int fast_activate(opl_mlp *ws, const double *params_dev, const double *x_dev, int64_t rows, double *out_dev) {
    FastState *f = ws->fast;
    if (!f->x32_scratch)
        OPL_CUDA(cudaMalloc(&f->x32_scratch, sizeof(float) * (size_t)ws->max_rows * (size_t)f->pw[0]));
    int rc = fast_convert_x(ws, x_dev, rows, f->x32_scratch);
    if (rc == OPL_OK) rc = fast_convert_w(ws, params_dev);
    if (rc == OPL_OK) rc = fast_forward(ws, f->x32_scratch, rows, false);
    if (rc == OPL_OK) rc = fast_output(ws, nullptr, rows, true, false, nullptr);
    if (rc != OPL_OK) return rc;
    const int L = ws->layers;
    unpad_convert_kernel<<<egrid(rows * ws->width[L]), 256, 0, ws->stream>>>(f->out_act, rows, ws->width[L], f->pw[L], out_dev);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    return OPL_OK;
}

}  // namespace
