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
                                                      (double)r.C, obj_out);
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
    const size_t smem = head_smem(H, R, sizeof(float));
    int grid = (int)std::min<int64_t>(ceil_div(ws->rows, R), 2 * (int64_t)sm_count());
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
    static bool configured = false;
    if (!configured) {
        OPL_CUDA(cudaFuncSetAttribute(narrow_head_kernel<float, 4, 16, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)head_smem(256, R, sizeof(float))));
        configured = true;
    }
    mark(ws, STAGE_HEAD * 100 + L - 1);
    narrow_head_kernel<float, 4, 16, 8, true><<<grid, 256, smem, ws->stream>>>(a);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    mark(ws, STAGE_LOSS * 100);
    loss_final_kernel<<<1, 256, 0, ws->stream>>>(ws->loss_partial, grid, ws->error_id, (double)ws->norm_rows, (double)C, obj_out);
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
    int rc = fast_convert_w(ws, ws->w_trial);
    if (rc != OPL_OK) return rc;
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
