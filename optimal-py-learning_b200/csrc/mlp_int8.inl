// FP64 layer GEMMs on the INT8 tensor cores (Ozaki splitting, ozaki.cu) -- included by mlp.cu.
//
// Parity mode keeps FP64 inputs, FP64 outputs and FP64-level accuracy (3e-14 relative on the forward shape,
// 1.5e-12 on heavy-tailed weight gradients; gate 1e-10), but the two GEMMs that hold 98 % of an evaluation's
// flops -- forward layer l (A_l W_l) and the weight gradient (A_l^T delta_l) -- run as 21 exact INT8 products
// on tcgen05 instead of DMMA.  What makes it pay:
//   * the digits of X (forward operand) and of [1 ; X^T] (weight-gradient operand, the row of ones yields the
//     bias gradient in the flat layout's place) are extracted ONCE per set_data;
//   * per evaluation only W_l (tiny) and delta_l are cut into digits;
//   * bias, transfer function, derivative and dropout mask are fused into the INT8 GEMM's epilogue.
// Layers that are too small for the tensor-core tile (the narrow head) stay on the DMMA / streaming kernels.

struct Int8State {
    std::vector<char> fwd, dw, da;     // per layer: shape qualifies (da[l]: delta_{l-1} = delta_l W_l^T)
    int64_t rows_pad = 0;              // pad128(max_rows)
    // layer-0 operands, constant between set_data calls
    int8_t *x_dig = nullptr;           // [S][rows_pad][pad(d0)]
    double *x_scale = nullptr;         // per pattern
    int8_t *xt_dig = nullptr;          // [S][pad(d0+1)][rows_pad]; row 0 = ones
    double *xt_scale = nullptr;        // per feature (+ the ones row)
    const double *x_for = nullptr;     // the X the digits were cut from
    int64_t x_rows = 0;
    // per evaluation
    int8_t *w_dig = nullptr;           // [S][pad(dout)][pad(din)], largest qualifying layer
    double *w_scale = nullptr;
    int8_t *a_dig = nullptr;           // act[l] rows (l >= 1): [S][rows_pad][pad(din)]
    double *a_scale = nullptr;
    int8_t *at_dig = nullptr;          // act[l]^T (l >= 1): [S][pad(din)][rows_pad]
    double *at_scale = nullptr;
    int8_t *d_dig = nullptr;           // delta_l^T: [S][pad(dout)][rows_pad]
    double *d_scale = nullptr;
    double *colmax = nullptr;          // column-max scratch
    double *classmax = nullptr;        // 16 entries: max |delta_L[:, c]| of the fused head when it leaves delta_prev to the digit kernel
    double *dmax = nullptr;            // max |delta_l[:, j]| gathered by the kernel that writes delta_l (atomicMax), or unused
    int dmax_layer = -1;               // layer whose delta the maxima in dmax belong to (-1: none)
    bool store_z0 = false;             // this evaluation's layer-0 GEMM stores z (plain epilogue): the fused head applies the transfer
    bool accumulators_cleared = false; // classmax and d_scale were zeroed by the evaluation's first launch (no memsets needed)
    bool w0_digits_ready = false;      // the evaluation's first launch already cut the digits of W_0^T (ozaki_trial_w0_digits)
    int64_t colmax_len = 0;
};
// Padding: the digit kernels write zeros along the CONTRACTION index themselves (pad columns of a row-digit array,
// pad rows-of-src of a transposed one); padded rows of the M / N side only feed accumulator rows / columns that are
// never stored, so they may hold anything and no array is ever cleared.

namespace {

constexpr int64_t INT8_MIN_ROWS = 1024;

bool int8_shape_ok(int64_t din, int64_t dout, int l) {
    if (dout < 64 || din < 64) return false;
    return l == 0 || din * dout >= 128 * 128;
}

void int8_free(Int8State *s) {
    if (!s) return;
    cudaFree(s->x_dig);
    cudaFree(s->x_scale);
    cudaFree(s->xt_dig);
    cudaFree(s->xt_scale);
    cudaFree(s->w_dig);
    cudaFree(s->w_scale);
    cudaFree(s->a_dig);
    cudaFree(s->a_scale);
    cudaFree(s->at_dig);
    cudaFree(s->at_scale);
    cudaFree(s->d_dig);
    cudaFree(s->d_scale);
    cudaFree(s->colmax);
    cudaFree(s->dmax);
    cudaFree(s->classmax);
    delete s;
}

struct Int8Plan {
    int splits;
    int k_per_split;
};
// split the (padded) contraction so that tiles x splits fills the SMs and no split exceeds the INT32-exact length
Int8Plan int8_plan(int64_t m, int64_t n, int64_t k_pad) {
    const int64_t tiles = ceil_div(m, OZ_BM) * ceil_div(n, OZ_BN);
    int64_t splits = ceil_div(k_pad, OZ_MAX_K);
    splits = std::max<int64_t>(splits, std::min<int64_t>(sm_count() / std::max<int64_t>(tiles, 1), k_pad / 1024));
    splits = std::max<int64_t>(splits, 1);
    Int8Plan p;
    p.k_per_split = (int)(ceil_div(ceil_div(k_pad, splits), 128) * 128);
    p.splits = (int)ceil_div(k_pad, p.k_per_split);
    return p;
}

// decide per layer, allocate; returns the split-K scratch (in doubles) the weight gradients may need
int int8_create(opl_mlp *ws, int64_t *partials_needed) {
    const int L = ws->layers;
    Int8State *s = new Int8State();
    s->fwd.assign(L, 0);
    s->dw.assign(L, 0);
    s->da.assign(L, 0);
    s->rows_pad = ozaki_pad(ws->max_rows);
    bool any = false;
    int64_t w_bytes = 0, a_cols = 0, d_cols = 0, at_rows = 0, scale_len = 0;
    for (int l = 0; l < L; ++l) {
        const int64_t din = ws->width[l], dout = ws->width[l + 1];
        if (!int8_shape_ok(din, dout, l) || ws->max_rows < INT8_MIN_ROWS) continue;
        s->fwd[l] = l < L - 1 && ozaki_pad(din) <= OZ_MAX_K;     // one tile contracts at most OZ_MAX_K (INT32 exactness);
        s->dw[l] = 1;                                              // wider layers stay on the DMMA kernels
        any = true;
        w_bytes = std::max<int64_t>(w_bytes, (int64_t)OZ_S * ozaki_pad(dout) * ozaki_pad(din));
        scale_len = std::max<int64_t>(scale_len, std::max(ozaki_pad(dout), ozaki_pad(din + 1)));
        d_cols = std::max<int64_t>(d_cols, ozaki_pad(dout));
        if (l >= 1) {
            s->da[l] = ozaki_pad(dout) <= OZ_MAX_K;
            a_cols = std::max<int64_t>(a_cols, std::max(ozaki_pad(din), ozaki_pad(dout)));   // A_l rows (forward), delta_l rows (dA)
            at_rows = std::max<int64_t>(at_rows, ozaki_pad(din));
        }
        const Int8Plan plan = int8_plan(din + (l == 0 ? 1 : 0), dout, s->rows_pad);
        *partials_needed = std::max<int64_t>(*partials_needed, (int64_t)plan.splits * (din * dout + dout));
    }
    if (!any) {
        delete s;
        return OPL_OK;
    }
    ws->int8 = s;
    const int64_t rp = s->rows_pad;
    if (s->dw[0]) {
        const int64_t d0p = ozaki_pad(ws->width[0]), mp = ozaki_pad(ws->width[0] + 1);
        OPL_CUDA(cudaMalloc(&s->x_dig, (size_t)OZ_S * rp * d0p));
        OPL_CUDA(cudaMalloc(&s->x_scale, sizeof(double) * (size_t)rp));
        OPL_CUDA(cudaMalloc(&s->xt_dig, (size_t)OZ_S * mp * rp));
        OPL_CUDA(cudaMalloc(&s->xt_scale, sizeof(double) * (size_t)mp));
    }
    OPL_CUDA(cudaMalloc(&s->w_dig, (size_t)w_bytes));
    OPL_CUDA(cudaMalloc(&s->w_scale, sizeof(double) * (size_t)scale_len));
    OPL_CUDA(cudaMalloc(&s->d_dig, (size_t)OZ_S * d_cols * rp));
    OPL_CUDA(cudaMalloc(&s->d_scale, sizeof(double) * (size_t)d_cols));
    if (a_cols > 0) {
        OPL_CUDA(cudaMalloc(&s->a_dig, (size_t)OZ_S * rp * a_cols));
        OPL_CUDA(cudaMalloc(&s->a_scale, sizeof(double) * (size_t)rp));
        OPL_CUDA(cudaMalloc(&s->at_dig, (size_t)OZ_S * at_rows * rp));
        OPL_CUDA(cudaMalloc(&s->at_scale, sizeof(double) * (size_t)at_rows));
    }
    s->colmax_len = 4 * (int64_t)sm_count() * std::max<int64_t>(std::max(d_cols, at_rows), ozaki_pad(ws->width[0] + 1));
    OPL_CUDA(cudaMalloc(&s->colmax, sizeof(double) * (size_t)s->colmax_len));
    OPL_CUDA(cudaMalloc(&s->dmax, sizeof(double) * (size_t)d_cols));
    OPL_CUDA(cudaMalloc(&s->classmax, sizeof(double) * 16));
    return OPL_OK;
}

// digits of the resident X (forward operand) and of [1 ; X^T] (weight-gradient operand): once per set_data
int int8_set_x(opl_mlp *ws) {
    Int8State *s = ws->int8;
    if (!s || !s->dw[0]) return OPL_OK;
    s->x_for = nullptr;
    if (ws->rows < INT8_MIN_ROWS) return OPL_OK;     // small batches stay on the DMMA kernels
    const int64_t d0 = ws->width[0], rows = ws->rows;
    const int64_t mp = ozaki_pad(d0 + 1);
    mark(ws, STAGE_DIGITS * 100);
    int rc = ozaki_slice_rows(ws->x, rows, d0, d0, s->x_dig, s->x_scale, false, ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    rc = ozaki_slice_cols_transposed(ws->x, rows, d0, d0, s->xt_dig, s->xt_scale, s->colmax, s->colmax_len, 1, mp, nullptr,
                                     ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    rc = ozaki_ones_row(s->xt_dig, s->xt_scale, rows, mp, ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    s->x_for = ws->x;
    s->x_rows = rows;
    return OPL_OK;
}

bool int8_use(const opl_mlp *ws, int64_t rows) {
    return ws->int8 && ws->gemm_engine == OPL_ENGINE_INT8 && rows >= INT8_MIN_ROWS;
}
bool int8_fwd_ok(const opl_mlp *ws, int l, const double *in, int64_t rows) {
    if (!int8_use(ws, rows) || !ws->int8->fwd[l]) return false;
    return l > 0 || (in == ws->int8->x_for && rows == ws->int8->x_rows);
}
bool int8_dw_ok(const opl_mlp *ws, int l, const double *in, int64_t rows) {
    if (!int8_use(ws, rows) || !ws->int8->dw[l]) return false;
    return l > 0 || (in == ws->int8->x_for && rows == ws->int8->x_rows);
}

// The kernel about to write delta_l may gather max |delta_l[:, j]| on the fly: returns the (zeroed) buffer, or null when
// layer l's weight gradient does not run on the INT8 engine.
double *int8_arm_delta_max(opl_mlp *ws, int l, int64_t rows) {
    if (!int8_use(ws, rows) || !ws->int8->dw[l]) return nullptr;
    Int8State *s = ws->int8;
    if (l == 0 && (s->x_for == nullptr || rows != s->x_rows)) return nullptr;
    if (cudaMemsetAsync(s->dmax, 0, sizeof(double) * (size_t)ws->width[l + 1], ws->stream) != cudaSuccess) return nullptr;
    s->dmax_layer = l;
    return s->dmax;
}

// act[l+1] = f_l(in . W_l (+ b)), deriv[l] = f_l'  -- the forward GEMM of layer l
int int8_forward_layer(opl_mlp *ws, int l, const double *w, const double *in, int64_t rows, bool keep_deriv, bool masked) {
    Int8State *s = ws->int8;
    const int64_t din = ws->width[l], dout = ws->width[l + 1];
    const int64_t kp = ozaki_pad(din), np = ozaki_pad(dout), mp = ozaki_pad(rows);
    const bool last = l == ws->layers - 1;
    // B operand: W_l (din x dout row-major) -> digits of its transpose, one scale per output unit
    int rc = OPL_OK;
    if (l == 0 && s->w0_digits_ready) {
        s->w0_digits_ready = false;              // cut together with the trial point
    } else {
        mark(ws, STAGE_DIGITS * 100 + l);
        rc = ozaki_slice_cols_transposed(w + ws->w_off[l], din, dout, dout, s->w_dig, s->w_scale, s->colmax, s->colmax_len, 0,
                                         np, nullptr, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
    }
    const int8_t *a_dig = s->x_dig;
    const double *a_scale = s->x_scale;
    int64_t a_rows_pad = ozaki_pad(s->x_rows);
    if (l > 0) {
        rc = ozaki_slice_rows(in, rows, din, din, s->a_dig, s->a_scale, false, ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        a_dig = s->a_dig;
        a_scale = s->a_scale;
        a_rows_pad = mp;
    }
    OzakiArgs g = {};
    g.C = ws->act[l + 1];
    g.ldc = dout;
    g.M = (int)rows;
    g.N = (int)dout;
    g.K = (int)kp;
    g.k_valid = (int)din;
    g.k_per_split = (int)kp;
    g.splits = 1;
    g.scale_a = a_scale;
    g.scale_b = s->w_scale;
    g.bias = l == 0 ? w : nullptr;
    g.ldaux = dout;
    g.tparam = ws->tparam[l];
    switch (last ? OPL_TRANSFER_LINEAR : ws->transfer[l]) {
        case OPL_TRANSFER_TANH: g.epi = OZ_EPI_TANH; break;
        case OPL_TRANSFER_SOFTPLUS: g.epi = OZ_EPI_SOFTPLUS; break;
        case OPL_TRANSFER_GAUSSIAN: g.epi = OZ_EPI_GAUSSIAN; break;
        default: g.epi = OZ_EPI_STORE;
    }
    g.aux_out = (!last && keep_deriv) ? ws->deriv[l] : nullptr;
    g.colmask = (!last && ws->masks && masked) ? ws->masks + ws->mask_off[l + 1] : nullptr;
    if (l == 0 && s->store_z0) {       // late transfer: z = X W_0 + b as it stands
        g.epi = OZ_EPI_STORE;
        g.aux_out = nullptr;
    }
    if (kp > OZ_MAX_K) return fail(OPL_E_STATE, "int8 forward: layer %d is wider than %d inputs", l, OZ_MAX_K);
    mark(ws, STAGE_FWD * 100 + l);
    return launch_ozaki_gemm(a_dig, a_rows_pad, s->w_dig, np, kp, g, ws->stream, &ws->launches);
}

bool int8_da_ok(const opl_mlp *ws, int l, int64_t rows) { return l >= 1 && int8_use(ws, rows) && ws->int8->da[l]; }

// delta_{l-1} = (delta_l W_l^T) o f'_{l-1}: A = delta_l (row digits, one scale per pattern), B = W_l, which is already
// [N = din][K = dout] with K contiguous, so its row digits (one scale per hidden unit) are the operand as it stands
int int8_da_layer(opl_mlp *ws, int l, const double *w, int64_t rows) {
    Int8State *s = ws->int8;
    const int64_t din = ws->width[l], dout = ws->width[l + 1];
    const int64_t kp = ozaki_pad(dout), np = ozaki_pad(din), mp = ozaki_pad(rows);
    if (kp > OZ_MAX_K) return fail(OPL_E_STATE, "int8 back-propagation: layer %d is wider than %d outputs", l, OZ_MAX_K);
    mark(ws, STAGE_DIGITS * 100 + 70 + l);
    int rc = ozaki_slice_rows(w + ws->w_off[l], din, dout, dout, s->w_dig, s->w_scale, false, ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    rc = ozaki_slice_rows(ws->delta[l], rows, dout, dout, s->a_dig, s->a_scale, false, ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    OzakiArgs g = {};
    g.C = ws->delta[l - 1];
    g.ldc = din;
    g.M = (int)rows;
    g.N = (int)din;
    g.K = (int)kp;
    g.k_valid = (int)dout;
    g.k_per_split = (int)kp;
    g.splits = 1;
    g.scale_a = s->a_scale;
    g.scale_b = s->w_scale;
    g.ldaux = din;
    switch (ws->transfer[l - 1]) {
        case OPL_TRANSFER_SOFTPLUS:
        case OPL_TRANSFER_GAUSSIAN:
            g.epi = OZ_EPI_MUL_AUX;
            g.aux_in = ws->deriv[l - 1];
            break;
        case OPL_TRANSFER_TANH:
            g.epi = OZ_EPI_MUL_DTANH;
            g.aux_in = ws->act[l];
            break;
        default:
            g.epi = OZ_EPI_STORE;
    }
    mark(ws, STAGE_DA * 100 + l);
    return launch_ozaki_gemm(s->a_dig, mp, s->w_dig, np, kp, g, ws->stream, &ws->launches);
}

// delta_l is not materialised: its transposed digits are cut from (delta_{l+1}, W_{l+1}, slope) by ozaki_slice_delta_prev
struct DeltaOnTheFly {
    const double *deltaL;   // rows x C
    int C;
    const double *W;        // d_{l+1} x C
    const double *aux;      // rows x d_{l+1}
    int mode;
};

// dW_l = A_l^T delta_l (+ db = 1^T delta_0 for l = 0) into the flat gradient
int int8_dw_layer(opl_mlp *ws, int l, const double *in, int64_t rows, double *grad, const DeltaOnTheFly *otf = nullptr) {
    Int8State *s = ws->int8;
    const double *colmax_raw = s->dmax_layer == l ? s->dmax : nullptr;
    s->dmax_layer = -1;
    const int64_t din = ws->width[l], dout = ws->width[l + 1];
    const int64_t kp = ozaki_pad(rows), np = ozaki_pad(dout);
    const int64_t lead_rows = l == 0 ? 1 : 0, lead = lead_rows * dout;
    const int64_t m = din + lead_rows, mp = ozaki_pad(m);
    // B operand: delta_l (rows x dout) -> digits of its transpose, one scale per unit (max over the patterns)
    mark(ws, STAGE_DIGITS * 100 + 50 + l);
    int rc;
    if (otf)
        rc = ozaki_slice_delta_prev(otf->deltaL, otf->C, otf->W, otf->aux, otf->mode, rows, dout, dout, s->classmax, s->d_dig, s->d_scale,
                                    np, ws->stream, &ws->launches, s->accumulators_cleared);
    else
        rc = ozaki_slice_cols_transposed(ws->delta[l], rows, dout, dout, s->d_dig, s->d_scale, s->colmax, s->colmax_len, 0, np, colmax_raw,
                                         ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    const int8_t *a_dig = s->xt_dig;
    const double *a_scale = s->xt_scale;
    if (l > 0) {
        rc = ozaki_slice_cols_transposed(in, rows, din, din, s->at_dig, s->at_scale, s->colmax, s->colmax_len, 0, mp, nullptr,
                                         ws->stream, &ws->launches);
        if (rc != OPL_OK) return rc;
        a_dig = s->at_dig;
        a_scale = s->at_scale;
    }
    const Int8Plan plan = int8_plan(m, dout, kp);
    const int64_t count = lead + din * dout;
    if ((int64_t)plan.splits * count > ws->partials_len) return fail(OPL_E_STATE, "split-K scratch too small (int8 layer %d)", l);
    double *dst = grad + ws->w_off[l] - lead;          // bias sits right before W_0 in the flat layout
    OzakiArgs g = {};
    g.C = plan.splits > 1 ? ws->partials : dst;
    g.ldc = dout;
    g.M = (int)m;
    g.N = (int)dout;
    g.K = (int)kp;
    g.k_valid = (int)rows;
    g.k_per_split = plan.k_per_split;
    g.splits = plan.splits;
    g.split_stride = count;
    g.scale_a = a_scale;
    g.scale_b = s->d_scale;
    g.epi = OZ_EPI_STORE;
    g.ldaux = dout;
    if (ws->fold_pending && plan.splits > 1) {
        // the fused head's folds (side stream) read the head's per-CTA partials out of ws->partials, which this GEMM is about to
        // overwrite with its split-K partials: they rejoin the launch stream HERE (they have had the digit kernel's time to run)
        ws->fold_pending = false;
        OPL_CUDA(cudaStreamWaitEvent(ws->stream, ws->ev_fold, 0));
    }
    mark(ws, STAGE_DW * 100 + l);
    // digit arrays are pitched by the PADDED row count of the resident batch
    rc = launch_ozaki_gemm(a_dig, mp, s->d_dig, np, kp, g, ws->stream, &ws->launches);
    if (rc != OPL_OK) return rc;
    if (plan.splits > 1) {
        mark(ws, STAGE_DW_REDUCE * 100 + l);
        rc = launch_reduce_partials(ws->partials, plan.splits, count, count, dst, ws->stream, &ws->launches);
    }
    return rc;
}

}  // namespace
