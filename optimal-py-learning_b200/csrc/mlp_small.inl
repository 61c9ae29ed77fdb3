// One-launch evaluation for SMALL networks (BASELINE config 1, the README iris MLP) -- included by mlp.cu.
//
// The general schedule spends ~10 launches per evaluation; for a (4,2,3) network on 90 patterns every one of them is pure
// launch latency, and a Wolfe line search pays that 13 times per optimizer step.  Here the WHOLE probe is one kernel:
//
//   trial point  w = x + alpha*dir                                (linesearch.py:394)      every CTA, into shared memory
//   forward      A_{l+1} = f_l(A_l W_l [+ b])                     (mlp.py:148-163)         one thread per pattern
//   error        MSE / cross entropy with the reference's nan_to_num rules, delta_L      (error.py:46-116, mlp.py:250-253)
//   backward     delta_{l-1} = (delta_l W_l^T) o f'  |  softmax contraction               (mlp.py:254-260, 279-295)
//   gradient     dW_l = A_l^T delta_l, db = 1^T delta_0 in the flat layout                (mlp.py:269-276, 313-315)
//   reduce       the last CTA to finish folds the per-CTA partials in CTA order (bit-reproducible, no float atomics),
//                forms obj, g.dir, ||g||^2 and the domain flag, and writes them to the output vector, to device scalars
//                and to MAPPED PINNED host memory followed by a sequence number -- the host reads its three scalars
//                without a memcpy.
//
// A CTA owns `rpb` consecutive patterns; activations and deltas live in shared memory TRANSPOSED ([unit][pattern], odd
// pitch) so that the thread-per-pattern phases and the entry-per-thread gradient phase are both conflict-free; weights are
// read from shared memory as broadcasts.  FP64 throughout, library exp / log / tanh: same formulas as the general path.

namespace {

constexpr int SMALL_MAX_LAYERS = 8;
constexpr int SMALL_THREADS = 128;
constexpr int SMALL_MAX_PARAMS = 4096;
constexpr int SMALL_MAX_WIDTH = 128;
constexpr int64_t SMALL_MAX_PARTIAL = 1 << 20;      // grid x n doubles folded by one CTA

struct SmallArgs {
    const double *x, *dir;       // parameters, direction (may be null)
    double alpha;
    const double *X, *Y;         // rows x d0, rows x dL
    int64_t rows;
    int L, n, rpb, pitch;        // pitch = rpb + 1 (odd)
    int width[SMALL_MAX_LAYERS + 1];
    int transfer[SMALL_MAX_LAYERS];
    double tparam[SMALL_MAX_LAYERS];
    int w_off[SMALL_MAX_LAYERS];
    int act_off[SMALL_MAX_LAYERS + 1];   // shared-memory offsets (doubles) of A_l, l = 0..L
    int del_off[SMALL_MAX_LAYERS];       // of delta_l (layer l's output side), l = 0..L-1
    int eid;
    double ce_div, mse_mul, norm_rows;
    int want_grad;
    double *partial;             // [grid][n + 2]: gradient partials, loss partial, bad-value count
    double *out;                 // n + 2: [gradient ; objective ; domain flag]
    double *scal;                // device: obj, g.dir, g.g, flag
    volatile double *host_scal;  // mapped pinned: obj, g.dir, g.g, flag, sequence number
    double seq;
    unsigned int *ticket;
};

__global__ void __launch_bounds__(SMALL_THREADS) smallnet_eval_kernel(const SmallArgs p) {
    extern __shared__ double ssm[];
    __shared__ double scratch[33];
    __shared__ int s_last;
    const int tid = threadIdx.x, T = SMALL_THREADS, P = p.pitch, L = p.L, n = p.n;
    double *w = ssm;
    const int64_t r0 = (int64_t)blockIdx.x * p.rpb;
    const int nr = (int)min((int64_t)p.rpb, p.rows - r0);

    // ---- trial point (two roundings, like NumPy's x + alpha * d) and the CTA's patterns ----
    for (int i = tid; i < n; i += T) w[i] = p.dir ? __dadd_rn(p.x[i], __dmul_rn(p.alpha, p.dir[i])) : p.x[i];
    {
        const int d0 = p.width[0];
        double *a0 = ssm + p.act_off[0];
        const double *src = p.X + r0 * d0;
        for (int k = tid; k < nr * d0; k += T) {
            const int r = k / d0, i = k - r * d0;
            a0[i * P + r] = src[k];
        }
    }
    __syncthreads();

    double loss = 0.0;
    bool bad = false;
    const int r = tid;
    if (r < nr) {
        // ---- forward ----
        for (int l = 0; l < L; ++l) {
            const int din = p.width[l], dout = p.width[l + 1], t = p.transfer[l];
            const double *W = w + p.w_off[l];
            const double *ain = ssm + p.act_off[l] + r;
            double *aout = ssm + p.act_off[l + 1] + r;
            double *dl = ssm + p.del_off[l] + r;
            const bool last = l == L - 1;
            double mx = -INFINITY;
            for (int j = 0; j < dout; ++j) {
                double z = l == 0 ? w[j] : 0.0;              // only the first layer has a bias (mlp.py:150-151)
                for (int i = 0; i < din; ++i) z = fma(ain[i * P], W[i * dout + j], z);
                aout[j * P] = z;
                mx = fmax(mx, z);
            }
            if (t == OPL_TRANSFER_SOFTMAX) {                 // calculate.py:152-153
                // a thread is ONE dependent FP64 chain here, so exp / reciprocal are the short table-driven forms of
                // oz_transcend.cuh (~11 and 3 dependent operations instead of ~40 and ~12; ~1 ulp each), as in the fused head
                double den = 0.0;
                for (int j = 0; j < dout; ++j) {
                    const double e = ozt_exp_nonpos(aout[j * P] - mx);
                    aout[j * P] = e;
                    den += e;
                }
                const bool regular_den = den >= 1.0 && den <= 1e300;           // (NaN / inf logits: literal division)
                const double rden = regular_den ? ozt_rcp(den) : 0.0;
                for (int j = 0; j < dout; ++j) aout[j * P] = regular_den ? aout[j * P] * rden : aout[j * P] / den;
            } else if (t == OPL_TRANSFER_SOFTPLUS) {         // calculate.py:128-141: value and slope share exp(-|z|) and 1/(1+t)
                for (int j = 0; j < dout; ++j) {
                    const double zin[1] = {aout[j * P]};
                    double sp[1], sg[1];
                    ozt_softplus_logistic_n<1>(zin, sp, sg);
                    aout[j * P] = sp[0];
                    if (p.want_grad || last) dl[j * P] = sg[0];
                }
            } else {
                for (int j = 0; j < dout; ++j) {
                    const double z = aout[j * P];
                    const double a = transfer_value(t, p.tparam[l], z);
                    aout[j * P] = a;
                    if (p.want_grad || last) dl[j * P] = transfer_slope(t, p.tparam[l], z, a);
                }
            }
        }
        // ---- error and delta_L ----
        {
            const int C = p.width[L], t = p.transfer[L - 1];
            const double *a = ssm + p.act_off[L] + r;
            double *dl = ssm + p.del_off[L - 1] + r;
            const double *y = p.Y + (r0 + r) * C;
            RowArgs ra = {};
            ra.eid = p.eid;
            ra.ce_div = p.ce_div;
            ra.mse_mul = p.mse_mul;
            if (t == OPL_TRANSFER_SOFTMAX) {                 // dsoftmax contracted with g (mlp.py:295): softmax_delta()
                double tt = 0.0;
                const bool ce = p.eid == OPL_ERROR_CROSS_ENTROPY;
                const double ce_rdiv = 1.0 / p.ce_div;
                for (int j = 0; j < C; ++j) {
                    const double aj = a[j * P], yj = y[j];
                    double g;
                    if (ce && aj >= 9.332636185032189e-302) {      // regular probability (>= 2^-1000): short log / reciprocal,
                        loss += ozt_log_pos(aj) * yj;              // same nan_to_num-free arithmetic as error.py:89-116 there
                        g = (yj * ozt_rcp(aj)) * ce_rdiv;
                    } else {
                        g = error_term(ra, aj, yj, loss, bad);     // 0, subnormal, negative, NaN: the literal rules
                    }
                    dl[j * P] = g;
                    tt += g * a[j * P];
                }
                for (int j = 0; j < C; ++j) dl[j * P] = softmax_delta(a[j * P], dl[j * P], tt);
            } else {
                for (int j = 0; j < C; ++j) dl[j * P] = error_term(ra, a[j * P], y[j], loss, bad) * dl[j * P];
            }
        }
        // ---- backward ----
        if (p.want_grad) {
            for (int l = L - 1; l >= 1; --l) {
                const int din = p.width[l], dout = p.width[l + 1], t = p.transfer[l - 1];
                const double *W = w + p.w_off[l];
                const double *dn = ssm + p.del_off[l] + r;
                double *dp = ssm + p.del_off[l - 1] + r;
                const double *a = ssm + p.act_off[l] + r;
                if (t == OPL_TRANSFER_SOFTMAX) {
                    double tt = 0.0;
                    for (int i = 0; i < din; ++i) {
                        double s = 0.0;
                        for (int j = 0; j < dout; ++j) s = fma(dn[j * P], W[i * dout + j], s);
                        dp[i * P] = s;
                        tt += s * a[i * P];
                    }
                    for (int i = 0; i < din; ++i) dp[i * P] = softmax_delta(a[i * P], dp[i * P], tt);
                } else {
                    for (int i = 0; i < din; ++i) {
                        double s = 0.0;
                        for (int j = 0; j < dout; ++j) s = fma(dn[j * P], W[i * dout + j], s);
                        dp[i * P] = s * dp[i * P];           // slope stored by the forward pass
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- this CTA's share of the flat gradient: one entry per thread and step, patterns summed in order ----
    double *part = p.partial + (size_t)blockIdx.x * (n + 2);
    if (p.want_grad) {
        for (int e = tid; e < n; e += T) {
            double s = 0.0;
            if (e < p.width[1]) {                             // bias gradient: column sums of delta_0 (mlp.py:276)
                const double *d0 = ssm + p.del_off[0] + e * P;
                for (int q = 0; q < nr; ++q) s += d0[q];
            } else {
                int l = 0;
                while (l + 1 < L && e >= p.w_off[l + 1]) ++l;
                const int dout = p.width[l + 1], k = e - p.w_off[l];
                const int i = k / dout, j = k - i * dout;
                const double *a = ssm + p.act_off[l] + i * P, *d = ssm + p.del_off[l] + j * P;
                for (int q = 0; q < nr; ++q) s = fma(a[q], d[q], s);
            }
            part[e] = s;
        }
    }
    loss = block_sum(loss, scratch);
    const double nbad = block_sum(bad ? 1.0 : 0.0, scratch);
    if (tid == 0) {
        part[n] = loss;
        part[n + 1] = nbad;
    }

    // ---- the last CTA folds the partials in CTA order ----
    const int G = gridDim.x;
    if (G > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(p.ticket, 1u) == (unsigned)(G - 1);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    } else {
        __syncthreads();
    }
    double gd = 0.0, gg = 0.0;
    if (p.want_grad) {
        for (int e = tid; e < n; e += T) {
            double g = 0.0;
            for (int b = 0; b < G; ++b) g += __ldcg(p.partial + (size_t)b * (n + 2) + e);
            p.out[e] = g;
            if (p.dir) gd = fma(g, p.dir[e], gd);
            gg = fma(g, g, gg);
        }
    }
    gd = block_sum(gd, scratch);
    gg = block_sum(gg, scratch);
    if (tid == 0) {
        double total = 0.0, flagged = 0.0;
        for (int b = 0; b < G; ++b) {
            total += __ldcg(p.partial + (size_t)b * (n + 2) + n);
            flagged += __ldcg(p.partial + (size_t)b * (n + 2) + n + 1);
        }
        const double obj = p.eid == OPL_ERROR_MSE ? total / (p.norm_rows * (double)p.width[L]) : -(total / p.norm_rows);
        const double flag = flagged != 0.0 ? 1.0 : 0.0;
        p.out[n] = obj;
        p.out[n + 1] = flag;
        p.scal[0] = obj;
        p.scal[1] = gd;
        p.scal[2] = gg;
        p.scal[3] = flag;
        if (p.host_scal) {
            p.host_scal[0] = obj;
            p.host_scal[1] = gd;
            p.host_scal[2] = gg;
            p.host_scal[3] = flag;
            __threadfence_system();
            p.host_scal[4] = p.seq;
        }
        if (G > 1) *p.ticket = 0;
    }
}

struct SmallPlan {
    bool ok = false;
    int rpb = 0, pitch = 0, grid = 0;
    size_t smem = 0;
    int act_off[SMALL_MAX_LAYERS + 1];
    int del_off[SMALL_MAX_LAYERS];
};

// Is the resident problem small enough for the one-launch kernel?  (parity mode, no dropout masks)
SmallPlan small_plan(const opl_mlp *ws, int64_t rows) {
    SmallPlan pl;
    if (ws->precision != OPL_PRECISION_F64 || ws->masks || !ws->small_path) return pl;
    if (ws->layers > SMALL_MAX_LAYERS || ws->n > SMALL_MAX_PARAMS) return pl;
    int64_t units = 0;
    for (int l = 0; l <= ws->layers; ++l) {
        if (ws->width[l] > SMALL_MAX_WIDTH) return pl;
        units += ws->width[l] * (l == 0 ? 1 : 2);
    }
    const int64_t npad = (ws->n + 1) & ~(int64_t)1;
    for (int rpb : {128, 64, 32}) {
        if (rpb > 32 && rows <= rpb / 2) continue;                 // few patterns: no point in a mostly empty block
        const size_t smem = sizeof(double) * (size_t)(npad + units * (rpb + 1));
        if (smem > 200 * 1024) continue;
        const int64_t grid = ceil_div(rows, rpb);
        if (grid * (ws->n + 2) > SMALL_MAX_PARTIAL || grid > 4096) return pl;
        pl.ok = true;
        pl.rpb = rpb;
        pl.pitch = rpb + 1;
        pl.grid = (int)grid;
        pl.smem = smem;
        int at = (int)npad;
        for (int l = 0; l <= ws->layers; ++l) {
            pl.act_off[l] = at;
            at += (int)ws->width[l] * pl.pitch;
        }
        for (int l = 0; l < ws->layers; ++l) {
            pl.del_off[l] = at;
            at += (int)ws->width[l + 1] * pl.pitch;
        }
        return pl;
    }
    return pl;
}

// One launch: evaluation at x + alpha*dir, [gradient ; objective ; flag] into grad_out_dev, scalars into ws->scal and the
// mapped host slot.  Returns the sequence number the kernel will publish (0: no host slot).
int small_evaluate(opl_mlp *ws, const SmallPlan &pl, const double *x_dev, double alpha, const double *dir_dev, bool want_grad,
                   double *grad_out_dev, double *seq_out) {
    if ((int64_t)pl.grid * (ws->n + 2) > ws->partials_len) return fail(OPL_E_STATE, "small-network path: scratch too small");
    SmallArgs a = {};
    a.x = x_dev;
    a.dir = dir_dev;
    a.alpha = alpha;
    a.X = ws->x;
    a.Y = ws->y;
    a.rows = ws->rows;
    a.L = ws->layers;
    a.n = (int)ws->n;
    a.rpb = pl.rpb;
    a.pitch = pl.pitch;
    for (int l = 0; l <= ws->layers; ++l) {
        a.width[l] = (int)ws->width[l];
        a.act_off[l] = pl.act_off[l];
    }
    for (int l = 0; l < ws->layers; ++l) {
        a.transfer[l] = ws->transfer[l];
        a.tparam[l] = ws->tparam[l];
        a.w_off[l] = (int)ws->w_off[l];
        a.del_off[l] = pl.del_off[l];
    }
    a.eid = ws->error_id;
    a.norm_rows = (double)ws->norm_rows;
    a.ce_div = -(double)ws->norm_rows;
    a.mse_mul = 2.0 / ((double)ws->norm_rows * (double)ws->width[ws->layers]);
    a.want_grad = want_grad ? 1 : 0;
    a.partial = ws->partials;
    a.out = grad_out_dev;
    a.scal = ws->scal;
    a.host_scal = ws->mapped_dev;
    ws->mapped_seq += 1.0;
    a.seq = ws->mapped_seq;
    a.ticket = ws->ticket;
    bool &configured = ws->small_configured;
    if (!configured || pl.smem > ws->small_smem) {
        OPL_CUDA(cudaFuncSetAttribute(smallnet_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        configured = true;
        ws->small_smem = 200 * 1024;
    }
    mark(ws, STAGE_SMALL * 100);
    smallnet_eval_kernel<<<pl.grid, SMALL_THREADS, pl.smem, ws->stream>>>(a);
    OPL_LAUNCH_CHECK();
    ++ws->launches;
    mark(ws, STAGE_END * 100);
    if (seq_out) *seq_out = a.seq;
    return OPL_OK;
}

// Wait for the kernel that publishes `seq` and read its four scalars from the mapped slot: a short spin on host memory
// (the kernel's last action is the sequence store), falling back to a stream synchronisation.
int small_read_scalars(opl_mlp *ws, double seq, double *host_out) {
    volatile double *m = ws->mapped_host;
    bool seen = false;
    if (m) {
        for (int spin = 0; spin < 2000000; ++spin) {
            if (m[4] == seq) {
                seen = true;
                break;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    if (!seen) {
        OPL_CUDA(cudaStreamSynchronize(ws->stream));
        if (!m || m[4] != seq) {      // no mapped slot (or an error ended the kernel): take the device copy
            OPL_CUDA(cudaMemcpy(ws->pinned, ws->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost));
            m = ws->pinned;
        }
    }
    host_out[0] = m[0];
    host_out[1] = m[1];
    host_out[2] = m[2];
    if (m[3] != 0.0) return fail(OPL_E_DOMAIN, "invalid value encountered in log (negative prediction in cross entropy)");
    return OPL_OK;
}

}  // namespace
