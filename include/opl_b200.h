/*
 * opl_b200.h -- C ABI of the B200-native hot path of optimal-py-learning.
 *
 * Drop-in boundary.  Every entry point replaces a NumPy call
 * site of the reference (`learning` 0.1.0, paths relative to /root/reference/learning);
 * the reference-side binding (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes only; no torch / C++ types cross this boundary.
 *   - `_host` entry points take HOST buffers (what the reference holds: NumPy float64
 *     arrays) and do the host<->device copies themselves; `_device` entry points take
 *     DEVICE pointers (float64) valid on the current CUDA device.
 *   - all work is enqueued on the stream given at create time (NULL = default stream);
 *     `_host` calls and calls with a `double *host_out` argument synchronise that stream
 *     before returning, all other calls are asynchronous.
 *   - every function returns OPL_OK (0) or a negative OPL_E_* code; opl_last_error()
 *     returns a thread-local message.  There is NO CPU fallback: without a CUDA device
 *     every compute entry point returns OPL_E_CUDA.
 *   - flat parameter / gradient layout is the reference's (architecture/mlp.py:313-327):
 *       [ bias(d1) ; W0 (d0 x d1 row-major) ; W1 (d1 x d2) ; ... ]   length n.
 */
#ifndef OPL_B200_H
#define OPL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPL_OK           0
#define OPL_E_ARG       -1   /* bad argument (shape mismatch, unknown id, null pointer) */
#define OPL_E_CUDA      -2   /* CUDA runtime error / no device                          */
#define OPL_E_STATE     -3   /* call sequence error (no data set, ...)                  */
#define OPL_E_DOMAIN    -4   /* log(negative) in cross entropy: the reference raises
                                FloatingPointError (error.py:89-91)                    */

/* transfer ids: transfer.py:47-130 / calculate.py:91-176 */
#define OPL_TRANSFER_LINEAR   0
#define OPL_TRANSFER_TANH     1
#define OPL_TRANSFER_SOFTPLUS 2   /* the reference's "ReluTransfer" is softplus */
#define OPL_TRANSFER_GAUSSIAN 3   /* param = variance */
#define OPL_TRANSFER_SOFTMAX  4
#define OPL_TRANSFER_LOGIT    5   /* output layers only: calculate.logit / dlogit (calculate.py:65-88),
                                     the equation of LogisticRegressionModel (regression.py:255-290) */

/* error ids: error.py:46-116 */
#define OPL_ERROR_MSE            0
#define OPL_ERROR_CROSS_ENTROPY  1

/* arithmetic modes */
#define OPL_PRECISION_F64   0   /* parity mode: FP64 in / out, matches NumPy float64 to 1e-10 */
#define OPL_PRECISION_TF32  1   /* fast mode: tcgen05 TF32, FP32 accumulate (1e-3)         */

/* which kernels run the large layer GEMMs of the parity mode (both are FP64-accurate; small layers always use DMMA) */
#define OPL_ENGINE_DMMA     0   /* mma.sync.m8n8k4.f64                                               */
#define OPL_ENGINE_INT8     1   /* Ozaki splitting: 21 exact INT8 products on tcgen05.mma.kind::i8, FP64 recombination
                                   (default; environment OPL_F64_ENGINE=dmma selects the other one)  */

const char *opl_last_error(void);
int opl_version(void);
/* number of CUDA devices visible (0 on a CPU-only host; never an error) */
int opl_device_count(void);

/* =====================================================================================
 * MLP objective / gradient workspace
 *   replaces MLP.activate / _get_obj / _get_obj_jac / _get_jacobians
 *   (architecture/mlp.py:134-163, 195-276)
 * ===================================================================================== */
typedef struct opl_mlp opl_mlp;

/* widths[0..n_widths-1] = the reference's `shape` tuple; transfer_ids / transfer_params have
 * n_widths-1 entries.  max_rows = capacity (patterns resident on this device). */
int opl_mlp_create(opl_mlp **out, int32_t n_widths, const int64_t *widths,
                   const int32_t *transfer_ids, const double *transfer_params,
                   int32_t error_id, int64_t max_rows, int32_t precision, void *cuda_stream);
int opl_mlp_destroy(opl_mlp *ws);
int64_t opl_mlp_param_count(const opl_mlp *ws);
/* select OPL_ENGINE_* for the parity mode's large GEMMs (takes effect at the next evaluation) */
int opl_mlp_set_engine(opl_mlp *ws, int32_t engine);
/* Re-bind the workspace to another CUDA stream (the caller entered a different stream context).  Work already
 * enqueued on the old stream is waited for first, so the new stream never races with it. */
int opl_mlp_set_stream(opl_mlp *ws, void *cuda_stream);

/* Make `rows` patterns resident (copied into the workspace).  `norm_rows` is the N of the
 * reference's 1/N (cross entropy, error.py:115-116) and 1/(N*C) (MSE, error.py:62) scales:
 * rows for a single device, the GLOBAL row count when the batch is sharded over ranks. */
int opl_mlp_set_data_host(opl_mlp *ws, const double *x, const double *y, int64_t rows, int64_t norm_rows);
int opl_mlp_set_data_device(opl_mlp *ws, const double *x_dev, const double *y_dev, int64_t rows, int64_t norm_rows);

/* ---- host-buffer calls: exactly what MLP._get_obj / _get_obj_jac / activate receive ---- */
/* _get_obj (mlp.py:195-199) */
int opl_mlp_obj_host(opl_mlp *ws, const double *params, double *obj_out);
/* _get_obj_jac (mlp.py:202-208): grad_out has n entries, flat layout */
int opl_mlp_obj_jac_host(opl_mlp *ws, const double *params, double *obj_out, double *grad_out);
/* activate (mlp.py:134-163) on `rows` fresh patterns (rows <= max_rows); out is rows x d_L.
 * Does not disturb the resident training data. */
int opl_mlp_activate_host(opl_mlp *ws, const double *params, const double *x, int64_t rows, double *out);

/* ---- device-pointer calls used by the device-resident optimizer stack ---- */
/* Evaluate at  x + alpha*dir  (dir_dev may be NULL => at x): the line-search probe
 * `obj_jac_func(parameters + step_size * step_dir)` (linesearch.py:392-399).
 * grad_out_dev has n+2 entries: [ flat gradient ; objective ; domain flag ].  The flag is 1.0 when
 * the cross entropy met log(negative) (error.py:89-91), else 0.0; it travels in the vector so that a
 * sharded evaluation all-reduces it and every rank raises at the same evaluation.  With
 * want_grad == 0 only entries n and n+1 are written (the _get_obj path used by backtracking,
 * linesearch.py:217).  Partial sums only: when the batch is sharded the caller all-reduces the
 * n+2 doubles.  Asynchronous. */
int opl_mlp_eval_device(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev,
                        int32_t want_grad, double *grad_out_dev);
/* Read back the probe scalars after (optionally all-reducing) grad_out_dev (n+2 entries):
 *   host_out[0] = objective, [1] = grad . dir (0 if dir NULL), [2] = ||grad||^2.
 * Synchronises; returns OPL_E_DOMAIN if the flag slot grad_dev[n+1] is non-zero. */
int opl_mlp_probe_scalars(opl_mlp *ws, const double *grad_dev, const double *dir_dev,
                          int32_t have_grad, double *host_out);
/* opl_mlp_eval_device + opl_mlp_probe_scalars in ONE call for an unsharded batch (the line-search probe of
 * linesearch.py:392-399 on a single device): evaluates at x + alpha*dir, writes [gradient ; objective ; flag] and returns
 * host_out[0..2] = objective, grad . dir, ||grad||^2.  Small networks (<= 4096 weights, layers <= 128 wide) run as ONE
 * kernel that also forms the three scalars and hands them over through mapped pinned memory (csrc/mlp_small.inl).
 * Synchronises; OPL_E_DOMAIN on log(negative). */
int opl_mlp_probe_device(opl_mlp *ws, const double *x_dev, double alpha, const double *dir_dev, int32_t want_grad,
                         double *grad_out_dev, double *host_out);
/* enable / disable the one-launch small-network kernel (default on; environment OPL_SMALL_PATH=0 turns it off) */
int opl_mlp_set_small_path(opl_mlp *ws, int32_t enable);
/* activate on device buffers: x_dev rows x d0 -> out_dev rows x d_L */
int opl_mlp_activate_device(opl_mlp *ws, const double *params_dev, const double *x_dev, int64_t rows,
                            double *out_dev);
/* DropoutMLP (mlp.py:330-445): 0/1 vectors of active neurons for one train_step, concatenated as
 * [input mask (d0) ; hidden mask 1 (d1) ; ... ; hidden mask L-1 (d_{L-1})] (no mask on the output
 * layer).  NULL clears them.  The input mask is folded into W0's rows (and dW0's), the hidden masks
 * multiply the transfer outputs in the forward epilogue; derivatives are NOT masked (mlp.py:438-445).
 * Parity mode only. */
int opl_mlp_set_masks_host(opl_mlp *ws, const double *masks, int64_t len);
/* number of kernels this workspace has launched since creation (bench.py gpu_launches) */
int64_t opl_mlp_launch_count(const opl_mlp *ws);
/* Per-launch device timing for the roofline report: while enabled, a CUDA event is recorded on
 * the workspace stream before every kernel of an evaluation.  profile_read synchronises and
 * returns, per launch, kind = stage*100 + layer (stage 0 trial point, 1 forward GEMM, 2 row
 * kernel, 3 dW GEMM, 4 dW split-K reduce, 5 dA GEMM, 6 bias column sum, 7 loss) and its
 * duration in ms, then clears the record. */
int opl_mlp_profile(opl_mlp *ws, int32_t enable);
int opl_mlp_profile_read(opl_mlp *ws, int32_t capacity, int32_t *kinds, float *ms, int32_t *count);

/* Benchmark hook (not on the training path): average ms of `iters` launches of one plain FP64 GEMM
 * of the family the MLP uses.  layout 0: C[M,N]=A[M,K].B[K,N]; 1: A[M,K].B[N,K]^T; 2: A[K,M]^T.B[K,N]
 * (split-K through scratch_dev when given).  All buffers are device pointers. */
int opl_bench_gemm_f64(int32_t layout, int32_t M, int32_t N, int32_t K, const double *a_dev, const double *b_dev,
                       double *c_dev, double *scratch_dev, int64_t scratch_len, int32_t iters, float *ms_out);

/* Same hook for the fast-mode TF32 family (tcgen05.mma + TMEM + TMA): FP32 row-major device buffers,
 * N % 32 == 0, K % 32 == 0 (M % 4 == 0 for layout 2); `splits` > 1 runs split-K through scratch_dev. */
int opl_bench_gemm_tf32(int32_t layout, int32_t M, int32_t N, int32_t K, const float *a_dev, const float *b_dev,
                        float *c_dev, float *scratch_dev, int64_t scratch_len, int32_t splits, int32_t iters,
                        float *ms_out);

/* Validation hook for the FP64-on-INT8 (Ozaki) GEMM: C[M,N] = epi(A[M,K].B[K,N]), row-major FP64 device buffers.
 * epi: 0 store, 1 tanh, 2 softplus, 3 gaussian (as the layer GEMMs fuse them).
 * slice_ms_out = time of the two digit-extraction passes, gemm_ms_out = tcgen05 INT8 GEMM (+ FP64 fold of K splits). */
int opl_bench_gemm_ozaki(int32_t M, int32_t N, int32_t K, const double *a_dev, const double *b_dev, double *c_dev,
                         int32_t epi, int32_t iters, float *slice_ms_out, float *gemm_ms_out);

/* ---- stand-alone host-buffer forms of the plugin objects' own methods (not on the training
 *      path; they run the same row kernel so the classes stay usable on their own) ---- */
/* Transfer.__call__ (transfer.py:33-130) on a rows x cols matrix */
int opl_transfer_apply_host(int32_t transfer_id, double param, const double *x, int64_t rows, int64_t cols,
                            double *out);
/* upstream . J_f(x): the product MLP backprop forms (mlp.py:279-295); with upstream = ones and an
 * element-wise transfer this is Transfer.derivative */
int opl_transfer_backprop_host(int32_t transfer_id, double param, const double *x, const double *upstream,
                               int64_t rows, int64_t cols, double *out);
/* ErrorFunc.__call__ / derivative (error.py:46-116); deriv_out (rows x cols) may be NULL */
int opl_error_host(int32_t error_id, const double *a, const double *b, int64_t rows, int64_t cols,
                   double *err_out, double *deriv_out);
/* calculate.dsoftmax (calculate.py:156-176): rows x cols x cols Jacobians from the softmax OUTPUT */
int opl_softmax_jacobian_host(const double *y, int64_t rows, int64_t cols, double *jac_out);

/* =====================================================================================
 * BFGS inverse-Hessian state   (optimize/optimizer.py:234-338)
 *   H is n x n float64 row-major, rows [row_begin,row_end) resident on this device.
 *   The rank-2 update  H+ = (I - rho s y^T) H (I - rho y s^T) + rho s s^T  is applied IN
 *   PLACE and lazily: one fused pass per iteration reads H once, applies the pending
 *   update, writes H+ and accumulates H+.[y g] and y^T.H+ for the next direction
 *   (2 n^2 8 bytes per iteration instead of the reference's 4 n^3 flop).
 * ===================================================================================== */
typedef struct opl_bfgs opl_bfgs;

#define OPL_BFGS_SEED_IDENTITY         0   /* initial_hessian_identity        optimizer.py:148 */
#define OPL_BFGS_SEED_SCALED_IDENTITY  1   /* initial_hessian_scaled_identity optimizer.py:154 */

int opl_bfgs_create(opl_bfgs **out, int64_t n, int64_t row_begin, int64_t row_end, void *cuda_stream);
int opl_bfgs_destroy(opl_bfgs *b);
/* BFGS.reset (optimizer.py:223-232): forget H */
int opl_bfgs_reset(opl_bfgs *b);
/* re-bind to another stream (see opl_mlp_set_stream) */
int opl_bfgs_set_stream(opl_bfgs *b, void *cuda_stream);

/* One call of BFGS._get_approx_inv_hessian + `-(H.dot(jac))` (optimizer.py:242-285), single
 * device.  prev_step_dev == NULL => first iteration (dir = -grad, identity never formed).
 * Otherwise s = prev_step, y = grad - prev_grad.  Asynchronous. */
int opl_bfgs_direction_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                              const double *prev_grad_dev, int32_t seed_mode, double *dir_out_dev);

/* g . dir of the direction the last opl_bfgs_direction_device call produced, when that call ran as the single-launch
 * small-problem kernel (n <= 512), which forms it on the fly: the line search's phi'(0) (linesearch.py:259) without a separate
 * dot product.  Waits for the kernel.  OPL_E_STATE when the streaming path produced the direction. */
int opl_bfgs_last_slope(opl_bfgs *b, double *host_out);

/* Row-sharded form (one rank per GPU): pass -> (caller: all-gather u,w slices; all-reduce v)
 * -> finish.  u_dev,v_dev,w_dev are full n-vectors owned by the caller; pass writes
 * u[row_begin:row_end] = (H+ y) rows, w[...] = (H+ g) rows and v = the LOCAL rows'
 * contribution to y^T H+ (all n columns).  On the first two iterations (no H yet) pass is a
 * no-op and finish works from vectors alone. */
int opl_bfgs_pass_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                         const double *prev_grad_dev, int32_t seed_mode,
                         double *u_dev, double *v_dev, double *w_dev);
int opl_bfgs_finish_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                           const double *prev_grad_dev, int32_t seed_mode,
                           const double *u_dev, const double *v_dev, const double *w_dev,
                           double *dir_out_dev);

/* Apply any pending update and copy the local rows of H (the reference's
 * `_prev_inv_hessian`) to h_out_dev ((row_end-row_begin) x n).  OPL_E_STATE if no H exists yet. */
int opl_bfgs_materialize_device(opl_bfgs *b, double *h_out_dev);
/* Test / benchmark hooks: load an explicit H (device, local rows) as the current state;
 * run only the fused update+products pass and report nothing (timed by the caller). */
int opl_bfgs_load_device(opl_bfgs *b, const double *h_dev);
int opl_bfgs_fill_synthetic(opl_bfgs *b, uint64_t seed);
int64_t opl_bfgs_launch_count(const opl_bfgs *b);
/* 0 = no H, 1 = seed pending (H not materialised), 2 = H resident */
int opl_bfgs_state(const opl_bfgs *b);

/* =====================================================================================
 * L-BFGS two-loop recursion   (optimize/optimizer.py:414-508)
 * ===================================================================================== */
typedef struct opl_lbfgs opl_lbfgs;

#define OPL_LBFGS_GAMMA_ONE     0   /* initial_hessian_one_scalar   optimizer.py:341 */
#define OPL_LBFGS_GAMMA_OLDEST  1   /* initial_hessian_gamma_scalar optimizer.py:346 (oldest pair) */

/* memory = num_remembered_iterations (<= 0: unbounded, grows on demand) */
int opl_lbfgs_create(opl_lbfgs **out, int64_t n, int32_t memory, void *cuda_stream);
int opl_lbfgs_destroy(opl_lbfgs *l);
int opl_lbfgs_reset(opl_lbfgs *l);
/* LBFGS._update_diffs + _lbfgs_step_dir: pop the oldest pair if the history is full, push
 * (s = prev_step, y = grad - prev_grad) when prev_step_dev != NULL, then run the two-loop
 * recursion on `grad`.  Asynchronous. */
int opl_lbfgs_direction_device(opl_lbfgs *l, const double *grad_dev, const double *prev_step_dev,
                               const double *prev_grad_dev, int32_t gamma_mode, double *dir_out_dev);
int32_t opl_lbfgs_history_len(const opl_lbfgs *l);
/* re-bind to another stream (see opl_mlp_set_stream) */
int opl_lbfgs_set_stream(opl_lbfgs *l, void *cuda_stream);
/* Checkpoint / teacher forcing: the reference pickles `_prev_param_diffs` / `_prev_jac_diffs`
 * (optimizer.py:397-399).  push appends one (s, y) pair as the NEWEST entry (copied; the oldest is dropped first when
 * the history is full, as _update_diffs does); get copies pair `index` (0 = oldest) out. */
int opl_lbfgs_push_pair_device(opl_lbfgs *l, const double *s_dev, const double *y_dev);
int opl_lbfgs_get_pair_device(const opl_lbfgs *l, int32_t index, double *s_out_dev, double *y_out_dev);
int64_t opl_lbfgs_launch_count(const opl_lbfgs *l);

/* =====================================================================================
 * n-vector helpers for the optimizer stack (float64, device pointers)
 *   replace the NumPy expressions at optimizer.py:97,136-141,251-253,429-431,
 *   linesearch.py:259,397 and initialstep.py:146,223.
 * ===================================================================================== */
/* out = x + fl(a*d) (+ fl(b*p) if p != NULL); products and sums rounded separately (no FMA
 * contraction) so the iterate matches NumPy bit for bit.  out may alias x. */
int opl_vec_axpy_device(int64_t n, const double *x, double a, const double *d, double b,
                        const double *p, double *out, void *cuda_stream);
/* step_out = fl(a*d), out = x + step_out: `self._prev_step = step_size * step_dir; return parameters + self._prev_step`
 * (optimizer.py:251-253, 429-431) in one launch */
int opl_vec_step_device(int64_t n, const double *x, double a, const double *d, double *step_out, double *out,
                        void *cuda_stream);
/* out = a * x */
int opl_vec_scale_device(int64_t n, double a, const double *x, double *out, void *cuda_stream);
/* Weight penalties of RegressionModel (error.py:122-193, regression.py:158-180):
 *   kind 1: L1Penalty  value = lambda * sum|w|,     grad += lambda * sign(w)
 *   kind 2: L2Penalty  value = lambda * ||w||_2,    grad += lambda * w / ||w||_2
 * grad_inout may be NULL (objective only).  host_out[0] = penalty value; synchronises. */
int opl_vec_penalty_device(int64_t n, const double *w, int32_t kind, double lambda, double *grad_inout,
                           double *host_out, void *cuda_stream);
/* host_out[0] = x . y   (deterministic two-stage reduction; synchronises) */
int opl_vec_dot_device(int64_t n, const double *x, const double *y, double *host_out, void *cuda_stream);
/* dst[i][:] = src[index[i]][:] for i < count: the mini-batch selection of Model.stochastic_train /
 * select_sample (base.py:39-50,100-135) on a device-resident dataset.  src: rows x cols row-major,
 * index_dev: count int64 row numbers in [0, rows), dst: count x cols.  Asynchronous. */
int opl_gather_rows_device(const double *src_dev, int64_t rows, int64_t cols, const int64_t *index_dev, int64_t count,
                           double *dst_dev, void *cuda_stream);

/* =====================================================================================
 * RBF network front end.  The RBF output layer is an opl_mlp of shape
 * (clusters, outputs) with a linear transfer fed with the similarity matrix (its flat layout
 * [bias ; W.ravel()], architecture/rbf.py:229-231, is the MLP layout); these two entry points
 * are the parts in front of it.
 * ===================================================================================== */
/* SOM.activate (architecture/som.py:61-85) for a matrix of patterns, optionally followed by
 * calculate.gaussian (calculate.py:101-102) and the similarity scaling of RBF.activate
 * (architecture/rbf.py:139-146):
 *   mode 0: out[i][j] = sqrt(sum_k (x[i][k] - centers[j][k])^2)
 *   mode 1: out[i][j] = exp(-(d^2 / variance))
 *   mode 2: mode 1 divided by its row sum, NaN (0/0) replaced by 1/neurons
 * x: rows x attrs, centers: neurons x attrs, out: rows x neurons, row-major FP64 device buffers. */
int opl_rbf_similarity_device(const double *x_dev, int64_t rows, int32_t attrs, const double *centers_dev,
                              int32_t neurons, int32_t mode, double variance, double *out_dev, void *cuda_stream);
/* `passes` passes of Model.train_step's per-pattern loop (base.py:289-331) around
 * SOM._train_increment / _move_neurons (architecture/som.py:87-113): for every pattern in order,
 * winner = argmin distance, neurons within `neighborhood` of the winner move by
 * rates[|i - winner|] * (pattern - weights[i]).  rates_dev: neighborhood + 1 doubles
 * (calculate.gaussian(distance, neighbor_move_rate) * move_rate, computed by the caller);
 * weights_dev: neurons x attrs, updated in place.  Asynchronous on the stream. */
int opl_som_train_device(const double *x_dev, int64_t rows, int32_t attrs, double *weights_dev, int32_t neurons,
                         int64_t passes, const double *rates_dev, int32_t neighborhood, void *cuda_stream);

/* =====================================================================================
 * Multi-GPU: the all-reduce of the batch-sharded evaluation as one kernel over
 * NVLink peer memory.  The reference has no counterpart (single process); it completes what
 * mlp.py:269-276 sums over ALL patterns when the patterns are sharded over ranks.
 * ===================================================================================== */
/* out[i] = sum over ranks r = 0..world-1 (in that order, so every rank gets the bit-identical
 * result) of buffer_r[i], i < count.  buffer_ptrs[r] / flag_ptrs[r]: this process's peer-mapped
 * device addresses of rank r's vector and of rank r's flag array (32 x uint32, zero-initialised
 * once); host arrays of `world` entries.  `seq` must increase by one per call on every rank and the
 * caller alternates between two buffers by its parity (see csrc/collective.cu).  Asynchronous on
 * the stream; the vector of this rank must have been produced by earlier work on the same stream. */
int opl_allreduce_sum_device(int32_t world, int32_t rank, const uint64_t *buffer_ptrs, const uint64_t *flag_ptrs,
                             int64_t count, uint32_t seq, double *out_dev, void *cuda_stream);

/* The same sum on an NVSwitch system with multicast (NVLS): rank r reduces its 1/world slice in the
 * switch (multimem.ld_reduce), broadcasts it into every rank's result buffer (multimem.st), and after
 * a second flag round every rank copies the complete result to out_dev.  mc_in / mc_res: multicast
 * addresses of this call's input / result buffers; res_local: this rank's result buffer;
 * flag_ptrs[r]: peer-mapped address of rank r's 32 x uint32 flag array (zero-initialised once);
 * counter_dev: one zero-initialised uint32 in local device memory.  seq / parity rules as above. */
int opl_allreduce_sum_nvls_device(int32_t world, int32_t rank, uint64_t mc_in, uint64_t mc_res, const double *res_local,
                                  const uint64_t *flag_ptrs, int64_t count, uint32_t seq, double *out_dev,
                                  uint32_t *counter_dev, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* OPL_B200_H */
