// BFGS inverse-Hessian state: O(n^2), in-place, lazily-applied rank-2 update fused with the
// products the next direction needs.  Replaces optimize/optimizer.py:234-338 (BFGS.next,
// _get_approx_inv_hessian, _bfgs_eq), whose literal evaluation costs 4 n^3 flop and five
// n x n temporaries per iteration.
//
// Algebra (H is NOT assumed symmetric):
//   rho = 1/(y.s)   u = H y   v = H^T y   c = y.u   kappa = rho^2 c + rho
//   H+  = H - rho s v^T - rho u s^T + kappa s s^T
//   H+ g = H g - rho s (v.g) - rho u (s.g) + kappa s (s.g)          (direction without H+)
//
// Schedule: the direction of iteration k only needs H_{k-1}.[y g] and y^T.H_{k-1}, so the update
// that produces H_k stays PENDING (vectors s,u,v + 2 scalars) and is applied by the pass of
// iteration k+1, which reads every H element once, applies the pending update, writes it back
// and accumulates the three products for the new (y, g) on the fly:
//   bytes per iteration = 2 n^2 * 8  (one read + one write of H), HBM-bound.
// The identity / gamma*I seed is never materialised: the first update is pending on a virtual
// diagonal and the first pass writes H without reading it.
#include <algorithm>
#include <cstdlib>

#include "gemm_f64.cuh"   // launch_reduce_partials

namespace opl {

constexpr int PASS_THREADS = 256;
constexpr int PASS_WARPS = PASS_THREADS / 32;
constexpr int PASS_COLS = 128;   // columns per chunk: 32 lanes x 4
constexpr int PASS_ROWS = 4;     // rows in flight per warp iteration

enum { PASS_PLAIN = 0, PASS_UPDATE = 1, PASS_SEED = 2 };

struct PassArgs {
    double *H;            // rows_local x n, row-major
    int64_t n;
    int64_t row_begin;    // global index of local row 0
    int64_t rows_local;
    int64_t rows_per_panel;
    int mode;
    int products;         // accumulate u2/w2/v2 ?
    const double *s, *u, *v;   // pending update vectors (length n)
    const double *coef;        // device: rho, kappa, gamma
    const double *y2, *g2;     // new pair (length n)
    double *u2, *w2;           // out, indexed by global row (col_splits == 1)
    double *vpart;             // out, [panels][n]
    // 2-D decomposition for SHORT shards (row-sharded H over several GPUs): blockIdx.y selects a column range, the row
    // products of the ranges go to upart / wpart [col_splits][rows_local] and are folded in range order afterwards
    int col_splits;
    int64_t cols_per_split;    // multiple of PASS_COLS
    double *upart, *wpart;
};

// Sum 8 per-lane values across the warp with 9 (instead of 40) 64-bit shuffles.
// Afterwards lanes with (lane & 3) == 0 hold the total of value index
//   ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)   in v[0].
__device__ __forceinline__ void warp_reduce8(double (&v)[8], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double send = hi ? v[k] : v[k + 4];
            const double keep = hi ? v[k + 4] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double send = hi ? v[k] : v[k + 2];
            const double keep = hi ? v[k + 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
        const double send = hi ? v[0] : v[1];
        const double keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// One CTA per row panel; streams the panel in chunks of PASS_COLS columns.
// VEC == 2: 16-byte loads/stores (n even, H 16-byte aligned); VEC == 1: scalar.
template <int VEC>
__global__ void __launch_bounds__(PASS_THREADS, 2) bfgs_pass_kernel(const PassArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int64_t R = p.rows_per_panel;
    double *s_row = sm, *u_row = sm + R, *y_row = sm + 2 * R, *acc_u = sm + 3 * R, *acc_w = sm + 4 * R;
    double *red = sm + 5 * R;   // [PASS_WARPS][PASS_COLS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * R;
    const int nr = (int)min(R, p.rows_local - r0);
    const int64_t n = p.n;
    const bool upd = p.mode != PASS_PLAIN;
    const bool prod = p.products != 0;

    for (int r = tid; r < nr; r += PASS_THREADS) {
        const int64_t gi = p.row_begin + r0 + r;
        s_row[r] = upd ? p.s[gi] : 0.0;
        u_row[r] = upd ? p.u[gi] : 0.0;
        y_row[r] = prod ? p.y2[gi] : 0.0;
        acc_u[r] = 0.0;
        acc_w[r] = 0.0;
    }
    double rho = 0.0, kappa = 0.0, gamma = 1.0;
    if (upd) {
        rho = p.coef[0];
        kappa = p.coef[1];
        gamma = p.coef[2];
    }
    __syncthreads();

    const int64_t c_begin = (int64_t)blockIdx.y * p.cols_per_split;
    const int64_t c_end = p.col_splits > 1 ? min(n, c_begin + p.cols_per_split) : n;
    for (int64_t c0 = c_begin; c0 < c_end; c0 += PASS_COLS) {
        int64_t col[4];
        if (VEC == 2) {
            col[0] = c0 + 2 * lane;
            col[1] = col[0] + 1;
            col[2] = c0 + 64 + 2 * lane;
            col[3] = col[2] + 1;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) col[k] = c0 + 32 * k + lane;
        }
        double vc[4], sc[4], yc[4], gc[4], vcol[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool in = col[k] < n;
            vc[k] = (upd && in) ? p.v[col[k]] : 0.0;
            sc[k] = (upd && in) ? p.s[col[k]] : 0.0;
            yc[k] = (prod && in) ? p.y2[col[k]] : 0.0;
            gc[k] = (prod && in) ? p.g2[col[k]] : 0.0;
            vcol[k] = 0.0;
        }

        for (int rb = warp; rb < nr; rb += PASS_WARPS * PASS_ROWS) {
            double h[PASS_ROWS][4];
            // ---- loads (all rows first: memory-level parallelism) ----
#pragma unroll
            for (int q = 0; q < PASS_ROWS; ++q) {
                const int r = rb + PASS_WARPS * q;
                const int64_t gi = p.row_begin + r0 + r;
                double *hp = p.H + (r0 + r) * n;
                if (r < nr && p.mode != PASS_SEED) {
                    if (VEC == 2) {
                        if (col[0] < n) {
                            const double2 t = *reinterpret_cast<const double2 *>(hp + col[0]);
                            h[q][0] = t.x;
                            h[q][1] = t.y;
                        } else h[q][0] = h[q][1] = 0.0;
                        if (col[2] < n) {
                            const double2 t = *reinterpret_cast<const double2 *>(hp + col[2]);
                            h[q][2] = t.x;
                            h[q][3] = t.y;
                        } else h[q][2] = h[q][3] = 0.0;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) h[q][k] = col[k] < n ? hp[col[k]] : 0.0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) h[q][k] = (r < nr && col[k] == gi) ? gamma : 0.0;   // virtual gamma*I
                }
            }
            // ---- update, store, accumulate ----
            double part[8];
#pragma unroll
            for (int q = 0; q < PASS_ROWS; ++q) {
                const int r = rb + PASS_WARPS * q;
                double pu = 0.0, pw = 0.0;
                if (r < nr) {
                    if (upd) {
                        const double si = s_row[r];
                        const double pi = -rho * si;
                        const double qi = kappa * si - rho * u_row[r];
#pragma unroll
                        for (int k = 0; k < 4; ++k) h[q][k] = h[q][k] + pi * vc[k] + qi * sc[k];
                        double *hp = p.H + (r0 + r) * n;
                        if (VEC == 2) {
                            if (col[0] < n) *reinterpret_cast<double2 *>(hp + col[0]) = make_double2(h[q][0], h[q][1]);
                            if (col[2] < n) *reinterpret_cast<double2 *>(hp + col[2]) = make_double2(h[q][2], h[q][3]);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (col[k] < n) hp[col[k]] = h[q][k];
                        }
                    }
                    if (prod) {
                        const double yi = y_row[r];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            pu += h[q][k] * yc[k];
                            pw += h[q][k] * gc[k];
                            vcol[k] += yi * h[q][k];
                        }
                    }
                }
                part[2 * q] = pu;
                part[2 * q + 1] = pw;
            }
            if (prod) {
                warp_reduce8(part, lane);
                if ((lane & 3) == 0) {
                    const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                    const int r = rb + PASS_WARPS * (idx >> 1);
                    if (r < nr) {
                        if (idx & 1) acc_w[r] += part[0];
                        else acc_u[r] += part[0];
                    }
                }
            }
        }

        if (prod) {   // fold the 8 warps' column sums in a fixed order
#pragma unroll
            for (int k = 0; k < 4; ++k) red[warp * PASS_COLS + (int)(col[k] - c0)] = vcol[k];
            __syncthreads();
            if (tid < PASS_COLS && c0 + tid < n) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < PASS_WARPS; ++w) t += red[w * PASS_COLS + tid];
                p.vpart[(int64_t)blockIdx.x * n + c0 + tid] = t;
            }
            __syncthreads();
        }
    }

    if (prod) {
        __syncwarp();
        __syncthreads();
        for (int r = tid; r < nr; r += PASS_THREADS) {
            if (p.col_splits > 1) {
                p.upart[(int64_t)blockIdx.y * p.rows_local + r0 + r] = acc_u[r];
                p.wpart[(int64_t)blockIdx.y * p.rows_local + r0 + r] = acc_w[r];
            } else {
                const int64_t gi = p.row_begin + r0 + r;
                p.u2[gi] = acc_u[r];
                p.w2[gi] = acc_w[r];
            }
        }
    }
}

// row products of a 2-D pass: u2[row] = sum over the column ranges, in range order (bit-reproducible)
__global__ void bfgs_fold_rows_kernel(const double *__restrict__ upart, const double *__restrict__ wpart, int splits,
                                      int64_t rows_local, int64_t row_begin, double *__restrict__ u2, double *__restrict__ w2) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows_local; r += (int64_t)gridDim.x * blockDim.x) {
        double su = 0.0, sw = 0.0;
        for (int k = 0; k < splits; ++k) {
            su += upart[(int64_t)k * rows_local + r];
            sw += wpart[(int64_t)k * rows_local + r];
        }
        u2[row_begin + r] = su;
        w2[row_begin + r] = sw;
    }
}

__global__ void sub_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                           double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] - b[i];
}

__global__ void negate_kernel(int64_t n, const double *__restrict__ a, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = -a[i];
}

// partial[b][0..5] = y.s, y.u, v.g, s.g, y.y, y.g   (u, v may be null on the seed call)
__global__ void bfgs_dots_kernel(int64_t n, const double *__restrict__ s, const double *__restrict__ y,
                                 const double *__restrict__ g, const double *__restrict__ u,
                                 const double *__restrict__ v, double *__restrict__ partial) {
    __shared__ double scratch[33];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double si = s[i], yi = y[i], gi = g[i];
        a[0] += yi * si;
        if (u) a[1] += yi * u[i];
        if (v) a[2] += v[i] * gi;
        a[3] += si * gi;
        a[4] += yi * yi;
        a[5] += yi * gi;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double t = block_sum(a[k], scratch);
        if (threadIdx.x == 0) partial[6 * blockIdx.x + k] = t;
    }
}

// Every block folds the partial dots in the same order, derives (rho, kappa), writes its slice of
// the direction and of the new pending vectors.  seed != 0: H is the virtual gamma*I
// (optimizer.py:148-167,266-271), so u = v = gamma*y, w = gamma*g.
__global__ void bfgs_apply_kernel(int64_t n, int nparts, const double *__restrict__ partial, int seed,
                                  int seed_mode, const double *__restrict__ s, const double *__restrict__ y,
                                  const double *__restrict__ g, const double *__restrict__ u,
                                  const double *__restrict__ v, const double *__restrict__ w,
                                  double *__restrict__ s_pend, double *__restrict__ u_pend,
                                  double *__restrict__ v_pend, double *__restrict__ coef,
                                  double *__restrict__ dir) {
    __shared__ double scratch[33];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] += partial[6 * i + k];
    double d[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k] = block_sum(a[k], scratch);
    const double ys = d[0], sg = d[3], yy = d[4];
    double gamma = 1.0;
    if (seed && seed_mode == OPL_BFGS_SEED_SCALED_IDENTITY)   // optimizer.py:346-362
        gamma = yy == 0.0 ? 1.0 : nan_to_num(ys / yy);
    const double c = seed ? gamma * yy : d[1];
    const double vg = seed ? gamma * d[5] : d[2];   // seed: v = gamma*y
    double rho = 0.0, kappa = 0.0;
    if (ys != 0.0) {   // optimizer.py:328-330: y.s == 0 keeps H unchanged
        rho = 1.0 / ys;
        kappa = rho * rho * c + rho;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        coef[0] = rho;
        coef[1] = kappa;
        coef[2] = gamma;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double si = s[i];
        const double ui = seed ? gamma * y[i] : u[i];
        const double vi = seed ? gamma * y[i] : v[i];
        const double wi = seed ? gamma * g[i] : w[i];
        dir[i] = -(wi - rho * si * vg - rho * ui * sg + kappa * si * sg);
        s_pend[i] = si;
        u_pend[i] = ui;
        v_pend[i] = vi;
    }
}

// ---------------------------------------------------------------------------------------
// Small problems (n <= BFGS_SMALL_N, the reference's make_optimizer picks BFGS up to 500 weights, optimizer.py:33-43):
// the whole of _get_approx_inv_hessian + `-(H.dot(jac))` (optimizer.py:242-285) as ONE single-CTA launch instead of five --
// y = g - g_prev, the six dots, u = H y, w = H g (a warp per row), v = H^T y (a thread per column), the rank-2 update
// applied in place at once (nothing stays pending), the direction, and its slope g.d, which also goes to mapped pinned
// memory so that the line search starts without a separate dot product.  Same formulas and the same element-wise update
// expression as the streaming pass, so both paths round alike.
// ---------------------------------------------------------------------------------------
constexpr int BFGS_SMALL_N = 512;
constexpr int BFGS_SMALL_THREADS = 512;

struct SmallBfgsArgs {
    double *H;
    int n;
    int seed;                 // 1: no H yet -> virtual gamma*I is updated and materialised
    int seed_mode;
    const double *g, *g_prev, *s;
    double *dir;
    double *slope_dev;        // device: g.d
    volatile double *slope_host;   // mapped pinned: g.d, sequence number
    double seq;
};

__global__ void __launch_bounds__(BFGS_SMALL_THREADS) bfgs_small_kernel(const SmallBfgsArgs p) {
    __shared__ double ys_[BFGS_SMALL_N], ss_[BFGS_SMALL_N], gs_[BFGS_SMALL_N], us_[BFGS_SMALL_N], vs_[BFGS_SMALL_N], ws_[BFGS_SMALL_N];
    __shared__ double scratch[33];
    const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BFGS_SMALL_THREADS / 32;
    for (int i = tid; i < n; i += BFGS_SMALL_THREADS) {
        const double gi = p.g[i];
        gs_[i] = gi;
        ys_[i] = gi - p.g_prev[i];
        ss_[i] = p.s[i];
    }
    __syncthreads();
    double a0 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0;
    for (int i = tid; i < n; i += BFGS_SMALL_THREADS) {
        a0 += ys_[i] * ss_[i];
        a3 += ss_[i] * gs_[i];
        a4 += ys_[i] * ys_[i];
        a5 += ys_[i] * gs_[i];
    }
    const double ys = block_sum(a0, scratch), sg = block_sum(a3, scratch), yy = block_sum(a4, scratch), yg = block_sum(a5, scratch);
    double gamma = 1.0, c, vg;
    if (p.seed) {
        if (p.seed_mode == OPL_BFGS_SEED_SCALED_IDENTITY) gamma = yy == 0.0 ? 1.0 : nan_to_num(ys / yy);   // optimizer.py:346-362
        for (int i = tid; i < n; i += BFGS_SMALL_THREADS) {
            us_[i] = vs_[i] = gamma * ys_[i];
            ws_[i] = gamma * gs_[i];
        }
        c = gamma * yy;
        vg = gamma * yg;
    } else {
        for (int i = warp; i < n; i += nwarps) {          // u = H y, w = H g: a warp per row
            const double *row = p.H + (size_t)i * n;
            double pu = 0.0, pw = 0.0;
            for (int j = lane; j < n; j += 32) {
                const double h = row[j];
                pu += h * ys_[j];
                pw += h * gs_[j];
            }
            pu = warp_sum(pu);
            pw = warp_sum(pw);
            if (lane == 0) {
                us_[i] = pu;
                ws_[i] = pw;
            }
        }
        for (int j = tid; j < n; j += BFGS_SMALL_THREADS) {   // v = H^T y: a thread per column
            double pv = 0.0;
            for (int i = 0; i < n; ++i) pv += ys_[i] * p.H[(size_t)i * n + j];
            vs_[j] = pv;
        }
        __syncthreads();
        double a1 = 0.0, a2 = 0.0;
        for (int i = tid; i < n; i += BFGS_SMALL_THREADS) {
            a1 += ys_[i] * us_[i];
            a2 += vs_[i] * gs_[i];
        }
        c = block_sum(a1, scratch);
        vg = block_sum(a2, scratch);
    }
    __syncthreads();
    double rho = 0.0, kappa = 0.0;
    if (ys != 0.0) {                                       // optimizer.py:328-330: y.s == 0 keeps H unchanged
        rho = 1.0 / ys;
        kappa = rho * rho * c + rho;
    }
    // H+ = H - rho s v^T - rho u s^T + kappa s s^T, element-wise as the streaming pass writes it
    if (p.seed || ys != 0.0) {
        const int total = n * n;
        for (int e = tid; e < total; e += BFGS_SMALL_THREADS) {
            const int i = e / n, j = e - i * n;
            const double h = p.seed ? (i == j ? gamma : 0.0) : p.H[e];
            const double pi = -rho * ss_[i], qi = kappa * ss_[i] - rho * us_[i];
            p.H[e] = h + pi * vs_[j] + qi * ss_[j];
        }
    }
    double gd = 0.0;
    for (int i = tid; i < n; i += BFGS_SMALL_THREADS) {
        const double d = -(ws_[i] - rho * ss_[i] * vg - rho * us_[i] * sg + kappa * ss_[i] * sg);
        p.dir[i] = d;
        gd += gs_[i] * d;
    }
    gd = block_sum(gd, scratch);
    if (tid == 0) {
        p.slope_dev[0] = gd;
        if (p.slope_host) {
            p.slope_host[0] = gd;
            __threadfence_system();
            p.slope_host[1] = p.seq;
        }
    }
}

__global__ void bfgs_fill_kernel(double *H, int64_t n, int64_t row_begin, int64_t rows_local, uint64_t seed) {
    const int64_t total = rows_local * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_begin + i / n, c = i % n;
        const int64_t lo = r < c ? r : c, hi = r < c ? c : r;
        uint64_t z = seed + (uint64_t)lo * 0x9E3779B97F4A7C15ull + (uint64_t)hi * 0xBF58476D1CE4E5B9ull;
        z ^= z >> 31;
        z *= 0x94D049BB133111EBull;
        z ^= z >> 29;
        const double noise = ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 1e-3;
        H[i] = (r == c ? 1.0 : 0.0) + noise;   // symmetric, diagonally dominant
    }
}

}  // namespace opl

using namespace opl;

struct opl_bfgs {
    int64_t n = 0, row_begin = 0, row_end = 0;
    cudaStream_t stream = nullptr;
    int state = 0;          // 0 none, 1 seed pending, 2 H resident
    bool pending = false;   // an update is waiting to be applied
    double *H = nullptr;
    double *y = nullptr;    // current y = g - g_prev
    double *s_pend = nullptr, *u_pend = nullptr, *v_pend = nullptr;
    double *u = nullptr, *v = nullptr, *w = nullptr;   // single-device scratch for direction_device
    double *coef = nullptr;
    double *vpart = nullptr;
    double *dots = nullptr;
    int dot_blocks = 0;
    int64_t panels = 0, rows_per_panel = 0;
    int col_splits = 1;
    int64_t cols_per_split = 0;
    double *upart = nullptr, *wpart = nullptr;
    int64_t launches = 0;
    // small-problem path: slope g.d of the last direction through mapped pinned memory
    double *slope_dev = nullptr;
    double *mapped_host = nullptr, *mapped_dev = nullptr;
    double mapped_seq = 0.0;
    bool slope_valid = false;
};

namespace {

int vgrid(int64_t n) {
    int64_t g = std::min<int64_t>(ceil_div(n, 256), 4 * (int64_t)sm_count());
    return (int)std::max<int64_t>(g, 1);
}

int ensure_matrix(opl_bfgs *b) {
    const int64_t rows = b->row_end - b->row_begin;
    if (!b->H) OPL_CUDA(cudaMalloc(&b->H, sizeof(double) * (size_t)rows * (size_t)b->n));
    if (!b->vpart) OPL_CUDA(cudaMalloc(&b->vpart, sizeof(double) * (size_t)b->panels * (size_t)b->n));
    if (b->col_splits > 1 && !b->upart) {
        OPL_CUDA(cudaMalloc(&b->upart, sizeof(double) * (size_t)b->col_splits * (size_t)rows));
        OPL_CUDA(cudaMalloc(&b->wpart, sizeof(double) * (size_t)b->col_splits * (size_t)rows));
    }
    return OPL_OK;
}

// Fused pass over the local rows: apply the pending update (if any), accumulate products
int run_pass(opl_bfgs *b, const double *y2, const double *g2, bool products, double *u2, double *v2, double *w2) {
    int rc = ensure_matrix(b);
    if (rc != OPL_OK) return rc;
    PassArgs a = {};
    a.H = b->H;
    a.n = b->n;
    a.row_begin = b->row_begin;
    a.rows_local = b->row_end - b->row_begin;
    a.rows_per_panel = b->rows_per_panel;
    a.mode = b->state == 1 ? PASS_SEED : (b->pending ? PASS_UPDATE : PASS_PLAIN);
    a.products = products ? 1 : 0;
    a.s = b->s_pend;
    a.u = b->u_pend;
    a.v = b->v_pend;
    a.coef = b->coef;
    a.y2 = y2;
    a.g2 = g2;
    a.u2 = u2;
    a.w2 = w2;
    a.vpart = b->vpart;
    a.col_splits = b->col_splits;
    a.cols_per_split = b->cols_per_split;
    a.upart = b->upart;
    a.wpart = b->wpart;
    const size_t smem = sizeof(double) * (size_t)(5 * b->rows_per_panel + PASS_WARPS * PASS_COLS);
    const bool vec = (b->n % 2 == 0) && ((reinterpret_cast<uintptr_t>(b->H) & 15u) == 0);
    const dim3 grid((unsigned)b->panels, (unsigned)b->col_splits);
    if (vec) bfgs_pass_kernel<2><<<grid, PASS_THREADS, smem, b->stream>>>(a);
    else bfgs_pass_kernel<1><<<grid, PASS_THREADS, smem, b->stream>>>(a);
    OPL_LAUNCH_CHECK();
    ++b->launches;
    if (products && b->col_splits > 1) {
        bfgs_fold_rows_kernel<<<vgrid(a.rows_local), 256, 0, b->stream>>>(b->upart, b->wpart, b->col_splits, a.rows_local,
                                                                           b->row_begin, u2, w2);
        OPL_LAUNCH_CHECK();
        ++b->launches;
    }
    if (products) {
        rc = launch_reduce_partials(b->vpart, (int)b->panels, b->n, b->n, v2, b->stream, &b->launches);
        if (rc != OPL_OK) return rc;
    }
    b->state = 2;
    b->pending = false;
    return OPL_OK;
}

int compute_y(opl_bfgs *b, const double *grad, const double *prev_grad) {
    sub_kernel<<<vgrid(b->n), 256, 0, b->stream>>>(b->n, grad, prev_grad, b->y);
    OPL_LAUNCH_CHECK();
    ++b->launches;
    return OPL_OK;
}

int finish(opl_bfgs *b, const double *grad, const double *prev_step, int seed_mode, const double *u,
           const double *v, const double *w, double *dir) {
    const bool seed = b->state == 0;
    bfgs_dots_kernel<<<b->dot_blocks, 256, 0, b->stream>>>(b->n, prev_step, b->y, grad, seed ? nullptr : u,
                                                           seed ? nullptr : v, b->dots);
    OPL_LAUNCH_CHECK();
    bfgs_apply_kernel<<<b->dot_blocks, 256, 0, b->stream>>>(b->n, b->dot_blocks, b->dots, seed ? 1 : 0, seed_mode,
                                                            prev_step, b->y, grad, u, v, w, b->s_pend, b->u_pend,
                                                            b->v_pend, b->coef, dir);
    OPL_LAUNCH_CHECK();
    b->launches += 2;
    if (seed) b->state = 1;
    b->pending = true;
    return OPL_OK;
}

}  // namespace

extern "C" {

int opl_bfgs_create(opl_bfgs **out, int64_t n, int64_t row_begin, int64_t row_end, void *cuda_stream) {
    if (!out || n < 1 || row_begin < 0 || row_end > n || row_begin >= row_end)
        return fail(OPL_E_ARG, "opl_bfgs_create: bad arguments");
    if (opl_device_count() < 1) return fail(OPL_E_CUDA, "no CUDA device: opl_b200 has no CPU fallback");
    opl_bfgs *b = new opl_bfgs();
    b->n = n;
    b->row_begin = row_begin;
    b->row_end = row_end;
    b->stream = (cudaStream_t)cuda_stream;
    const int64_t rows = row_end - row_begin;
    const int64_t slots = 2 * (int64_t)sm_count();        // two resident CTAs per SM
    int64_t panels = std::min<int64_t>(ceil_div(rows, 8), slots);
    int64_t rpp = ceil_div(ceil_div(rows, panels), 8) * 8;
    if (rpp > 1024) rpp = 1024;
    b->rows_per_panel = rpp;
    b->panels = ceil_div(rows, rpp);
    // A SHORT shard (H row-sharded over several GPUs: 10416 rows per GPU at n = 83328 on 8) would give every CTA a panel of
    // only ~40 rows: the per-chunk synchronisation and the column-sum fold are then amortised over 40 x 128 elements, a
    // third of the row slots of the last warp iteration is empty and [panels][n] column partials cost 5 % extra traffic
    // (measured: 0.73 of the aggregate HBM rate at 8 GPUs against 0.96 on one).  Tall panels x column ranges instead.
    if (rpp < 192 && n >= 8 * PASS_COLS && rows >= 64) {
        const int64_t tall = 288;                           // = the single-GPU panel height at n = 83328
        const int64_t panels_r = ceil_div(rows, tall);
        int64_t splits = std::max<int64_t>(1, std::min<int64_t>(slots / panels_r, n / (4 * PASS_COLS)));
        if (splits > 1) {
            b->rows_per_panel = tall;
            b->panels = panels_r;
            b->col_splits = (int)splits;
            b->cols_per_split = ceil_div(ceil_div(n, splits), PASS_COLS) * PASS_COLS;
            b->col_splits = (int)ceil_div(n, b->cols_per_split);
        }
    }
    b->dot_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 1024), (int64_t)sm_count()));
    cudaError_t e = cudaSuccess;
    double **vecs[] = {&b->y, &b->s_pend, &b->u_pend, &b->v_pend, &b->u, &b->v, &b->w};
    for (double **p : vecs)
        if (e == cudaSuccess) e = cudaMalloc(p, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&b->coef, sizeof(double) * 4);
    if (e == cudaSuccess) e = cudaMemset(b->coef, 0, sizeof(double) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b->dots, sizeof(double) * 6 * (size_t)b->dot_blocks);
    if (e == cudaSuccess) e = cudaMalloc(&b->slope_dev, sizeof(double) * 2);
    if (e == cudaSuccess) e = cudaHostAlloc(&b->mapped_host, sizeof(double) * 2, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        b->mapped_host[0] = b->mapped_host[1] = 0.0;
        e = cudaHostGetDevicePointer(&b->mapped_dev, b->mapped_host, 0);
    }
    if (e != cudaSuccess) {
        opl_bfgs_destroy(b);
        return fail(OPL_E_CUDA, "opl_bfgs_create: %s", cudaGetErrorString(e));
    }
    *out = b;
    return OPL_OK;
}

int opl_bfgs_destroy(opl_bfgs *b) {
    if (!b) return OPL_OK;
    cudaStreamSynchronize(b->stream);
    double *all[] = {b->H, b->y, b->s_pend, b->u_pend, b->v_pend, b->u, b->v, b->w, b->coef, b->vpart, b->dots, b->upart, b->wpart};
    for (double *p : all) cudaFree(p);
    cudaFree(b->slope_dev);
    cudaFreeHost(b->mapped_host);
    delete b;
    return OPL_OK;
}

int opl_bfgs_reset(opl_bfgs *b) {
    if (!b) return fail(OPL_E_ARG, "opl_bfgs_reset: null");
    b->state = 0;
    b->pending = false;
    return OPL_OK;
}

int opl_bfgs_set_stream(opl_bfgs *b, void *cuda_stream) {
    if (!b) return fail(OPL_E_ARG, "opl_bfgs_set_stream: null");
    if (b->stream != (cudaStream_t)cuda_stream) {
        OPL_CUDA(cudaStreamSynchronize(b->stream));
        b->stream = (cudaStream_t)cuda_stream;
    }
    return OPL_OK;
}

int opl_bfgs_state(const opl_bfgs *b) { return b ? b->state : -1; }
int64_t opl_bfgs_launch_count(const opl_bfgs *b) { return b ? b->launches : -1; }

int opl_bfgs_pass_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                         const double *prev_grad_dev, int32_t seed_mode, double *u_dev, double *v_dev,
                         double *w_dev) {
    (void)seed_mode;
    if (!b || !grad_dev) return fail(OPL_E_ARG, "opl_bfgs_pass_device: null pointer");
    if (!prev_step_dev) return OPL_OK;   // first iteration: nothing to do
    if (!prev_grad_dev || !u_dev || !v_dev || !w_dev) return fail(OPL_E_ARG, "opl_bfgs_pass_device: null pointer");
    int rc = compute_y(b, grad_dev, prev_grad_dev);
    if (rc != OPL_OK) return rc;
    if (b->state == 0) return OPL_OK;    // second iteration: seed handled from vectors in finish
    return run_pass(b, b->y, grad_dev, true, u_dev, v_dev, w_dev);
}

int opl_bfgs_finish_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                           const double *prev_grad_dev, int32_t seed_mode, const double *u_dev,
                           const double *v_dev, const double *w_dev, double *dir_out_dev) {
    (void)prev_grad_dev;
    if (!b || !grad_dev || !dir_out_dev) return fail(OPL_E_ARG, "opl_bfgs_finish_device: null pointer");
    if (!prev_step_dev) {   // optimizer.py:258-263: identity on the first call => dir = -grad
        negate_kernel<<<vgrid(b->n), 256, 0, b->stream>>>(b->n, grad_dev, dir_out_dev);
        OPL_LAUNCH_CHECK();
        ++b->launches;
        return OPL_OK;
    }
    return finish(b, grad_dev, prev_step_dev, seed_mode, u_dev, v_dev, w_dev, dir_out_dev);
}

int opl_bfgs_direction_device(opl_bfgs *b, const double *grad_dev, const double *prev_step_dev,
                              const double *prev_grad_dev, int32_t seed_mode, double *dir_out_dev) {
    if (!b) return fail(OPL_E_ARG, "opl_bfgs_direction_device: null");
    if (b->row_begin != 0 || b->row_end != b->n)
        return fail(OPL_E_STATE, "opl_bfgs_direction_device needs the whole H on this device; use pass/finish");
    b->slope_valid = false;
    static int small_enabled = -1;
    if (small_enabled < 0) {
        const char *env = getenv("OPL_SMALL_PATH");
        small_enabled = (env && atoi(env) == 0) ? 0 : 1;
    }
    if (small_enabled && b->n <= BFGS_SMALL_N && grad_dev && dir_out_dev && prev_step_dev && prev_grad_dev) {
        if (b->state != 0 && b->pending) {      // (an update left pending by the streaming path: apply it first)
            int rc = run_pass(b, nullptr, nullptr, false, nullptr, nullptr, nullptr);
            if (rc != OPL_OK) return rc;
        }
        int rc = ensure_matrix(b);
        if (rc != OPL_OK) return rc;
        SmallBfgsArgs a = {};
        a.H = b->H;
        a.n = (int)b->n;
        a.seed = b->state == 0 ? 1 : 0;
        a.seed_mode = seed_mode;
        a.g = grad_dev;
        a.g_prev = prev_grad_dev;
        a.s = prev_step_dev;
        a.dir = dir_out_dev;
        a.slope_dev = b->slope_dev;
        a.slope_host = b->mapped_dev;
        b->mapped_seq += 1.0;
        a.seq = b->mapped_seq;
        bfgs_small_kernel<<<1, BFGS_SMALL_THREADS, 0, b->stream>>>(a);
        OPL_LAUNCH_CHECK();
        ++b->launches;
        b->state = 2;
        b->pending = false;
        b->slope_valid = true;
        return OPL_OK;
    }
    int rc = opl_bfgs_pass_device(b, grad_dev, prev_step_dev, prev_grad_dev, seed_mode, b->u, b->v, b->w);
    if (rc != OPL_OK) return rc;
    return opl_bfgs_finish_device(b, grad_dev, prev_step_dev, prev_grad_dev, seed_mode, b->u, b->v, b->w,
                                  dir_out_dev);
}

int opl_bfgs_last_slope(opl_bfgs *b, double *host_out) {
    if (!b || !host_out) return fail(OPL_E_ARG, "opl_bfgs_last_slope: null pointer");
    if (!b->slope_valid) return fail(OPL_E_STATE, "the last direction did not come with its slope (streaming path)");
    volatile double *m = b->mapped_host;
    bool seen = false;
    for (int spin = 0; spin < 2000000 && !seen; ++spin) seen = m[1] == b->mapped_seq;
    if (!seen) {
        OPL_CUDA(cudaStreamSynchronize(b->stream));
        if (m[1] != b->mapped_seq) return fail(OPL_E_CUDA, "the direction kernel did not publish its slope");
    }
    *host_out = m[0];
    return OPL_OK;
}

int opl_bfgs_materialize_device(opl_bfgs *b, double *h_out_dev) {
    if (!b || !h_out_dev) return fail(OPL_E_ARG, "opl_bfgs_materialize_device: null pointer");
    if (b->state == 0) return fail(OPL_E_STATE, "no inverse Hessian yet (fewer than two BFGS iterations)");
    if (b->state == 1 || b->pending) {
        int rc = run_pass(b, nullptr, nullptr, false, nullptr, nullptr, nullptr);
        if (rc != OPL_OK) return rc;
    }
    const size_t bytes = sizeof(double) * (size_t)(b->row_end - b->row_begin) * (size_t)b->n;
    OPL_CUDA(cudaMemcpyAsync(h_out_dev, b->H, bytes, cudaMemcpyDeviceToDevice, b->stream));
    return OPL_OK;
}

int opl_bfgs_load_device(opl_bfgs *b, const double *h_dev) {
    if (!b || !h_dev) return fail(OPL_E_ARG, "opl_bfgs_load_device: null pointer");
    int rc = ensure_matrix(b);
    if (rc != OPL_OK) return rc;
    const size_t bytes = sizeof(double) * (size_t)(b->row_end - b->row_begin) * (size_t)b->n;
    OPL_CUDA(cudaMemcpyAsync(b->H, h_dev, bytes, cudaMemcpyDeviceToDevice, b->stream));
    b->state = 2;
    b->pending = false;
    return OPL_OK;
}

int opl_bfgs_fill_synthetic(opl_bfgs *b, uint64_t seed) {
    if (!b) return fail(OPL_E_ARG, "opl_bfgs_fill_synthetic: null");
    int rc = ensure_matrix(b);
    if (rc != OPL_OK) return rc;
    bfgs_fill_kernel<<<8 * sm_count(), 256, 0, b->stream>>>(b->H, b->n, b->row_begin, b->row_end - b->row_begin, seed);
    OPL_LAUNCH_CHECK();
    ++b->launches;
    b->state = 2;
    b->pending = false;
    return OPL_OK;
}

}  // extern "C"
