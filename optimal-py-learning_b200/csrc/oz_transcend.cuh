// Table-driven FP64 softplus + logistic for the forward GEMM epilogue (reference: ReluTransfer = softplus,
// learning/calculate.py:128-141: log(1.0 + exp(z)) with overflowing entries replaced by z; derivative
// 1 / (1 + exp(-z)) from the INPUT).
//
// Why it is written out: on B200 the FP64 work of the epilogue does not hide under the tensor pipe, so its
// cost is its FP64 instruction count.  Library exp + log + division are ~200 FP64 instructions per element, the
// polynomial-only version this replaces was ~60; this one is ~26:
//   t  = exp(-|z|)         n = rint(-|z| 64/ln2), r = -|z| - n ln2/64 (two-part constant), 2^(n/64) from a 64-entry table,
//                          exp(r) - 1 by a degree-5 Taylor polynomial (|r| <= ln2/128: remainder 3.5e-17)      10 FP64
//   u  = 1 + t             in [1, 2]  (rounded exactly like the reference's 1.0 + exp(z) for z <= 0)           1
//   1/u                    MUFU seed (2^-20) + one cubic Newton step                                           3
//   log u                  i = top 8 mantissa bits of u, w = u rc_i - 1 (one FMA, |w| <= 2^-8), degree-7 Taylor of log1p(w),
//                          + L_i = -log(rc_i)                                                                  9
//   softplus = max(z,0) + log u,  logistic = z >= 0 ? 1/u : t/u                                                2
// Selections and special cases (|z| >= 708, NaN) are integer operations on the high words, not FP64 compares.
// Errors ~1 ulp (tests/test_host_logic.py compiles this header for the host and checks it against mpmath).
//
// The header compiles for the device (nvcc) and for the host (g++, used only by that test).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OZT_FN __device__ __forceinline__
#define OZT_TABLE_QUAL __device__ const __align__(16)
#else
#include <math.h>
#include <string.h>
#define OZT_FN static inline
#define OZT_TABLE_QUAL static const __attribute__((aligned(16)))
#endif

#include "oz_tables.inc"

namespace opl {

#if defined(__CUDACC__)
OZT_FN int ozt_hi(double x) { return __double2hiint(x); }
OZT_FN int ozt_lo(double x) { return __double2loint(x); }
OZT_FN double ozt_make(int hi, int lo) { return __hiloint2double(hi, lo); }
OZT_FN double ozt_fma(double a, double b, double c) { return fma(a, b, c); }
OZT_FN double ozt_rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
OZT_FN double ozt_exp_tab(int j) { return __longlong_as_double((long long)__ldg(&OZT_EXP_BITS[j])); }
OZT_FN void ozt_log_tab(int i, double &rc, double &L) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(&OZT_LOG_BITS[i][0]));
    rc = __longlong_as_double((long long)v.x);
    L = __longlong_as_double((long long)v.y);
}
OZT_FN void ozt_logh_tab(int i, double &rc, double &L) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(&OZT_LOGH_BITS[i][0]));
    rc = __longlong_as_double((long long)v.x);
    L = __longlong_as_double((long long)v.y);
}
#else
OZT_FN int ozt_hi(double x) { uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
OZT_FN int ozt_lo(double x) { uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)b; }
OZT_FN double ozt_make(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    memcpy(&x, &b, 8);
    return x;
}
OZT_FN double ozt_fma(double a, double b, double c) { return fma(a, b, c); }
// host model of MUFU.RCP64H: the low 32 bits of the operand are ignored, the result carries ~20 good bits
OZT_FN double ozt_rcp_seed(double x) { return ozt_make(ozt_hi(1.0 / ozt_make(ozt_hi(x), 0)), 0); }
OZT_FN double ozt_exp_tab(int j) { double x; memcpy(&x, &OZT_EXP_BITS[j], 8); return x; }
OZT_FN void ozt_log_tab(int i, double &rc, double &L) { memcpy(&rc, &OZT_LOG_BITS[i][0], 8); memcpy(&L, &OZT_LOG_BITS[i][1], 8); }
OZT_FN void ozt_logh_tab(int i, double &rc, double &L) { memcpy(&rc, &OZT_LOGH_BITS[i][0], 8); memcpy(&L, &OZT_LOGH_BITS[i][1], 8); }
#endif

// ---- single-value forms for latency-bound row stages (fused head: softmax + cross entropy) ---------------------------
// Short dependency chains are the point there: ~11 dependent FP64 operations for exp, ~10 for log, 3 for 1/x, against
// ~40 / ~45 / ~12 for the library routines (a dependent FP64 operation costs ~36 cycles on this part).

// exp(x) for x <= 0 (softmax: logit minus the row maximum), -inf -> 0, NaN -> NaN; subnormal results are rounded by the
// hardware (two-step scaling), as the reference's exp does
OZT_FN double ozt_exp_nonpos(double x) {
    const double t = ozt_fma(x, 92.332482616893656877, 6755399441055744.0);
    const double nd = t - 6755399441055744.0;
    double r = ozt_fma(nd, -6.93147180369123816490e-01 / 64.0, x);
    r = ozt_fma(nd, -1.90821492927058770002e-10 / 64.0, r);
    const double r2 = r * r;
    double p = ozt_fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = ozt_fma(p, r, 1.6666666666666666e-01);
    p = ozt_fma(p, r, 0.5);
    const double s = ozt_fma(r2, p, r);
    const int n = ozt_lo(t), k = n >> 6;
    const double T = ozt_exp_tab(n & 63);
    const double q = ozt_fma(T, s, T);
    const bool sub = k < -1000;
    const double v = ozt_make(ozt_hi(q) + ((k + (sub ? 128 : 0)) << 20), ozt_lo(q)) * (sub ? 2.938735877055719e-39 : 1.0);   // 2^-128
    const uint32_t ax = (uint32_t)(ozt_hi(x) & 0x7fffffff);
    const bool isnan = (ax | (ozt_lo(x) != 0 ? 1u : 0u)) > 0x7ff00000u;
    return isnan ? x : (ax >= 0x40875000u ? 0.0 : v);       // |x| >= 746: below the smallest subnormal
}

// 1/a for a normal a whose reciprocal is finite: MUFU seed (2^-20) + one cubic Newton step, ~1 ulp
OZT_FN double ozt_rcp(double a) {
    const double y = ozt_rcp_seed(a);
    const double e = ozt_fma(-a, y, 1.0);
    return ozt_fma(y, ozt_fma(e, e, e), y);
}

// log(a) for a positive finite a (normal or subnormal): a = u 2^(kk-1), u in [1, 2); log a = kk ln2 + (log u - ln2) with the
// bracket from the table centred on a = 1 (relative accuracy kept for a -> 1-, where cross entropy lives)
OZT_FN double ozt_log_pos(double a) {
    int hi = ozt_hi(a), adj = 0;
    if (hi < 0x00100000) {                 // subnormal: rare
        a *= 18014398509481984.0;          // 2^54
        hi = ozt_hi(a);
        adj = -54;
    }
    const int kk = (hi >> 20) - 1022 + adj;
    const double u = ozt_make((hi & 0x000fffff) | 0x3ff00000, ozt_lo(a));
    double rc, L;
    ozt_logh_tab((hi >> 12) & 0xff, rc, L);
    const double w = ozt_fma(u, rc, -1.0), w2 = w * w;
    double p = ozt_fma(w, 1.4285714285714285e-01, -1.6666666666666666e-01);
    p = ozt_fma(p, w, 0.2);
    p = ozt_fma(p, w, -0.25);
    p = ozt_fma(p, w, 3.3333333333333331e-01);
    p = ozt_fma(p, w, -0.5);
    const double lw = ozt_fma(w2, p, w);
    const double kd = (double)kk;
    return ozt_fma(kd, 6.93147180369123816490e-01, L + ozt_fma(kd, 1.90821492927058770002e-10, lw));
}

// W elements advance in lockstep through every step, so W independent FP64 chains are adjacent in program order
// (a dependent DFMA has ~36 cycles of latency on this part).
template <int W>
OZT_FN void ozt_softplus_logistic_n(const double (&z)[W], double (&sp)[W], double (&sg)[W]) {
    double a[W], t[W], r[W], s[W], u[W], y[W];
    // ---- t = exp(-|z|) ----
#pragma unroll
    for (int i = 0; i < W; ++i) a[i] = -fabs(z[i]);
#pragma unroll
    for (int i = 0; i < W; ++i) t[i] = ozt_fma(a[i], 92.332482616893656877, 6755399441055744.0);   // 64/ln2; low word = n
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = t[i] - 6755399441055744.0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double nd = r[i];
        r[i] = ozt_fma(nd, -6.93147180369123816490e-01 / 64.0, a[i]);     // ln2_hi / 64: 21 trailing zero bits, n ln2_hi/64 exact
        r[i] = ozt_fma(nd, -1.90821492927058770002e-10 / 64.0, r[i]);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double x = r[i], x2 = x * x;
        double p = ozt_fma(x, 8.3333333333333332e-03, 4.1666666666666664e-02);    // 1/5! x + 1/4!
        p = ozt_fma(p, x, 1.6666666666666666e-01);
        p = ozt_fma(p, x, 0.5);
        s[i] = ozt_fma(x2, p, x);                                                 // exp(r) - 1
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const int n = ozt_lo(t[i]);
        const double T = ozt_exp_tab(n & 63);
        const double q = ozt_fma(T, s[i], T);                                     // 2^(j/64) exp(r) in [0.5, 1]
        const double scaled = ozt_make(ozt_hi(q) + ((n >> 6) << 20), ozt_lo(q));   // * 2^(n >> 6), no underflow above -708
        const bool tiny = (uint32_t)(ozt_hi(z[i]) & 0x7fffffff) >= 0x40862000u;   // |z| >= 708 (inf, NaN included)
        t[i] = tiny ? 0.0 : scaled;
    }
    // ---- u = 1 + t in [1, 2];  1/u ----
#pragma unroll
    for (int i = 0; i < W; ++i) u[i] = 1.0 + t[i];
#pragma unroll
    for (int i = 0; i < W; ++i) y[i] = ozt_rcp_seed(u[i]);
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double e = ozt_fma(-u[i], y[i], 1.0);
        y[i] = ozt_fma(y[i], ozt_fma(e, e, e), y[i]);
    }
    // ---- log u ----
#pragma unroll
    for (int i = 0; i < W; ++i) {
        uint32_t idx = (uint32_t)(ozt_hi(u[i]) - 0x3ff00000) >> 12;
        idx = idx > 256u ? 256u : idx;
        double rc, L;
        ozt_log_tab((int)idx, rc, L);
        const double w = ozt_fma(u[i], rc, -1.0), w2 = w * w;
        double p = ozt_fma(w, 1.4285714285714285e-01, -1.6666666666666666e-01);    // w/7 - 1/6
        p = ozt_fma(p, w, 0.2);
        p = ozt_fma(p, w, -0.25);
        p = ozt_fma(p, w, 3.3333333333333331e-01);
        p = ozt_fma(p, w, -0.5);
        r[i] = L + ozt_fma(w2, p, w);
    }
    // ---- assemble ----
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const int hz = ozt_hi(z[i]);
        const bool neg = hz < 0;
        const bool isnan = ((uint32_t)(hz & 0x7fffffff) | (ozt_lo(z[i]) != 0 ? 1u : 0u)) > 0x7ff00000u;
        const double spv = (neg ? 0.0 : z[i]) + r[i];
        const double ty = t[i] * y[i];
        const double sgv = neg ? ty : y[i];
        sp[i] = isnan ? z[i] : spv;
        sg[i] = isnan ? z[i] : sgv;
    }
}

// The same pair for REGULAR arguments only (finite, |z| < 708): no special-value selections, and max(z, 0) formed
// arithmetically (z + |z| is 2 max(z, 0) exactly).  The GEMM epilogue runs this form and re-evaluates a whole register
// group with ozt_softplus_logistic_n when any of its values is not regular (warp-uniform, practically never taken):
// the finish phase is bound by its instruction count, and the selections were a quarter of the non-FP64 instructions.
OZT_FN bool ozt_is_regular(double z) { return (uint32_t)(ozt_hi(z) & 0x7fffffff) < 0x40862000u; }   // |z| < 708, not inf / NaN

template <int W>
OZT_FN void ozt_softplus_logistic_regular_n(const double (&z)[W], double (&sp)[W], double (&sg)[W]) {
    double a[W], t[W], r[W], s[W], u[W], y[W];
#pragma unroll
    for (int i = 0; i < W; ++i) a[i] = -fabs(z[i]);
#pragma unroll
    for (int i = 0; i < W; ++i) t[i] = ozt_fma(a[i], 92.332482616893656877, 6755399441055744.0);
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = t[i] - 6755399441055744.0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double nd = r[i];
        r[i] = ozt_fma(nd, -6.93147180369123816490e-01 / 64.0, a[i]);
        r[i] = ozt_fma(nd, -1.90821492927058770002e-10 / 64.0, r[i]);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double x = r[i], x2 = x * x;
        double p = ozt_fma(x, 8.3333333333333332e-03, 4.1666666666666664e-02);
        p = ozt_fma(p, x, 1.6666666666666666e-01);
        p = ozt_fma(p, x, 0.5);
        s[i] = ozt_fma(x2, p, x);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const int n = ozt_lo(t[i]);
        const double T = ozt_exp_tab(n & 63);
        const double q = ozt_fma(T, s[i], T);
        t[i] = ozt_make(ozt_hi(q) + ((n >> 6) << 20), ozt_lo(q));
    }
#pragma unroll
    for (int i = 0; i < W; ++i) u[i] = 1.0 + t[i];
#pragma unroll
    for (int i = 0; i < W; ++i) y[i] = ozt_rcp_seed(u[i]);
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double e = ozt_fma(-u[i], y[i], 1.0);
        y[i] = ozt_fma(y[i], ozt_fma(e, e, e), y[i]);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        uint32_t idx = (uint32_t)(ozt_hi(u[i]) - 0x3ff00000) >> 12;
        idx = idx > 256u ? 256u : idx;
        double rc, L;
        ozt_log_tab((int)idx, rc, L);
        const double w = ozt_fma(u[i], rc, -1.0), w2 = w * w;
        double p = ozt_fma(w, 1.4285714285714285e-01, -1.6666666666666666e-01);
        p = ozt_fma(p, w, 0.2);
        p = ozt_fma(p, w, -0.25);
        p = ozt_fma(p, w, 3.3333333333333331e-01);
        p = ozt_fma(p, w, -0.5);
        r[i] = L + ozt_fma(w2, p, w);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        sp[i] = ozt_fma(0.5, z[i] - a[i], r[i]);          // max(z, 0) + log(1 + t)
        sg[i] = ozt_hi(z[i]) < 0 ? t[i] * y[i] : y[i];
    }
}

}  // namespace opl
