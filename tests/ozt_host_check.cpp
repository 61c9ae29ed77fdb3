// Host build of csrc/oz_transcend.cuh for tests/test_host_logic.py (test infrastructure: the product compiles the same
// header for the device).  g++ -O2 -ffp-contract=off -shared -fPIC
#include "oz_transcend.cuh"

extern "C" void ozt_softplus_logistic(const double *z, long n, double *sp, double *sg) {
    long i = 0;
    for (; i + 8 <= n; i += 8) {
        double zz[8], a[8], b[8];
        for (int j = 0; j < 8; ++j) zz[j] = z[i + j];
        opl::ozt_softplus_logistic_n<8>(zz, a, b);
        for (int j = 0; j < 8; ++j) { sp[i + j] = a[j]; sg[i + j] = b[j]; }
    }
    for (; i < n; ++i) {
        double zz[1] = {z[i]}, a[1], b[1];
        opl::ozt_softplus_logistic_n<1>(zz, a, b);
        sp[i] = a[0];
        sg[i] = b[0];
    }
}

extern "C" void ozt_scalar_forms(const double *x, long n, double *e, double *l, double *r) {
    for (long i = 0; i < n; ++i) {
        e[i] = opl::ozt_exp_nonpos(-(x[i] < 0 ? -x[i] : x[i]));      // exp(-|x|)
        l[i] = opl::ozt_log_pos(x[i] < 0 ? -x[i] : x[i]);            // log |x|
        r[i] = opl::ozt_rcp(x[i]);
    }
}

// The form the GEMM epilogue runs: groups of 8 values, the regular-argument version when every value of the group is
// finite with |z| < 708, else the full version (csrc/ozaki.cu finish phase).
extern "C" void ozt_softplus_logistic_dispatch(const double *z, long n, double *sp, double *sg, long *regular_groups) {
    long regular_count = 0;
    for (long i = 0; i < n; i += 8) {
        double in[8], a[8], b[8];
        for (int j = 0; j < 8; ++j) in[j] = i + j < n ? z[i + j] : 0.0;
        bool regular = true;
        for (int j = 0; j < 8; ++j) regular = regular && opl::ozt_is_regular(in[j]);
        if (regular) {
            opl::ozt_softplus_logistic_regular_n<8>(in, a, b);
            ++regular_count;
        } else {
            opl::ozt_softplus_logistic_n<8>(in, a, b);
        }
        for (int j = 0; j < 8 && i + j < n; ++j) {
            sp[i + j] = a[j];
            sg[i + j] = b[j];
        }
    }
    if (regular_groups) *regular_groups = regular_count;
}
