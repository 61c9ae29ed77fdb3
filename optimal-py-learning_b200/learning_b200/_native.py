"""ctypes binding of the C-ABI library ``libopl_b200.so`` (``include/opl_b200.h``).

This is the ONLY compute backend of the package: there is no NumPy / CPU fallback.  If the
shared library is missing or older than its sources it is (re)built in-tree with nvcc; if that is
impossible the import fails loudly, and without a CUDA device every compute entry point raises
``RuntimeError``.
"""
import ctypes
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libopl_b200.so")

OPL_OK, OPL_E_ARG, OPL_E_CUDA, OPL_E_STATE, OPL_E_DOMAIN = 0, -1, -2, -3, -4

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p
i32, i64, f64, u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_uint64

# name -> (restype, argtypes); one entry per declaration in include/opl_b200.h
SIGNATURES = {
    "opl_last_error": (ctypes.c_char_p, []),
    "opl_version": (ctypes.c_int, []),
    "opl_device_count": (ctypes.c_int, []),
    "opl_mlp_create": (ctypes.c_int, [ctypes.POINTER(vp), i32, c_int64_p, c_int32_p, c_double_p, i32, i64, i32, vp]),
    "opl_mlp_destroy": (ctypes.c_int, [vp]),
    "opl_mlp_param_count": (i64, [vp]),
    "opl_mlp_set_engine": (ctypes.c_int, [vp, i32]),
    "opl_mlp_set_stream": (ctypes.c_int, [vp, vp]),
    "opl_mlp_set_data_host": (ctypes.c_int, [vp, vp, vp, i64, i64]),
    "opl_mlp_set_data_device": (ctypes.c_int, [vp, vp, vp, i64, i64]),
    "opl_mlp_obj_host": (ctypes.c_int, [vp, vp, c_double_p]),
    "opl_mlp_obj_jac_host": (ctypes.c_int, [vp, vp, c_double_p, vp]),
    "opl_mlp_activate_host": (ctypes.c_int, [vp, vp, vp, i64, vp]),
    "opl_mlp_eval_device": (ctypes.c_int, [vp, vp, f64, vp, i32, vp]),
    "opl_mlp_probe_scalars": (ctypes.c_int, [vp, vp, vp, i32, c_double_p]),
    "opl_mlp_probe_device": (ctypes.c_int, [vp, vp, f64, vp, i32, vp, c_double_p]),
    "opl_mlp_set_small_path": (ctypes.c_int, [vp, i32]),
    "opl_mlp_activate_device": (ctypes.c_int, [vp, vp, vp, i64, vp]),
    "opl_mlp_launch_count": (i64, [vp]),
    "opl_mlp_set_masks_host": (ctypes.c_int, [vp, vp, i64]),
    "opl_mlp_profile": (ctypes.c_int, [vp, i32]),
    "opl_mlp_profile_read": (ctypes.c_int, [vp, i32, c_int32_p, ctypes.POINTER(ctypes.c_float), c_int32_p]),
    "opl_bench_gemm_f64": (ctypes.c_int, [i32, i32, i32, i32, vp, vp, vp, vp, i64, i32, ctypes.POINTER(ctypes.c_float)]),
    "opl_bench_gemm_tf32": (ctypes.c_int, [i32, i32, i32, i32, vp, vp, vp, vp, i64, i32, i32, ctypes.POINTER(ctypes.c_float)]),
    "opl_bench_gemm_ozaki": (ctypes.c_int, [i32, i32, i32, vp, vp, vp, i32, i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "opl_transfer_apply_host": (ctypes.c_int, [i32, f64, vp, i64, i64, vp]),
    "opl_transfer_backprop_host": (ctypes.c_int, [i32, f64, vp, vp, i64, i64, vp]),
    "opl_error_host": (ctypes.c_int, [i32, vp, vp, i64, i64, c_double_p, vp]),
    "opl_softmax_jacobian_host": (ctypes.c_int, [vp, i64, i64, vp]),
    "opl_bfgs_create": (ctypes.c_int, [ctypes.POINTER(vp), i64, i64, i64, vp]),
    "opl_bfgs_destroy": (ctypes.c_int, [vp]),
    "opl_bfgs_reset": (ctypes.c_int, [vp]),
    "opl_bfgs_set_stream": (ctypes.c_int, [vp, vp]),
    "opl_bfgs_direction_device": (ctypes.c_int, [vp, vp, vp, vp, i32, vp]),
    "opl_bfgs_last_slope": (ctypes.c_int, [vp, c_double_p]),
    "opl_bfgs_pass_device": (ctypes.c_int, [vp, vp, vp, vp, i32, vp, vp, vp]),
    "opl_bfgs_finish_device": (ctypes.c_int, [vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "opl_bfgs_materialize_device": (ctypes.c_int, [vp, vp]),
    "opl_bfgs_load_device": (ctypes.c_int, [vp, vp]),
    "opl_bfgs_fill_synthetic": (ctypes.c_int, [vp, u64]),
    "opl_bfgs_launch_count": (i64, [vp]),
    "opl_bfgs_state": (ctypes.c_int, [vp]),
    "opl_lbfgs_create": (ctypes.c_int, [ctypes.POINTER(vp), i64, i32, vp]),
    "opl_lbfgs_destroy": (ctypes.c_int, [vp]),
    "opl_lbfgs_reset": (ctypes.c_int, [vp]),
    "opl_lbfgs_direction_device": (ctypes.c_int, [vp, vp, vp, vp, i32, vp]),
    "opl_lbfgs_history_len": (i32, [vp]),
    "opl_lbfgs_set_stream": (ctypes.c_int, [vp, vp]),
    "opl_lbfgs_push_pair_device": (ctypes.c_int, [vp, vp, vp]),
    "opl_lbfgs_get_pair_device": (ctypes.c_int, [vp, i32, vp, vp]),
    "opl_lbfgs_launch_count": (i64, [vp]),
    "opl_vec_axpy_device": (ctypes.c_int, [i64, vp, f64, vp, f64, vp, vp, vp]),
    "opl_vec_step_device": (ctypes.c_int, [i64, vp, f64, vp, vp, vp, vp]),
    "opl_vec_scale_device": (ctypes.c_int, [i64, f64, vp, vp, vp]),
    "opl_vec_penalty_device": (ctypes.c_int, [i64, vp, i32, f64, vp, c_double_p, vp]),
    "opl_vec_dot_device": (ctypes.c_int, [i64, vp, vp, c_double_p, vp]),
    "opl_gather_rows_device": (ctypes.c_int, [vp, i64, i64, vp, i64, vp, vp]),
    "opl_rbf_similarity_device": (ctypes.c_int, [vp, i64, i32, vp, i32, i32, f64, vp, vp]),
    "opl_som_train_device": (ctypes.c_int, [vp, i64, i32, vp, i32, i64, vp, i32, vp]),
    "opl_allreduce_sum_device": (ctypes.c_int, [i32, i32, ctypes.POINTER(u64), ctypes.POINTER(u64), i64, ctypes.c_uint32, vp, vp]),
    "opl_allreduce_sum_nvls_device": (ctypes.c_int, [i32, i32, u64, u64, vp, ctypes.POINTER(u64), i64, ctypes.c_uint32, vp, vp, vp]),
}


def _load():
    # build.build() compares a digest of csrc/ + the header with the stamp of the last build and recompiles when they
    # differ, so a stale binary is never loaded silently after a source edit (a matching stamp costs a few ms)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_opl_b200_build", os.path.join(_PKG_ROOT, "build.py"))
    try:
        builder = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(builder)
        builder.build()
    except Exception as exc:  # no nvcc / compile error: fail loudly, never fall back
        if not os.path.exists(LIB_PATH) or os.environ.get("OPL_ALLOW_STALE_LIB") != "1":
            raise ImportError(
                "learning_b200: CUDA library %s is missing or out of date and could not be built (%s). "
                "There is no CPU fallback." % (LIB_PATH, exc))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def last_error():
    msg = lib.opl_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Map a C-ABI status to the exception the reference raises at the same point."""
    if rc == OPL_OK:
        return
    msg = last_error()
    if rc == OPL_E_ARG:
        raise ValueError(msg)
    if rc == OPL_E_DOMAIN:
        raise FloatingPointError(msg)   # error.py:89-91 under numpy.errstate(invalid='raise')
    raise RuntimeError(msg)


def device_count():
    return int(lib.opl_device_count())


def require_device():
    if device_count() < 1:
        raise RuntimeError("learning_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
