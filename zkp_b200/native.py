"""ctypes binding of libzkp_b200.so (include/zkp_b200.h).  There is no CPU fallback: if the CUDA library is
missing or no GPU is present the product path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzkp_b200.so")

ZKP_OK, ZKP_ERR_POINT, ZKP_ERR_SIZE, ZKP_ERR_SCALAR = 0, 1, 2, 3
ZKP_ERR_NOGPU, ZKP_ERR_CUDA, ZKP_ERR_NOMEM = -1, -2, -3
ZKP_POINTS_COMPRESSED, ZKP_POINTS_LIMBS51 = 0, 1

# every symbol include/zkp_b200.h declares
SYMBOLS = [
    "zkp_device_count", "zkp_ctx_create", "zkp_ctx_destroy", "zkp_ctx_set_stream", "zkp_ctx_set_option",
    "zkp_ctx_synchronize", "zkp_last_error", "zkp_ctx_launch_count", "zkp_decompress_batch", "zkp_compress_batch",
    "zkp_msm_vartime", "zkp_msm_vartime_dev", "zkp_msm_vartime_batched", "zkp_msm_ct_batched", "zkp_batch_verify",
    "zkp_bench_field", "zkp_ctx_stage_ms", "zkp_bench_dual", "zkp_batch_verify_proofs", "zkp_selftest_hash", "zkp_prove_batch", "zkp_batch_verify_partial", "zkp_partials_verdict", "zkp_selftest_bv_script", "zkp_selftest_chunk_schedule",
    "zkp_msm_vartime_partial_dev", "zkp_partials_verdict_dev",
]

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load the in-tree CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: run `python -m zkp_b200.build` (or __graft_entry__.build()). "
            "zkp_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_void_p, c_int32, c_int64, c_size_t, c_char_p = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                                                      ctypes.c_size_t, ctypes.c_char_p)
    P = ctypes.POINTER
    lib.zkp_device_count.restype = c_int32
    lib.zkp_ctx_create.argtypes = [P(c_void_p), c_int32]
    lib.zkp_ctx_create.restype = c_int32
    lib.zkp_ctx_destroy.argtypes = [c_void_p]
    lib.zkp_ctx_destroy.restype = None
    lib.zkp_ctx_set_stream.argtypes = [c_void_p, c_void_p]
    lib.zkp_ctx_set_stream.restype = c_int32
    lib.zkp_ctx_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.zkp_ctx_set_option.restype = c_int32
    lib.zkp_ctx_synchronize.argtypes = [c_void_p]
    lib.zkp_ctx_synchronize.restype = c_int32
    lib.zkp_last_error.argtypes = [c_void_p]
    lib.zkp_last_error.restype = c_char_p
    lib.zkp_ctx_launch_count.argtypes = [c_void_p]
    lib.zkp_ctx_launch_count.restype = ctypes.c_uint64
    lib.zkp_decompress_batch.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.zkp_decompress_batch.restype = c_int32
    lib.zkp_compress_batch.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.zkp_compress_batch.restype = c_int32
    lib.zkp_msm_vartime.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, P(c_int32), P(c_int64)]
    lib.zkp_msm_vartime.restype = c_int32
    lib.zkp_msm_vartime_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.zkp_msm_vartime_dev.restype = c_int32
    lib.zkp_msm_vartime_batched.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.zkp_msm_vartime_batched.restype = c_int32
    lib.zkp_msm_ct_batched.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]
    lib.zkp_msm_ct_batched.restype = c_int32
    lib.zkp_batch_verify.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t,
                                     P(c_int32), P(c_int64)]
    lib.zkp_batch_verify.restype = c_int32
    lib.zkp_batch_verify_partial.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t,
                                             c_void_p, P(c_int64)]
    lib.zkp_batch_verify_partial.restype = c_int32
    lib.zkp_partials_verdict.argtypes = [c_void_p, c_void_p, c_size_t, P(c_int32), c_void_p]
    lib.zkp_partials_verdict.restype = c_int32
    lib.zkp_msm_vartime_partial_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.zkp_msm_vartime_partial_dev.restype = c_int32
    lib.zkp_partials_verdict_dev.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.zkp_partials_verdict_dev.restype = c_int32
    lib.zkp_bench_field.argtypes = [c_void_p, c_int32, c_int32, P(ctypes.c_double)]
    lib.zkp_bench_field.restype = c_int32
    lib.zkp_bench_dual.argtypes = [c_void_p, c_int32, c_int32, P(ctypes.c_double)]
    lib.zkp_bench_dual.restype = c_int32
    lib.zkp_ctx_stage_ms.argtypes = [c_void_p, c_int32]
    lib.zkp_ctx_stage_ms.restype = ctypes.c_double
    _lib = lib
    return lib
