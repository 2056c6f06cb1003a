"""Python face of the C ABI (include/zkp_b200.h): thin, numpy in / numpy out, no arithmetic here.

`Engine` owns one `zkp_ctx` (one device, one stream).  Method names follow the reference's entry points:
  msm_vartime           RistrettoPoint::optional_multiscalar_mul / vartime_multiscalar_mul
                        (/root/reference/src/toolbox/batch_verifier.rs:219, verifier.rs:97,162)
  msm_ct_batched        RistrettoPoint::multiscalar_mul + compress (/root/reference/src/toolbox/prover.rs:93-103)
  decompress_batch      CompressedRistretto::decompress (verifier.rs:87-92)
  compress_batch        RistrettoPoint::compress (toolbox/mod.rs:180,204)
  batch_verify          BatchVerifier::verify_batchable from the MSM on (batch_verifier.rs:208-234)
"""
import ctypes

import numpy as np

from . import native


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zkp_b200 error %d: %s" % (code, msg))
        self.code = code


def _u8(a, width):
    """Accept bytes / list of bytes / ndarray; return C-contiguous uint8 array of shape (n, width)."""
    if isinstance(a, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(bytes(a), dtype=np.uint8)
    elif isinstance(a, np.ndarray):
        arr = a
    else:
        arr = np.frombuffer(b"".join(bytes(x) for x in a), dtype=np.uint8)
    arr = np.ascontiguousarray(arr, dtype=np.uint8).reshape(-1, width)
    return arr


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None and a.size else None


class Engine:
    def __init__(self, device=0):
        self._lib = native.load()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.zkp_ctx_create(ctypes.byref(self._ctx), int(device))
        if rc != native.ZKP_OK:
            self._ctx = None
            raise EngineError(rc, "zkp_ctx_create failed (no CUDA device?) -- zkp_b200 has no CPU fallback")
        self.device = int(device)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.zkp_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, ok=(native.ZKP_OK,)):
        if rc not in ok:
            raise EngineError(rc, self._lib.zkp_last_error(self._ctx).decode(errors="replace"))
        return rc

    # ---- configuration -----------------------------------------------------------------------------------
    def set_option(self, key, value):
        self._check(self._lib.zkp_ctx_set_option(self._ctx, key.encode(), int(value)))

    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.zkp_ctx_set_stream(self._ctx, ctypes.c_void_p(cuda_stream_handle or 0)))

    def synchronize(self):
        self._check(self._lib.zkp_ctx_synchronize(self._ctx))

    STAGES = ["decompress", "recode_hist", "scan", "scatter", "accumulate", "bucket_reduce", "finish"]

    def stage_ms(self):
        """Per-stage device times of the last vartime MSM (needs set_option('profile', 1))."""
        d = {n: float(self._lib.zkp_ctx_stage_ms(self._ctx, i)) for i, n in enumerate(self.STAGES)}
        d["window"] = int(self._lib.zkp_ctx_stage_ms(self._ctx, 100))
        d["lanes"] = int(self._lib.zkp_ctx_stage_ms(self._ctx, 101))
        return d

    def live_ms(self):
        """Device times of the dominant kernels of the LAST device-resident MSM, from events that are always recorded
        (no profiling mode): the two ingestion launches and the bucket accumulation.  Call after synchronising."""
        names = ("ingest_phase1", "ingest_phase2", "accumulate")
        return {n: float(self._lib.zkp_ctx_stage_ms(self._ctx, 7 + i)) for i, n in enumerate(names)}

    @property
    def launch_count(self):
        return int(self._lib.zkp_ctx_launch_count(self._ctx))

    # ---- the hot path ------------------------------------------------------------------------------------
    def msm_vartime(self, scalars, points):
        """Returns (encoding: bytes | None, is_identity: bool, first_bad: int).  None == the reference's `None`."""
        s, p = _u8(scalars, 32), _u8(points, 32)
        if s.shape[0] != p.shape[0]:
            raise EngineError(native.ZKP_ERR_SIZE, "scalars/points length mismatch")
        out = np.zeros(32, dtype=np.uint8)
        ident, bad = ctypes.c_int32(0), ctypes.c_int64(-1)
        rc = self._lib.zkp_msm_vartime(self._ctx, _ptr(s), _ptr(p), s.shape[0], _ptr(out), ctypes.byref(ident),
                                       ctypes.byref(bad))
        self._check(rc, ok=(native.ZKP_OK, native.ZKP_ERR_POINT))
        if rc == native.ZKP_ERR_POINT:
            return None, False, int(bad.value)
        return out.tobytes(), bool(ident.value), -1

    def msm_vartime_dev(self, d_scalars_ptr, d_points_ptr, n, d_result_ptr):
        """Asynchronous, device pointers (ints).  See zkp_msm_vartime_dev for the 48-byte result layout."""
        self._check(self._lib.zkp_msm_vartime_dev(self._ctx, ctypes.c_void_p(d_scalars_ptr),
                                                  ctypes.c_void_p(d_points_ptr), int(n),
                                                  ctypes.c_void_p(d_result_ptr)))

    def msm_vartime_partial_dev(self, d_scalars_ptr, d_points_ptr, n, d_result_ptr, d_partial_ptr):
        """msm_vartime_dev that also leaves the sum as a 160-byte limb-form point at d_partial_ptr (single-verdict mode)."""
        self._check(self._lib.zkp_msm_vartime_partial_dev(self._ctx, ctypes.c_void_p(d_scalars_ptr),
                                                          ctypes.c_void_p(d_points_ptr), int(n),
                                                          ctypes.c_void_p(d_result_ptr), ctypes.c_void_p(d_partial_ptr)))

    def partials_verdict_dev(self, d_partials_ptr, count, d_result_ptr):
        """Sum of `count` 160-byte limb-form points on the device -> the 48-byte result record (asynchronous)."""
        self._check(self._lib.zkp_partials_verdict_dev(self._ctx, ctypes.c_void_p(d_partials_ptr), int(count),
                                                       ctypes.c_void_p(d_result_ptr)))

    def batch_verify(self, static_coeffs, static_points, instance_coeffs, instance_points, rows, batch):
        """Returns (accept: bool, status).  status ZKP_ERR_POINT == VerificationFailure from a bad encoding."""
        sc, sp = _u8(static_coeffs, 32), _u8(static_points, 32)
        ic, ip = _u8(instance_coeffs, 32), _u8(instance_points, 32)
        if sc.shape[0] != sp.shape[0] or ic.shape[0] != rows * batch or ip.shape[0] != rows * batch:
            raise EngineError(native.ZKP_ERR_SIZE, "batch_verify: inconsistent sizes")
        acc, bad = ctypes.c_int32(0), ctypes.c_int64(-1)
        rc = self._lib.zkp_batch_verify(self._ctx, _ptr(sc), _ptr(sp), sc.shape[0], _ptr(ic), _ptr(ip), int(rows),
                                        int(batch), ctypes.byref(acc), ctypes.byref(bad))
        self._check(rc, ok=(native.ZKP_OK, native.ZKP_ERR_POINT))
        return bool(acc.value) and rc == native.ZKP_OK, rc

    def batch_verify_partial(self, static_coeffs, static_points, instance_coeffs, instance_points, rows, batch):
        """One shard of a batch in single-verdict mode: returns the shard's MSM sum as a (4, 5) uint64 limb-form point
        (X, Y, Z, T), or None for an invalid point encoding (VerificationFailure)."""
        sc, sp = _u8(static_coeffs, 32), _u8(static_points, 32)
        ic, ip = _u8(instance_coeffs, 32), _u8(instance_points, 32)
        if sc.shape[0] != sp.shape[0] or ic.shape[0] != rows * batch or ip.shape[0] != rows * batch:
            raise EngineError(native.ZKP_ERR_SIZE, "batch_verify_partial: inconsistent sizes")
        out = np.zeros((4, 5), dtype=np.uint64)
        bad = ctypes.c_int64(-1)
        rc = self._lib.zkp_batch_verify_partial(self._ctx, _ptr(sc), _ptr(sp), sc.shape[0], _ptr(ic), _ptr(ip), int(rows),
                                                int(batch), _ptr(out), ctypes.byref(bad))
        self._check(rc, ok=(native.ZKP_OK, native.ZKP_ERR_POINT))
        return out if rc == native.ZKP_OK else None

    def partials_verdict(self, partials):
        """Identity test (batch_verifier.rs:230) on the sum of the shards' partial sums: (accept, encoding of the sum)."""
        p = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, 20)
        acc = ctypes.c_int32(0)
        enc = np.zeros(32, dtype=np.uint8)
        self._check(self._lib.zkp_partials_verdict(self._ctx, _ptr(p), p.shape[0], ctypes.byref(acc), _ptr(enc)))
        return bool(acc.value), enc.tobytes()

    def decompress_batch(self, encodings):
        e = _u8(encodings, 32)
        n = e.shape[0]
        limbs = np.zeros((n, 4, 5), dtype=np.uint64)
        valid = np.zeros(n, dtype=np.uint8)
        self._check(self._lib.zkp_decompress_batch(self._ctx, _ptr(e), n, _ptr(limbs), _ptr(valid)))
        return limbs, valid

    def compress_batch(self, limbs):
        l = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4, 5)
        out = np.zeros((l.shape[0], 32), dtype=np.uint8)
        self._check(self._lib.zkp_compress_batch(self._ctx, _ptr(l), l.shape[0], _ptr(out)))
        return out

    def msm_vartime_batched(self, scalars, points, offsets):
        s, p = _u8(scalars, 32), _u8(points, 32)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        M = off.shape[0] - 1
        if M < 0 or s.shape[0] != p.shape[0] or (M >= 0 and int(off[-1]) != s.shape[0]):
            raise EngineError(native.ZKP_ERR_SIZE, "msm_vartime_batched: inconsistent sizes")
        out = np.zeros((max(M, 0), 32), dtype=np.uint8)
        valid = np.zeros(max(M, 0), dtype=np.uint8)
        self._check(self._lib.zkp_msm_vartime_batched(self._ctx, _ptr(s), _ptr(p), _ptr(off), M, _ptr(out),
                                                      _ptr(valid)))
        return out, valid

    def msm_ct_batched(self, scalars, points, offsets, limbs=False):
        s = _u8(scalars, 32)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        M = off.shape[0] - 1
        if limbs:
            p = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 4, 5)
            fmt = native.ZKP_POINTS_LIMBS51
        else:
            p = _u8(points, 32)
            fmt = native.ZKP_POINTS_COMPRESSED
        if M < 0 or s.shape[0] != p.shape[0] or int(off[-1]) != s.shape[0]:
            raise EngineError(native.ZKP_ERR_SIZE, "msm_ct_batched: inconsistent sizes")
        out = np.zeros((M, 32), dtype=np.uint8)
        rc = self._lib.zkp_msm_ct_batched(self._ctx, _ptr(s), _ptr(p), fmt, _ptr(off), M, _ptr(out))
        self._check(rc)
        return out

    def bench_field(self, kind, iters=4096):
        r = ctypes.c_double(0)
        self._check(self._lib.zkp_bench_field(self._ctx, int(kind), int(iters), ctypes.byref(r)))
        return r.value
