//! `cuda_backend`: the ristretto255 hot path of zkp's toolbox on NVIDIA B200 GPUs, through the C ABI of
//! `libzkp_b200.so` (`include/zkp_b200.h`).
//!
//! This module is what the nine call sites of `toolbox::{prover, verifier, batch_verifier}` call instead of
//! `RistrettoPoint::{multiscalar_mul, vartime_multiscalar_mul, optional_multiscalar_mul}`, `decompress()`,
//! `compress()` and `is_identity()` when the crate is built with `--features cuda_backend` (see
//! `zkp-cuda-backend.patch`).  Everything that leaves or enters the library is a canonical 32-byte encoding
//! (scalars little-endian and reduced, points as `CompressedRistretto`), so results are bit-identical to the
//! CPU backends; `define_proof!` and the public `toolbox` signatures are untouched.
//!
//! One device context per host thread (contexts are not shared between threads; calls on one context are
//! serialised by the library).  The library never unwinds or aborts across the FFI: every entry point returns
//! a status, mapped here to `None` (= the `None` of `optional_multiscalar_mul`, i.e. an undecodable point) or a panic
//! for conditions that have no `ProofError` (no GPU, out of device memory): like a failed allocation in the CPU
//! backends, they are not recoverable inside a proof.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_void};
use std::ptr;

use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;

#[repr(C)]
pub struct zkp_ctx {
    _private: [u8; 0],
}

pub const ZKP_OK: i32 = 0;
/// an undecodable point encoding: the reference's `None` / `ProofError::VerificationFailure`
pub const ZKP_ERR_POINT: i32 = 1;
pub const ZKP_ERR_SIZE: i32 = 2;
/// a scalar that is not canonical (cannot come out of `Scalar::as_bytes`)
pub const ZKP_ERR_SCALAR: i32 = 3;
const ZKP_POINTS_COMPRESSED: i32 = 0;

#[link(name = "zkp_b200")]
extern "C" {
    fn zkp_ctx_create(out: *mut *mut zkp_ctx, device: i32) -> i32;
    fn zkp_ctx_destroy(ctx: *mut zkp_ctx);
    fn zkp_last_error(ctx: *mut zkp_ctx) -> *const c_char;
    fn zkp_decompress_batch(ctx: *mut zkp_ctx, enc: *const u8, n: usize, limbs_out: *mut u64, valid_out: *mut u8) -> i32;
    fn zkp_msm_vartime(
        ctx: *mut zkp_ctx,
        scalars: *const u8,
        points: *const u8,
        n: usize,
        out32: *mut u8,
        is_identity: *mut i32,
        first_bad: *mut i64,
    ) -> i32;
    fn zkp_msm_vartime_batched(
        ctx: *mut zkp_ctx,
        scalars: *const u8,
        points: *const u8,
        offsets: *const u64,
        m: usize,
        out: *mut u8,
        valid: *mut u8,
    ) -> i32;
    fn zkp_msm_ct_batched(
        ctx: *mut zkp_ctx,
        scalars: *const u8,
        points: *const c_void,
        point_format: i32,
        offsets: *const u64,
        m: usize,
        out: *mut u8,
    ) -> i32;
    fn zkp_batch_verify(
        ctx: *mut zkp_ctx,
        static_coeffs: *const u8,
        static_points: *const u8,
        num_s: usize,
        instance_coeffs: *const u8,
        instance_points: *const u8,
        rows: usize,
        batch: usize,
        accept: *mut i32,
        first_bad: *mut i64,
    ) -> i32;
}

/// Owner of one `zkp_ctx` (device `ZKP_B200_DEVICE`, default 0).
struct Context(*mut zkp_ctx);

impl Context {
    fn new() -> Context {
        let device = std::env::var("ZKP_B200_DEVICE")
            .ok()
            .and_then(|s| s.parse::<i32>().ok())
            .unwrap_or(0);
        let mut raw: *mut zkp_ctx = ptr::null_mut();
        let rc = unsafe { zkp_ctx_create(&mut raw, device) };
        if rc != ZKP_OK || raw.is_null() {
            panic!("zkp cuda_backend: zkp_ctx_create(device {}) failed with status {}", device, rc);
        }
        Context(raw)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { zkp_ctx_destroy(self.0) }
    }
}

thread_local! {
    static CONTEXT: Context = Context::new();
}

fn with_ctx<R>(f: impl FnOnce(*mut zkp_ctx) -> R) -> R {
    CONTEXT.with(|c| f(c.0))
}

fn engine_failure(ctx: *mut zkp_ctx, what: &str, rc: i32) -> ! {
    let msg = unsafe {
        let p = zkp_last_error(ctx);
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    };
    panic!("zkp cuda_backend: {} failed with status {}: {}", what, rc, msg);
}

fn scalar_bytes<'a, I: IntoIterator<Item = &'a Scalar>>(scalars: I) -> Vec<u8> {
    let mut out = Vec::new();
    for s in scalars {
        out.extend_from_slice(s.as_bytes());
    }
    out
}

fn point_bytes<'a, I: IntoIterator<Item = &'a CompressedRistretto>>(points: I) -> Vec<u8> {
    let mut out = Vec::new();
    for p in points {
        out.extend_from_slice(p.as_bytes());
    }
    out
}

fn encodings(bytes: &[u8]) -> Vec<CompressedRistretto> {
    bytes.chunks(32).map(CompressedRistretto::from_slice).collect()
}

/// Replaces the `decompress()` of every public point at `verifier.rs:87-92`: `true` iff every encoding is a
/// valid ristretto255 point (the decompressed points themselves stay on the device side of the later calls).
pub fn all_points_decompress(points: &[CompressedRistretto]) -> bool {
    if points.is_empty() {
        return true;
    }
    let enc = point_bytes(points.iter());
    let mut limbs = vec![0u64; 20 * points.len()];
    let mut valid = vec![0u8; points.len()];
    with_ctx(|ctx| {
        let rc = unsafe { zkp_decompress_batch(ctx, enc.as_ptr(), points.len(), limbs.as_mut_ptr(), valid.as_mut_ptr()) };
        if rc != ZKP_OK {
            engine_failure(ctx, "zkp_decompress_batch", rc);
        }
    });
    valid.iter().all(|&v| v == 1)
}

/// Replaces the per-constraint `RistrettoPoint::multiscalar_mul` + `compress()` of `prover.rs:93-103`
/// (constant time: no branch or address in the kernels depends on a scalar).  MSM `j` covers the terms
/// `offsets[j] .. offsets[j + 1]` of `scalars` / `points`; returns one encoding per MSM.
pub fn multiscalar_mul_compressed(
    scalars: &[Scalar],
    points: &[CompressedRistretto],
    offsets: &[u64],
) -> Vec<CompressedRistretto> {
    assert_eq!(scalars.len(), points.len());
    let m = offsets.len() - 1;
    assert_eq!(offsets[m] as usize, scalars.len());
    let sc = scalar_bytes(scalars.iter());
    let pt = point_bytes(points.iter());
    let mut out = vec![0u8; 32 * m];
    with_ctx(|ctx| {
        let rc = unsafe {
            zkp_msm_ct_batched(
                ctx,
                sc.as_ptr(),
                pt.as_ptr() as *const c_void,
                ZKP_POINTS_COMPRESSED,
                offsets.as_ptr(),
                m,
                out.as_mut_ptr(),
            )
        };
        // the prover's points were compressed by this crate a moment ago: an undecodable one is a bug, not an input error
        if rc != ZKP_OK {
            engine_failure(ctx, "zkp_msm_ct_batched", rc);
        }
    });
    encodings(&out)
}

/// Replaces the per-constraint `RistrettoPoint::vartime_multiscalar_mul` + `compress()` of
/// `verifier.rs:97-109`.  `None` if a point of any MSM does not decompress.
pub fn vartime_multiscalar_mul_compressed(
    scalars: &[Scalar],
    points: &[CompressedRistretto],
    offsets: &[u64],
) -> Option<Vec<CompressedRistretto>> {
    assert_eq!(scalars.len(), points.len());
    let m = offsets.len() - 1;
    assert_eq!(offsets[m] as usize, scalars.len());
    let sc = scalar_bytes(scalars.iter());
    let pt = point_bytes(points.iter());
    let mut out = vec![0u8; 32 * m];
    let mut valid = vec![0u8; m];
    with_ctx(|ctx| {
        let rc = unsafe {
            zkp_msm_vartime_batched(ctx, sc.as_ptr(), pt.as_ptr(), offsets.as_ptr(), m, out.as_mut_ptr(), valid.as_mut_ptr())
        };
        if rc != ZKP_OK {
            engine_failure(ctx, "zkp_msm_vartime_batched", rc);
        }
    });
    if valid.iter().all(|&v| v == 1) {
        Some(encodings(&out))
    } else {
        None
    }
}

/// Replaces `RistrettoPoint::optional_multiscalar_mul(coeffs, points.map(decompress))` followed by
/// `is_identity()` at `verifier.rs:162-172`: `None` if a point does not decompress, else whether the sum is the
/// identity of ristretto255 (coset-aware, like `RistrettoPoint::is_identity`).
pub fn optional_multiscalar_mul_is_identity<'a, I, J>(scalars: I, points: J) -> Option<bool>
where
    I: IntoIterator<Item = &'a Scalar>,
    J: IntoIterator<Item = &'a CompressedRistretto>,
{
    let sc = scalar_bytes(scalars);
    let pt = point_bytes(points);
    assert_eq!(sc.len(), pt.len());
    let n = sc.len() / 32;
    let mut enc = [0u8; 32];
    let mut is_identity: i32 = 0;
    let mut first_bad: i64 = -1;
    with_ctx(|ctx| {
        let rc = unsafe {
            zkp_msm_vartime(ctx, sc.as_ptr(), pt.as_ptr(), n, enc.as_mut_ptr(), &mut is_identity, &mut first_bad)
        };
        match rc {
            ZKP_OK => Some(is_identity == 1),
            ZKP_ERR_POINT => None,
            _ => engine_failure(ctx, "zkp_msm_vartime", rc),
        }
    })
}

/// Replaces the one large MSM + identity test of `batch_verifier.rs:219-234`.  `instance_coeffs` is the
/// row-major `(rows x batch)` coefficient matrix (`Matrix::row_major_entries`), `instance_points` the
/// matching flat point list (instance rows, then commitment rows).  `None` if a point does not decompress.
pub fn batch_verify<'a, I>(
    static_coeffs: &[Scalar],
    static_points: &[CompressedRistretto],
    instance_coeffs: I,
    instance_points: &[CompressedRistretto],
    rows: usize,
    batch: usize,
) -> Option<bool>
where
    I: IntoIterator<Item = &'a Scalar>,
{
    assert_eq!(static_coeffs.len(), static_points.len());
    assert_eq!(instance_points.len(), rows * batch);
    let ssc = scalar_bytes(static_coeffs.iter());
    let spt = point_bytes(static_points.iter());
    let isc = scalar_bytes(instance_coeffs);
    let ipt = point_bytes(instance_points.iter());
    assert_eq!(isc.len(), ipt.len());
    let mut accept: i32 = 0;
    let mut first_bad: i64 = -1;
    with_ctx(|ctx| {
        let rc = unsafe {
            zkp_batch_verify(
                ctx,
                ssc.as_ptr(),
                spt.as_ptr(),
                static_points.len(),
                isc.as_ptr(),
                ipt.as_ptr(),
                rows,
                batch,
                &mut accept,
                &mut first_bad,
            )
        };
        match rc {
            ZKP_OK => Some(accept == 1),
            ZKP_ERR_POINT => None,
            _ => engine_failure(ctx, "zkp_batch_verify", rc),
        }
    })
}
