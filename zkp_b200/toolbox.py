"""Python face of the C++ host mirror of zkp's toolbox (include/zkp_b200_host.h -> zkp_b200/csrc/host/).

`Statement` is what `define_proof!{name, label, (secrets), (instance), (common) : lhs = (s*P + ...), ...}` expands to
(/root/reference/src/macros.rs:74-370): the same five module functions, same argument meaning, same errors.  All
arithmetic on points happens on the GPU through the C ABI; hashing and scalar arithmetic run in the C++ host
library.  Nothing here falls back to a CPU implementation of the hot path.
"""
import ctypes

import numpy as np

from . import native
from .engine import Engine, _u8


class ProofError(Exception):
    """zkp::ProofError (/root/reference/src/errors.rs:4-11)."""


class VerificationFailure(ProofError):
    pass


class BatchSizeMismatch(ProofError):
    pass


class MalformedProof(ValueError):
    """Input the reference's `bincode::deserialize` would refuse (truncated, trailing bytes, non-canonical scalar)."""


def _raise(rc):
    if rc == 0:
        return
    if rc == 1:
        raise VerificationFailure()
    if rc == 2:
        raise BatchSizeMismatch()
    if rc == 4:
        raise MalformedProof()
    raise RuntimeError("zkp_b200 host/engine failure (code %d)" % rc)


_bound = False


def _lib():
    global _bound
    lib = native.load()
    if not _bound:
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32
        lib.zkph_statement_new.restype = vp
        lib.zkph_statement_new.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, i32, i32, i32, vp, vp, vp, vp]
        lib.zkph_statement_free.argtypes = [vp]
        lib.zkph_prove.argtypes = [vp, vp, vp, sz, vp, vp, vp, sz, i32, vp, vp, vp, vp, vp]
        lib.zkph_verify_compact.argtypes = [vp, vp, vp, sz, vp, vp, vp, sz]
        lib.zkph_verify_batchable.argtypes = [vp, vp, vp, sz, vp, vp, sz, vp, sz, vp, sz]
        lib.zkph_batch_verify.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, sz, i32, vp, vp, vp]
        lib.zkph_batch_verify_t.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, sz, i32]
        lib.zkph_prove_many.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, i32, vp, vp, vp]
        lib.zkph_batch_verify_device.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp]
        lib.zkph_prove_many_device.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp]
        lib.zkph_compact_proof_size.restype = sz
        lib.zkph_compact_proof_size.argtypes = [sz]
        lib.zkph_batchable_proof_size.restype = sz
        lib.zkph_batchable_proof_size.argtypes = [sz, sz]
        lib.zkph_compact_proof_serialize.argtypes = [vp, vp, sz, vp]
        lib.zkph_batchable_proof_serialize.argtypes = [vp, sz, vp, sz, vp]
        lib.zkph_compact_proof_parse.argtypes = [vp, sz, sz, vp, vp, vp, vp]
        lib.zkph_batchable_proofs_parse.argtypes = [vp, sz, sz, sz, sz, vp, vp, i32, vp]
        lib.zkph_transcript_new.restype = vp
        lib.zkph_transcript_new.argtypes = [vp, sz]
        lib.zkph_transcript_clone.restype = vp
        lib.zkph_transcript_clone.argtypes = [vp]
        lib.zkph_transcript_free.argtypes = [vp]
        lib.zkph_transcript_append_message.argtypes = [vp, vp, sz, vp, sz]
        lib.zkph_transcript_challenge_bytes.argtypes = [vp, vp, sz, vp, sz]
        lib.zkph_prove_t.argtypes = [vp, vp, vp, vp, vp, vp, sz, i32, vp, vp, vp, vp, vp]
        lib.zkph_verify_compact_t.argtypes = [vp, vp, vp, vp, vp, vp, sz]
        lib.zkph_verify_batchable_t.argtypes = [vp, vp, vp, vp, vp, sz, vp, sz, vp, sz]
        lib.zkph_scalar_mul.argtypes = [vp, vp, vp]
        lib.zkph_scalar_from_wide.argtypes = [vp, vp]
        lib.zkph_merlin_test_vector.argtypes = [vp]
        lib.zkph_rng_bytes.argtypes = [vp, sz, vp, sz]
        _bound = True
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _seed(seed):
    """(pointer argument, length) of an optional seed: None lets the library draw from the OS CSPRNG (the default a
    deployment wants; see RANDOMNESS in include/zkp_b200_host.h).  A seed that is given must be fresh and secret."""
    if seed is None:
        return None, 0
    seed = bytes(seed)
    return seed, len(seed)


class Transcript:
    """merlin::Transcript (re-exported by the reference as zkp::Transcript, /root/reference/src/lib.rs:35), hashed
    by the C++ host library.  Pass an instance wherever a transcript label (bytes) is accepted to drive the
    transcript yourself, as the reference's `&mut Transcript` arguments allow."""

    def __init__(self, label=None, _handle=None):
        self._h = _handle if _handle is not None else _lib().zkph_transcript_new(bytes(label), len(label))
        if not self._h:
            raise MemoryError("zkph_transcript_new / zkph_transcript_clone returned NULL")

    def clone(self):
        h = _lib().zkph_transcript_clone(self._h)
        if not h:
            raise MemoryError("zkph_transcript_clone returned NULL")
        return Transcript(_handle=h)

    def append_message(self, label, message):
        _lib().zkph_transcript_append_message(self._h, bytes(label), len(label), bytes(message), len(message))

    def challenge_bytes(self, label, n):
        out = ctypes.create_string_buffer(n)
        _lib().zkph_transcript_challenge_bytes(self._h, bytes(label), len(label), out, n)
        return out.raw

    def __del__(self):
        try:
            _lib().zkph_transcript_free(self._h)
        except Exception:
            pass


class Statement:
    """define_proof! mirror.  constraints: [(lhs_name, [(secret_name, point_name), ...]), ...]."""

    def __init__(self, name, label, secrets, instance, common, constraints):
        self.name, self.label = name, label
        self.secrets, self.instance, self.common = list(secrets), list(instance), list(common)
        self.points = self.instance + self.common
        self.constraints = [(l, list(r)) for l, r in constraints]
        lib = _lib()
        labels = b"".join(s.encode() + b"\0" for s in self.secrets + self.instance + self.common)
        lhs = np.array([self.points.index(l) for l, _ in self.constraints], dtype=np.int32)
        off = np.cumsum([0] + [len(r) for _, r in self.constraints]).astype(np.int32)
        ts = np.array([self.secrets.index(s) for _, r in self.constraints for s, _ in r], dtype=np.int32)
        tp = np.array([self.points.index(p) for _, r in self.constraints for _, p in r], dtype=np.int32)
        self._keep = (labels, lhs, off, ts, tp)
        self._h = lib.zkph_statement_new(name.encode(), label.encode(), labels, len(self.secrets), len(self.instance),
                                         len(self.common), len(self.constraints), _p(lhs), _p(off), _p(ts), _p(tp))
        if not self._h:
            raise ValueError("inconsistent statement description (zkph_statement_new refused it)")
        self.m, self.p, self.k = len(self.secrets), len(self.points), len(self.constraints)

    def __del__(self):
        try:
            _lib().zkph_statement_free(self._h)
        except Exception:
            pass

    # ---- module::prove_compact / prove_batchable ------------------------------------------------------------
    def _prove(self, eng, transcript_label, secrets, points_limbs, rng_seed, batchable):
        lib = _lib()
        sec = _u8(secrets, 32)
        pts = np.ascontiguousarray(points_limbs, dtype=np.uint64).reshape(self.p, 20)
        assert sec.shape[0] == self.m
        enc = np.zeros((self.p, 32), np.uint8)
        chal = np.zeros(32, np.uint8)
        com = np.zeros((self.k, 32), np.uint8)
        resp = np.zeros((self.m, 32), np.uint8)
        blind = np.zeros((self.m, 32), np.uint8)
        seed, seed_len = _seed(rng_seed)
        if isinstance(transcript_label, Transcript):
            _raise(lib.zkph_prove_t(eng._ctx, self._h, transcript_label._h, _p(sec), _p(pts), seed, seed_len,
                                    1 if batchable else 0, _p(enc), _p(chal), _p(com), _p(resp), _p(blind)))
        else:
            _raise(lib.zkph_prove(eng._ctx, self._h, transcript_label, len(transcript_label), _p(sec), _p(pts), seed,
                                  seed_len, 1 if batchable else 0, _p(enc), _p(chal), _p(com), _p(resp), _p(blind)))
        return enc, chal, com, resp, blind

    def prove_compact(self, eng, transcript_label, secrets, points_limbs, rng_seed=None):
        """-> ((challenge, responses), encodings)   (macros.rs:261-268)"""
        enc, chal, _, resp, _ = self._prove(eng, transcript_label, secrets, points_limbs, rng_seed, False)
        return (chal.tobytes(), resp), enc

    def prove_batchable(self, eng, transcript_label, secrets, points_limbs, rng_seed=None):
        """-> ((commitments, responses), encodings)   (macros.rs:271-278)"""
        enc, _, com, resp, _ = self._prove(eng, transcript_label, secrets, points_limbs, rng_seed, True)
        return (com, resp), enc

    # ---- module::verify_* -------------------------------------------------------------------------------------
    def verify_compact(self, eng, proof, transcript_label, encodings):
        chal, resp = proof
        resp = _u8(resp, 32)
        enc = _u8(encodings, 32)
        assert enc.shape[0] == self.p
        if isinstance(transcript_label, Transcript):
            _raise(_lib().zkph_verify_compact_t(eng._ctx, self._h, transcript_label._h, _p(enc), bytes(chal), _p(resp),
                                                resp.shape[0]))
        else:
            _raise(_lib().zkph_verify_compact(eng._ctx, self._h, transcript_label, len(transcript_label), _p(enc),
                                              bytes(chal), _p(resp), resp.shape[0]))

    def verify_batchable(self, eng, proof, transcript_label, encodings, rng_seed=None):
        com, resp = proof
        com, resp, enc = _u8(com, 32), _u8(resp, 32), _u8(encodings, 32)
        assert enc.shape[0] == self.p
        seed, seed_len = _seed(rng_seed)
        if isinstance(transcript_label, Transcript):
            _raise(_lib().zkph_verify_batchable_t(eng._ctx, self._h, transcript_label._h, _p(enc), _p(com), com.shape[0],
                                                  _p(resp), resp.shape[0], seed, seed_len))
        else:
            _raise(_lib().zkph_verify_batchable(eng._ctx, self._h, transcript_label, len(transcript_label), _p(enc), _p(com),
                                                com.shape[0], _p(resp), resp.shape[0], seed, seed_len))

    # ---- module::batch_verify ---------------------------------------------------------------------------------
    def batch_verify(self, eng, proofs_commitments, proofs_responses, transcript_label, instance_enc, common_enc,
                     rng_seed=None, threads=0, want_msm_inputs=False):
        """proofs_commitments (N,k,32), proofs_responses (N,m,32), instance_enc (n_instance,N,32), common_enc
        (n_common,32).  Raises VerificationFailure / BatchSizeMismatch like the reference.  With
        want_msm_inputs returns (scalars, points, host_seconds) exactly as fed to the MSM.
        transcript_label: one label (bytes) shared by all proofs, or -- the reference's own signature, macros.rs:336-346 --
        a list of Transcript objects, one per proof, each with its own prior state; they are left advanced, and a list
        whose length differs from the number of proofs is BatchSizeMismatch (batch_verifier.rs:72-74)."""
        com = np.ascontiguousarray(proofs_commitments, dtype=np.uint8)
        resp = np.ascontiguousarray(proofs_responses, dtype=np.uint8)
        inst = np.ascontiguousarray(instance_enc, dtype=np.uint8)
        comm = np.ascontiguousarray(common_enc, dtype=np.uint8)
        N = com.shape[0]
        if resp.shape[0] != N or inst.shape[1] != N:
            raise BatchSizeMismatch()
        assert com.shape[1:] == (self.k, 32) and resp.shape[1:] == (self.m, 32) and inst.shape[0] == len(self.instance)
        n = len(self.common) + (len(self.instance) + self.k) * N
        co = np.zeros((n, 32), np.uint8) if want_msm_inputs else None
        po = np.zeros((n, 32), np.uint8) if want_msm_inputs else None
        hs = ctypes.c_double(0)
        seed, seed_len = _seed(rng_seed)
        if isinstance(transcript_label, (list, tuple)):
            if want_msm_inputs:
                raise ValueError("want_msm_inputs needs the shared-label form")
            handles = (ctypes.c_void_p * len(transcript_label))(*[t._h for t in transcript_label])
            _raise(_lib().zkph_batch_verify_t(eng._ctx, self._h, handles, len(transcript_label), N, _p(inst), _p(comm),
                                              _p(com), _p(resp), seed, seed_len, int(threads)))
            return 0.0
        rc = _lib().zkph_batch_verify(eng._ctx, self._h, transcript_label, len(transcript_label), N, _p(inst), _p(comm),
                                      _p(com), _p(resp), seed, seed_len, int(threads), _p(co), _p(po),
                                      ctypes.byref(hs))
        _raise(rc)
        if want_msm_inputs:
            return co, po, hs.value
        return hs.value

    def batch_verify_device(self, eng, proofs_commitments, proofs_responses, transcript_label, instance_enc, common_enc,
                            rho_seed32=None, want_msm_inputs=False):
        """module::batch_verify with transcripts, challenges, weights and the coefficient fold on the GPU
        (zkp_batch_verify_proofs).  Same arguments as batch_verify; rho_seed32 seeds the per-proof weights."""
        com = np.ascontiguousarray(proofs_commitments, dtype=np.uint8)
        resp = np.ascontiguousarray(proofs_responses, dtype=np.uint8)
        inst = np.ascontiguousarray(instance_enc, dtype=np.uint8)
        comm = np.ascontiguousarray(common_enc, dtype=np.uint8)
        N = com.shape[0]
        if resp.shape[0] != N or inst.shape[1] != N:
            raise BatchSizeMismatch()
        assert com.shape[1:] == (self.k, 32) and resp.shape[1:] == (self.m, 32) and inst.shape[0] == len(self.instance)
        assert rho_seed32 is None or len(rho_seed32) == 32       # None: 32 fresh bytes from the OS CSPRNG
        n = len(self.common) + (len(self.instance) + self.k) * N
        co = np.zeros((n, 32), np.uint8) if want_msm_inputs else None
        po = np.zeros((n, 32), np.uint8) if want_msm_inputs else None
        rc = _lib().zkph_batch_verify_device(eng._ctx, self._h, transcript_label, len(transcript_label), N, _p(inst),
                                             _p(comm), _p(com), _p(resp), _seed(rho_seed32)[0], _p(co), _p(po))
        if want_msm_inputs and rc in (0, 1):
            self._last_msm_inputs = (co, po)
        _raise(rc)
        return (co, po) if want_msm_inputs else None

    # ---- N proofs at once (no counterpart in the reference; = N x prove_batchable) ----------------------------
    def prove_many(self, eng, transcript_label, secrets, points_limbs, entropy=None, threads=0):
        """entropy (N, 32): the 32 bytes per proof the reference takes from thread_rng (prover.rs:82); None = OS CSPRNG."""
        sec = np.ascontiguousarray(secrets, dtype=np.uint8).reshape(-1, self.m, 32)
        N = sec.shape[0]
        pts = np.ascontiguousarray(points_limbs, dtype=np.uint64).reshape(N, self.p, 20)
        ent = None if entropy is None else np.ascontiguousarray(entropy, dtype=np.uint8).reshape(N, 32)
        enc = np.zeros((N, self.p, 32), np.uint8)
        com = np.zeros((N, self.k, 32), np.uint8)
        resp = np.zeros((N, self.m, 32), np.uint8)
        _raise(_lib().zkph_prove_many(eng._ctx, self._h, transcript_label, len(transcript_label), N, _p(sec), _p(pts),
                                      _p(ent), int(threads), _p(enc), _p(com), _p(resp)))
        return enc, com, resp


    def prove_many_device(self, eng, transcript_label, secrets, points_limbs, entropy=None, out=None):
        """prove_many with the per-proof transcript, nonce and response work on the GPU (zkp_prove_batch): the host hashes
        only the batch-wide transcript prefix.  Byte-identical to prove_many.  `out` = preallocated (encodings[N][p][32],
        commitments[N][k][32], responses[N][m][32]) uint8 arrays; inputs and outputs in pinned host memory avoid the
        staging copies of pageable transfers, which otherwise dominate the call."""
        sec = np.ascontiguousarray(secrets, dtype=np.uint8).reshape(-1, self.m, 32)
        N = sec.shape[0]
        pts = np.ascontiguousarray(points_limbs, dtype=np.uint64).reshape(N, self.p, 20)
        ent = None if entropy is None else np.ascontiguousarray(entropy, dtype=np.uint8).reshape(N, 32)
        if out is not None:
            enc, com, resp = out
            for a, shp in ((enc, (N, self.p, 32)), (com, (N, self.k, 32)), (resp, (N, self.m, 32))):
                if a.dtype != np.uint8 or a.shape != shp or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError("out arrays must be C-contiguous uint8 of shapes [N][p][32], [N][k][32], [N][m][32]")
        else:
            enc, com, resp = (np.zeros((N, self.p, 32), np.uint8), np.zeros((N, self.k, 32), np.uint8),
                              np.zeros((N, self.m, 32), np.uint8))
        _raise(_lib().zkph_prove_many_device(eng._ctx, self._h, transcript_label, len(transcript_label), N, _p(sec), _p(pts),
                                             _p(ent), _p(enc), _p(com), _p(resp)))
        return enc, com, resp


# ---- wire format of src/proofs.rs (bincode 1.x defaults, as /root/reference/tests/zkp.rs:53-54, :96-97 use it) --------
def serialize_compact(challenge, responses):
    r = np.ascontiguousarray(responses, dtype=np.uint8).reshape(-1, 32)
    c = np.ascontiguousarray(challenge, dtype=np.uint8).reshape(32)
    out = np.zeros(int(_lib().zkph_compact_proof_size(r.shape[0])), np.uint8)
    _raise(_lib().zkph_compact_proof_serialize(_p(c), _p(r), r.shape[0], _p(out)))
    return out.tobytes()


def parse_compact(buf, m):
    """-> (challenge[32], responses[m][32]); VerificationFailure if the proof carries another number of responses."""
    b = np.frombuffer(bytes(buf), dtype=np.uint8)
    c, r = np.zeros(32, np.uint8), np.zeros((m, 32), np.uint8)
    m_out, used = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _raise(_lib().zkph_compact_proof_parse(_p(b), b.size, m, _p(c), _p(r), ctypes.byref(m_out), ctypes.byref(used)))
    if m_out.value != m:
        raise VerificationFailure()
    if used.value != b.size:
        raise MalformedProof()
    return c, r


def serialize_batchable(commitments, responses):
    k = np.ascontiguousarray(commitments, dtype=np.uint8).reshape(-1, 32)
    r = np.ascontiguousarray(responses, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(int(_lib().zkph_batchable_proof_size(k.shape[0], r.shape[0])), np.uint8)
    _raise(_lib().zkph_batchable_proof_serialize(_p(k), k.shape[0], _p(r), r.shape[0], _p(out)))
    return out.tobytes()


def parse_batchable_many(buf, N, k, m, threads=0):
    """N concatenated serialized BatchableProofs of one statement -> (commitments[N][k][32], responses[N][m][32]), the
    arrays batch_verify / batch_verify_device take."""
    b = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    com, resp = np.zeros((N, k, 32), np.uint8), np.zeros((N, m, 32), np.uint8)
    bad = ctypes.c_int64(-1)
    _raise(_lib().zkph_batchable_proofs_parse(_p(b), b.size, N, k, m, _p(com), _p(resp), int(threads), ctypes.byref(bad)))
    return com, resp


def dleq_statement():
    """define_proof! {dleq, "DLEQ proof", (x), (A, B, H), (G) : A = (x * G), B = (x * H)}  (benches/zkp.rs:49)"""
    return Statement("dleq", "DLEQ proof", ["x"], ["A", "B", "H"], ["G"],
                     [("A", [("x", "G")]), ("B", [("x", "H")])])


def cmz10_statement():
    """CMZ'13 credential presentation with 10 hidden attributes (/root/reference/benches/zkp.rs:27-46)."""
    rng = range(1, 11)
    return Statement(
        "cred_show_10", "CMZ cred show n=10",
        ["m_%d" % i for i in rng] + ["z_%d" % i for i in rng] + ["minus_z_Q"],
        ["C_%d" % i for i in rng] + ["P", "Q", "V"],
        ["X_%d" % i for i in rng] + ["A", "B"],
        [("C_%d" % i, [("m_%d" % i, "P"), ("z_%d" % i, "A")]) for i in rng]
        + [("V", [("m_%d" % i, "X_%d" % i) for i in rng] + [("minus_z_Q", "Q")])])
