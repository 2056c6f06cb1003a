"""The reference's constraint-system callers, restated over the big-int oracle.  TEST INFRASTRUCTURE.

Follows, line for line in behaviour (not in code):
  TranscriptProtocol  /root/reference/src/toolbox/mod.rs:165-228
  Prover              /root/reference/src/toolbox/prover.rs:41-142
  Verifier            /root/reference/src/toolbox/verifier.rs:47-183
  BatchVerifier       /root/reference/src/toolbox/batch_verifier.rs:67-245
Randomness the reference draws from thread_rng (prover.rs:82, verifier.rs:153, batch_verifier.rs:179) is
injected through `rng` (an object with .bytes(n)) so that the product and the oracle can be fed the same
bytes and compared byte for byte.
"""
from . import ristretto as R
from . import scalar as S
from . import msm as M
from .merlin import Transcript


class ProofError(Exception):
    pass


class VerificationFailure(ProofError):
    pass


class BatchSizeMismatch(ProofError):
    pass


class SeededRng:
    """Deterministic byte source (SHAKE-256 stream) standing in for rand::thread_rng."""

    def __init__(self, seed: bytes):
        import hashlib
        self._h = hashlib.shake_256(seed)
        self._off = 0

    def bytes(self, n):
        out = self._h.digest(self._off + n)[self._off:]
        self._off += n
        return out

    def u128(self):
        return int.from_bytes(self.bytes(16), "little")


class PerProofRng:
    """The weight derivation of zkp_batch_verify_proofs: rho_(i,j) = bytes [16i, 16i+16) of
    SHAKE256(seed32 || le64(j)) -- a parallel stand-in for rand::thread_rng (batch_verifier.rs:179)."""

    def __init__(self, seed32: bytes):
        assert len(seed32) == 32
        self.seed = bytes(seed32)

    def rho(self, i, j):
        import hashlib
        return int.from_bytes(hashlib.shake_256(self.seed + int(j).to_bytes(8, "little")).digest(16 * (i + 1))[16 * i:], "little")


# ---- TranscriptProtocol (toolbox/mod.rs:165-228) ---------------------------------------------------------
def domain_sep(t, label):
    t.append_message(b"dom-sep", b"schnorrzkp/1.0/ristretto255")
    t.append_message(b"dom-sep", label)


def append_scalar_var(t, label):
    t.append_message(b"scvar", label)


def append_point_var(t, label, point):
    enc = R.compress(point)
    t.append_message(b"ptvar", label)
    t.append_message(b"val", enc)
    return enc


def validate_and_append_point_var(t, label, enc):
    if R.is_identity_encoding(enc):
        raise VerificationFailure()
    t.append_message(b"ptvar", label)
    t.append_message(b"val", enc)


def append_blinding_commitment(t, label, point):
    enc = R.compress(point)
    t.append_message(b"blindcom", label)
    t.append_message(b"val", enc)
    return enc


def validate_and_append_blinding_commitment(t, label, enc):
    if R.is_identity_encoding(enc):
        raise VerificationFailure()
    t.append_message(b"blindcom", label)
    t.append_message(b"val", enc)


def get_challenge(t, label):
    return S.from_bytes_mod_order_wide(t.challenge_bytes(label, 64))


class CompactProof:
    def __init__(self, challenge, responses):
        self.challenge, self.responses = challenge, list(responses)


class BatchableProof:
    def __init__(self, commitments, responses):
        self.commitments, self.responses = list(commitments), list(responses)


class Prover:
    def __init__(self, proof_label, transcript):
        domain_sep(transcript, proof_label)
        self.transcript = transcript
        self.scalars, self.points, self.point_labels, self.constraints = [], [], [], []

    def allocate_scalar(self, label, assignment):
        append_scalar_var(self.transcript, label)
        self.scalars.append(assignment % R.L)
        return len(self.scalars) - 1

    def allocate_point(self, label, assignment):
        enc = append_point_var(self.transcript, label, assignment)
        self.points.append(assignment)
        self.point_labels.append(label)
        return len(self.points) - 1, enc

    def constrain(self, lhs, linear_combination):
        self.constraints.append((lhs, list(linear_combination)))

    def _prove_impl(self, rng):
        b = self.transcript.build_rng()
        for s in self.scalars:
            b = b.rekey_with_witness_bytes(b"", S.to_bytes(s))
        trng = b.finalize(rng.bytes(32))
        blindings = [S.from_bytes_mod_order_wide(trng.fill_bytes(64)) for _ in self.scalars]
        commitments = []
        for lhs, rhs in self.constraints:
            com = M.multiscalar_mul([blindings[sc] for sc, _ in rhs], [self.points[pt] for _, pt in rhs])
            commitments.append(append_blinding_commitment(self.transcript, self.point_labels[lhs], com))
        challenge = get_challenge(self.transcript, b"chal")
        responses = [(s * challenge + bl) % R.L for s, bl in zip(self.scalars, blindings)]
        return challenge, responses, commitments, blindings

    def prove_compact(self, rng):
        c, r, _, _ = self._prove_impl(rng)
        return CompactProof(c, r)

    def prove_batchable(self, rng):
        _, r, com, _ = self._prove_impl(rng)
        return BatchableProof(com, r)


class Verifier:
    def __init__(self, proof_label, transcript):
        domain_sep(transcript, proof_label)
        self.transcript = transcript
        self.num_scalars = 0
        self.points, self.point_labels, self.constraints = [], [], []

    def allocate_scalar(self, label):
        append_scalar_var(self.transcript, label)
        self.num_scalars += 1
        return self.num_scalars - 1

    def allocate_point(self, label, enc):
        validate_and_append_point_var(self.transcript, label, enc)
        self.points.append(bytes(enc))
        self.point_labels.append(label)
        return len(self.points) - 1

    def constrain(self, lhs, linear_combination):
        self.constraints.append((lhs, list(linear_combination)))

    def verify_compact(self, proof):
        if len(proof.responses) != self.num_scalars:
            raise VerificationFailure()
        pts = [R.decompress(e) for e in self.points]
        if any(p is None for p in pts):
            raise VerificationFailure()
        minus_c = S.neg(proof.challenge)
        for lhs, rhs in self.constraints:
            com = M.vartime_multiscalar_mul(
                [proof.responses[sc] for sc, _ in rhs] + [minus_c],
                [pts[pt] for _, pt in rhs] + [pts[lhs]])
            append_blinding_commitment(self.transcript, self.point_labels[lhs], com)
        if get_challenge(self.transcript, b"chal") != proof.challenge:
            raise VerificationFailure()

    def batchable_coeffs(self, proof, rng):
        """verifier.rs:125-160 -- everything before the MSM; returns (coeffs, encodings)."""
        if len(proof.responses) != self.num_scalars:
            raise VerificationFailure()
        if len(proof.commitments) != len(self.constraints):
            raise VerificationFailure()
        for i, com in enumerate(proof.commitments):
            validate_and_append_blinding_commitment(self.transcript, self.point_labels[self.constraints[i][0]], com)
        minus_c = S.neg(get_challenge(self.transcript, b"chal"))
        off = len(self.points)
        coeffs = [0] * (off + len(proof.commitments))
        for i, (lhs, rhs) in enumerate(self.constraints):
            rho = rng.u128()
            coeffs[off + i] = (coeffs[off + i] - rho) % R.L
            coeffs[lhs] = (coeffs[lhs] + rho * minus_c) % R.L
            for sc, pt in rhs:
                coeffs[pt] = (coeffs[pt] + rho * proof.responses[sc]) % R.L
        return coeffs, self.points + [bytes(c) for c in proof.commitments]

    def verify_batchable(self, proof, rng):
        coeffs, encs = self.batchable_coeffs(proof, rng)
        check = M.optional_multiscalar_mul(coeffs, [R.decompress(e) for e in encs])
        if check is None or not R.is_identity(check):
            raise VerificationFailure()


class BatchVerifier:
    def __init__(self, proof_label, batch_size, transcripts):
        if len(transcripts) != batch_size:
            raise BatchSizeMismatch()
        for t in transcripts:
            domain_sep(t, proof_label)
        self.batch_size, self.transcripts = batch_size, transcripts
        self.num_scalars = 0
        self.static_points, self.static_point_labels = [], []
        self.instance_points, self.instance_point_labels = [], []
        self.constraints = []

    def allocate_scalar(self, label):
        for t in self.transcripts:
            append_scalar_var(t, label)
        self.num_scalars += 1
        return self.num_scalars - 1

    def allocate_static_point(self, label, enc):
        for t in self.transcripts:
            validate_and_append_point_var(t, label, enc)
        self.static_points.append(bytes(enc))
        self.static_point_labels.append(label)
        return ("S", len(self.static_points) - 1)

    def allocate_instance_point(self, label, encs):
        if len(encs) != self.batch_size:
            raise BatchSizeMismatch()
        for t, e in zip(self.transcripts, encs):
            validate_and_append_point_var(t, label, e)
        self.instance_points.append([bytes(e) for e in encs])
        self.instance_point_labels.append(label)
        return ("I", len(self.instance_points) - 1)

    def constrain(self, lhs, linear_combination):
        self.constraints.append((lhs, list(linear_combination)))

    def _label(self, var):
        return self.static_point_labels[var[1]] if var[0] == "S" else self.instance_point_labels[var[1]]

    def batch_coeffs(self, proofs, rng):
        """batch_verifier.rs:138-217 -- everything before the MSM; returns (scalars, encodings) in the exact
        order the reference feeds optional_multiscalar_mul (static ++ row-major instance matrix)."""
        if len(proofs) != self.batch_size:
            raise BatchSizeMismatch()
        for pr in proofs:
            if len(pr.commitments) != len(self.constraints) or len(pr.responses) != self.num_scalars:
                raise VerificationFailure()
        for j in range(self.batch_size):
            for i, com in enumerate(proofs[j].commitments):
                validate_and_append_blinding_commitment(self.transcripts[j], self._label(self.constraints[i][0]), com)
        minus_c = [S.neg(get_challenge(t, b"chal")) for t in self.transcripts]
        num_s, num_i, num_c = len(self.static_points), len(self.instance_points), len(self.constraints)
        N = self.batch_size
        static_coeffs = [0] * num_s
        inst = [[0] * N for _ in range(num_i + num_c)]
        for i, (lhs, rhs) in enumerate(self.constraints):
            for j in range(N):
                rho = rng.rho(i, j) if hasattr(rng, "rho") else rng.u128()
                inst[num_i + i][j] = (inst[num_i + i][j] - rho) % R.L
                if lhs[0] == "S":
                    static_coeffs[lhs[1]] = (static_coeffs[lhs[1]] + rho * minus_c[j]) % R.L
                else:
                    inst[lhs[1]][j] = (inst[lhs[1]][j] + rho * minus_c[j]) % R.L
                for sc, pt in rhs:
                    resp = proofs[j].responses[sc]
                    if pt[0] == "S":
                        static_coeffs[pt[1]] = (static_coeffs[pt[1]] + rho * resp) % R.L
                    else:
                        inst[pt[1]][j] = (inst[pt[1]][j] + rho * resp) % R.L
        rows = [list(r) for r in self.instance_points]
        for i in range(num_c):
            rows.append([bytes(pr.commitments[i]) for pr in proofs])
        scalars = static_coeffs + [c for row in inst for c in row]
        encs = list(self.static_points) + [e for row in rows for e in row]
        return scalars, encs

    def verify_batchable(self, proofs, rng):
        scalars, encs = self.batch_coeffs(proofs, rng)
        check = M.optional_multiscalar_mul(scalars, [R.decompress(e) for e in encs])
        if check is None or not R.is_identity(check):
            raise VerificationFailure()


# ---- define_proof! mirror (/root/reference/src/macros.rs:74-370) ---------------------------------------
class Statement:
    """What `define_proof!{name, label, (secrets), (instance), (common) : lhs = (s*P + ...), ...}` expands to.
    Allocation order secrets -> instance -> common and labels = stringify!(var) follow macros.rs:206-258,
    :280-311, :336-370; they determine the transcript byte stream."""

    def __init__(self, name, label, secrets, instance, common, constraints):
        self.name, self.label = name, label.encode()
        self.secrets, self.instance, self.common = list(secrets), list(instance), list(common)
        self.constraints = [(lhs, list(rhs)) for lhs, rhs in constraints]  # (lhs_name, [(secret_name, point_name)])

    def _constrain(self, cs, svars, pvars):
        for lhs, rhs in self.constraints:
            cs.constrain(pvars[lhs], [(svars[s], pvars[p]) for s, p in rhs])

    def build_prover(self, transcript, secrets, points):
        pr = Prover(self.label, transcript)
        svars = {n: pr.allocate_scalar(n.encode(), secrets[n]) for n in self.secrets}
        pvars, enc = {}, {}
        for n in self.instance + self.common:
            pvars[n], enc[n] = pr.allocate_point(n.encode(), points[n])
        self._constrain(pr, svars, pvars)
        return pr, enc

    def prove_compact(self, transcript, secrets, points, rng):
        pr, enc = self.build_prover(transcript, secrets, points)
        return pr.prove_compact(rng), enc

    def prove_batchable(self, transcript, secrets, points, rng):
        pr, enc = self.build_prover(transcript, secrets, points)
        return pr.prove_batchable(rng), enc

    def build_verifier(self, transcript, encs):
        v = Verifier(self.label, transcript)
        svars = {n: v.allocate_scalar(n.encode()) for n in self.secrets}
        pvars = {n: v.allocate_point(n.encode(), encs[n]) for n in self.instance + self.common}
        self._constrain(v, svars, pvars)
        return v

    def verify_compact(self, proof, transcript, encs):
        self.build_verifier(transcript, encs).verify_compact(proof)

    def verify_batchable(self, proof, transcript, encs, rng):
        self.build_verifier(transcript, encs).verify_batchable(proof, rng)

    def build_batch_verifier(self, proofs_len, transcripts, encs):
        bv = BatchVerifier(self.label, proofs_len, transcripts)
        svars = {n: bv.allocate_scalar(n.encode()) for n in self.secrets}
        pvars = {}
        for n in self.instance:
            pvars[n] = bv.allocate_instance_point(n.encode(), encs[n])
        for n in self.common:
            pvars[n] = bv.allocate_static_point(n.encode(), encs[n])
        self._constrain(bv, svars, pvars)
        return bv

    def batch_verify(self, proofs, transcripts, encs, rng):
        self.build_batch_verifier(len(proofs), transcripts, encs).verify_batchable(proofs, rng)


# benches/zkp.rs:49 / tests/zkp.rs:28
DLEQ = Statement("dleq", "DLEQ proof", ["x"], ["A", "B", "H"], ["G"],
                 [("A", [("x", "G")]), ("B", [("x", "H")])])

# benches/zkp.rs:27-46 -- CMZ'13 credential presentation with 10 hidden attributes
CMZ10 = Statement(
    "cred_show_10", "CMZ cred show n=10",
    ["m_%d" % i for i in range(1, 11)] + ["z_%d" % i for i in range(1, 11)] + ["minus_z_Q"],
    ["C_%d" % i for i in range(1, 11)] + ["P", "Q", "V"],
    ["X_%d" % i for i in range(1, 11)] + ["A", "B"],
    [("C_%d" % i, [("m_%d" % i, "P"), ("z_%d" % i, "A")]) for i in range(1, 11)]
    + [("V", [("m_%d" % i, "X_%d" % i) for i in range(1, 11)] + [("minus_z_Q", "Q")])],
)


def dleq_statement(cs, x, A, B, G, H):
    """benches/dleq.rs:37-47 (hand-written constraint-API form)."""
    cs.constrain(A, [(x, G)])
    cs.constrain(B, [(x, H)])


# ---- wire format of /root/reference/src/proofs.rs:14-32 under bincode 1.x defaults (tests/zkp.rs:53-54, :96-97) ---------
# little-endian fixed-width integers; Vec<T> = u64 length + items; Scalar / CompressedRistretto = 32 raw bytes
# (curve25519-dalek 2.x serde impls [ext]; Scalar deserialisation rejects non-canonical encodings).
def serialize_compact(challenge, responses):
    return S.to_bytes(challenge) + len(responses).to_bytes(8, "little") + b"".join(S.to_bytes(r) for r in responses)


def serialize_batchable(proof):
    return (len(proof.commitments).to_bytes(8, "little") + b"".join(proof.commitments)
            + len(proof.responses).to_bytes(8, "little") + b"".join(S.to_bytes(r) for r in proof.responses))


def _scalar_canonical(b):
    v = int.from_bytes(b, "little")
    if v >= R.L:
        raise ValueError("scalar was not canonically encoded")
    return v


def parse_batchable(buf, off=0):
    """-> (BatchableProof, next offset); raises ValueError like bincode::deserialize would."""
    def take(n):
        nonlocal off
        if off + n > len(buf):
            raise ValueError("unexpected end of input")
        b = buf[off:off + n]
        off += n
        return b
    k = int.from_bytes(take(8), "little")
    if k > (len(buf) - off) // 32:
        raise ValueError("length prefix exceeds input")
    coms = [bytes(take(32)) for _ in range(k)]
    m = int.from_bytes(take(8), "little")
    if m > (len(buf) - off) // 32:
        raise ValueError("length prefix exceeds input")
    resp = [_scalar_canonical(take(32)) for _ in range(m)]
    return BatchableProof(coms, resp), off
