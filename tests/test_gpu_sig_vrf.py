"""Port of /root/reference/tests/sig_and_vrf_example.rs to the product path (host mirror over the CUDA engine):
a Schnorr signature is `sig_proof` in batchable form, a VRF is `vrf_proof` in compact form; the transcripts carry
state (messages appended before proving, stateful signature chains).  Same acceptance / rejection cases as the
reference (:177-222 create_and_verify_sig, :243-281 counterparty_signature_chain, :284-385 create_and_verify_vrf)."""
import numpy as np
import pytest

from oracle import ristretto as R          # only for hash-to-group (user-side code in the reference's test)
from zkp_b200 import toolbox as PT

pytestmark = pytest.mark.gpu

B_ENC = np.frombuffer(R.BASEPOINT_COMPRESSED, dtype=np.uint8)


@pytest.fixture(scope="module")
def stmts():
    sig = PT.Statement("sig_proof", "Sig", ["x"], ["A"], ["B"], [("A", [("x", "B")])])                       # :24
    vrf = PT.Statement("vrf_proof", "VRF", ["x"], ["A", "G", "H"], ["B"], [("A", [("x", "B")]), ("G", [("x", "H")])])  # :25
    return sig, vrf


def scalar(seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, size=32, dtype=np.uint8)
    s[31] &= 0x0F
    return s


def mul(engine, k, enc):
    out, valid = engine.msm_vartime_batched(k.reshape(1, 32), np.asarray(enc, dtype=np.uint8).reshape(1, 32),
                                            np.arange(2, dtype=np.uint64))
    assert valid.all()
    return out[0]


class KeyPair:
    def __init__(self, engine, seed):
        self.eng, self.sk = engine, scalar(seed)
        self.pk = mul(engine, self.sk, B_ENC)                                           # PublicKey::from(&sk)

    def limbs(self, *encs):
        l, v = self.eng.decompress_batch(np.stack(encs))
        assert v.all()
        return l

    def sign(self, stmts, message, transcript, seed=b"s"):
        transcript.append_message(b"msg", message)                                      # append_message_example
        (com, resp), _ = stmts[0].prove_batchable(self.eng, transcript, self.sk, self.limbs(self.pk, B_ENC), seed)
        return com, resp

    def vrf(self, stmts, function_transcript, message, proof_transcript, seed=b"v"):
        function_transcript.append_message(b"msg", message)
        H = R.compress(R.from_uniform_bytes(function_transcript.challenge_bytes(b"output", 64)))   # hash_to_group
        H = np.frombuffer(H, dtype=np.uint8)
        G = mul(self.eng, self.sk, H)
        (chal, resp), enc = stmts[1].prove_compact(self.eng, proof_transcript, self.sk,
                                                   self.limbs(self.pk, G, H, B_ENC), seed)
        return enc[1].copy(), (chal, resp)                                              # (VrfOutput(points.G), proof)


def sig_verify(engine, stmts, sig, message, pk, transcript):
    transcript.append_message(b"msg", message)
    try:
        stmts[0].verify_batchable(engine, sig, transcript, np.stack([pk, B_ENC]), b"w")
        return True
    except PT.ProofError:
        return False


def vrf_verify(engine, stmts, output, function_transcript, message, pk, proof_transcript, proof):
    function_transcript.append_message(b"msg", message)
    H = np.frombuffer(R.compress(R.from_uniform_bytes(function_transcript.challenge_bytes(b"output", 64))), dtype=np.uint8)
    try:
        stmts[1].verify_compact(engine, proof, proof_transcript, np.stack([pk, output, H, B_ENC]))
        return True
    except PT.ProofError:
        return False


def test_create_and_verify_sig(engine, stmts):
    T = PT.Transcript
    dom, msg1, msg2 = b"My Sig Application", b"Test Message 1", b"Test Message 2"
    kp1, kp2 = KeyPair(engine, 1), KeyPair(engine, 2)
    sig1 = kp1.sign(stmts, msg1, T(dom))
    sig2 = kp2.sign(stmts, msg2, T(dom))
    assert sig_verify(engine, stmts, sig1, msg1, kp1.pk, T(dom))
    assert sig_verify(engine, stmts, sig2, msg2, kp2.pk, T(dom))
    assert not sig_verify(engine, stmts, sig1, msg1, kp2.pk, T(dom))        # wrong public key
    assert not sig_verify(engine, stmts, sig2, msg2, kp1.pk, T(dom))
    assert not sig_verify(engine, stmts, sig1, msg2, kp1.pk, T(dom))        # wrong message
    assert not sig_verify(engine, stmts, sig2, msg1, kp2.pk, T(dom))
    assert not sig_verify(engine, stmts, sig1, msg1, kp1.pk, T(b"Wrong"))   # wrong domain separator
    assert not sig_verify(engine, stmts, sig2, msg2, kp2.pk, T(b"Wrong"))


def test_counterparty_signature_chain(engine, stmts):
    msgs = [b"In this test, two counterparties exchange signatures.", b"However, the counterparties sign and verify messages",
            b"using stateful transcript objects.", b"When party 1 signs, the party 1 transcript changes;",
            b"when party 2 verifies, the party 2 transcript syncs.",
            b"In this way, the transcript states ratchet stateful signatures."]
    kp1, kp2 = KeyPair(engine, 3), KeyPair(engine, 4)
    trans1, trans2 = PT.Transcript(b"Counterparty Example"), PT.Transcript(b"Counterparty Example")
    for i in range(0, 6, 2):
        s1 = kp1.sign(stmts, msgs[i], trans1, seed=b"a%d" % i)
        assert sig_verify(engine, stmts, s1, msgs[i], kp1.pk, trans2)
        s2 = kp2.sign(stmts, msgs[i + 1], trans2, seed=b"b%d" % i)
        assert sig_verify(engine, stmts, s2, msgs[i + 1], kp2.pk, trans1)
    # the two transcripts stayed in lock-step
    assert trans1.challenge_bytes(b"sync", 32) == trans2.challenge_bytes(b"sync", 32)


def test_create_and_verify_vrf(engine, stmts):
    T = PT.Transcript
    dom, msg1, msg2 = b"My VRF Application", b"Test Message 1", b"Test Message 2"
    kp1, kp2 = KeyPair(engine, 5), KeyPair(engine, 6)
    out1, proof1 = kp1.vrf(stmts, T(dom), msg1, T(dom))
    out2, proof2 = kp2.vrf(stmts, T(dom), msg2, T(dom))
    assert vrf_verify(engine, stmts, out1, T(dom), msg1, kp1.pk, T(dom), proof1)
    assert vrf_verify(engine, stmts, out2, T(dom), msg2, kp2.pk, T(dom), proof2)
    assert not vrf_verify(engine, stmts, out1, T(dom), msg1, kp2.pk, T(dom), proof1)      # swap pubkey
    assert not vrf_verify(engine, stmts, out2, T(dom), msg2, kp1.pk, T(dom), proof2)
    assert not vrf_verify(engine, stmts, out2, T(dom), msg1, kp1.pk, T(dom), proof1)      # swap output
    assert not vrf_verify(engine, stmts, out1, T(dom), msg2, kp2.pk, T(dom), proof2)
    assert not vrf_verify(engine, stmts, out1, T(dom), msg1, kp1.pk, T(b"A different application"), proof1)
    assert not vrf_verify(engine, stmts, out2, T(dom), msg2, kp2.pk, T(b"A different application"), proof2)
