"""Generate the committed golden fixtures under tests/golden/.

Run from the repo root:  python tests/golden/make_golden.py
Everything here is produced by the oracle (oracle/*.py) and cross-checked, while generating, against
libsodium 1.0.20's independent ristretto255 implementation and the constants published in RFC 9496
Appendix A (recalled constants are asserted equal to both implementations, so a recall error cannot slip
in).  The reference crate itself cannot run in this image (no rustc), so no fixture is reference-produced:
see oracle/__init__.py "Pinning status".
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import merlin, msm as M, ristretto as R, scalar as S, sodium, toolbox as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

RFC_MULTIPLES = [
    "0000000000000000000000000000000000000000000000000000000000000000",
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
    "da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57",
    "e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e",
    "f64746d3c92b13050ed8d80236a7f0007c3b3f962f5ba793d19a601ebb1df403",
    "44f53520926ec81fbd5a387845beb7df85a96a24ece18738bdcfa6a7822a176d",
    "903293d8f2287ebe10e2374dc1a53e0bc887e592699f02d077d5263cdd55601c",
    "02622ace8f7303a31cafc63f8fc48fdc16e1c8c8d234b2f0d6685282a9076031",
    "20706fd788b2720a1ed2a5dad4952b01f413bcf0e7564de8cdc816689e2db95f",
    "bce83f8ba5dd2fa572864c24ba1810f9522bc6004afe95877ac73241cafdab42",
    "e4549ee16b9aa03099ca208c67adafcafa4c3f3e4e5303de6026e3ca8ff84460",
    "aa52e000df2e16f55fb1032fc33bc42742dad6bd5a8fc0be0167436c5948501f",
    "46376b80f409b29dc2b5f6f0c52591990896e5716f41477cd30085ab7f10301e",
    "e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e",
]

# RFC 9496 A.3 style invalid encodings: non-canonical field elements, negative field elements, and strings
# whose decoding fails the square / sign / zero checks.  Each is asserted invalid by oracle AND libsodium.
BAD_ENCODINGS = [
    "00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
    "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "f3ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "0100000000000000000000000000000000000000000000000000000000000000",
    "01ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "ed57ffd8c914fb201471d1c3d245ce3c746fcbe63a3679d51b6a516ebebe0e20",
    "c34c4e1826e5d403b78e246e88aa051c36ccf0aafebffe137d148a2bf9104562",
    "c940e5a4404157cfb1628b108db051a8d439e1a421394ec4ebccb9ec92a8ac78",
    "47cfc5497c53dc8e61c91d17fd626ffb1c49e2bca94eed052281b510b1117a24",
    "f1c6165d33367351b0da8f6e4511010c68174a03b6581212c71c0e1d026c3c72",
    "87260f7a2f12495118360f02c26a470f450dadf34a413d21042b43b9d93e1309",
    "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371",
    "4eac077a713c57b4f4397629a4145982c661f48044dd3f96427d40b147d9742f",
    "de6a7b00deadc788eb6b6c8d20c8ae50ed8ff6f9ab2a3ed3b0ffc5d1b6bbd37d",
    "bcab477be20861e01e4a0e295284146a510150d9817763caf1a6f4b422d67042",
    "2a292df7e32cababbd9de088d1d1abec9fc0440f637ed2fba145094dc14bea08",
    "f4a9e534fc0d216c44b218fa0c42d99635a0127ee2e53c712f70609649fdff22",
    "8268436f8c4126196cf64b3c7ddbda90746a378625f9813dd9b8457077256731",
    "2810e5cbc2cc4d4eece54f61c6f69758e289aa7ab440b3cbeaa21995c2f4232b",
    "3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e",
    "a45fdc55c76448c049a1ab33f17023edfb2be3581e9c7aade8a6125215e04220",
    "d483fe813c6ba647ebbfd3ec41adca1c6130c2beeee9d9bf065c8d151c5f396e",
    "8a2e1d30050198c65a54483123960ccc38aef6848e1ec8f5f780e8523769ba32",
    "32888462f8b486c68ad7dd9610be5192bbeaf3b443951ac1a8118419d9fa097b",
    "227142501b9d4355ccba290404bde41575b037693cef1f438c47f8fbf35d1165",
    "5c37cc491da847cfeb9281d407efc41e15144c876e0170b499a96a22ed31e01e",
    "445425117cb8c90edcbc7c1cc0e74f747f2c1efa5630a967c64f287792a48a4b",
]

HASH_TO_GROUP = [
    ("Ristretto is traditionally a short shot of espresso coffee",
     "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"),
    ("A VRF input, for instance",  # /root/reference/tests/zkp.rs:34
     "8062d869a1a967d6a60604a3ec8d0316cef712e094f4cb991de60a9f52555068"),
]


def hexs(bs):
    return [bytes(b).hex() for b in bs]


def seeded_points(k, seed):
    return [R.compress(R.from_uniform_bytes(hashlib.sha512(seed + i.to_bytes(8, "little")).digest()))
            for i in range(k)]


def seeded_scalars(n, seed, bits=None):
    out = []
    for i in range(n):
        v = int.from_bytes(hashlib.sha512(seed + b"/s/" + i.to_bytes(8, "little")).digest(), "little")
        out.append(v % R.L if bits is None else v % (1 << bits))
    return out


def main():
    na = sodium.load()
    # ---- group vectors ---------------------------------------------------------------------------------
    assert R.compress(R.IDENTITY).hex() == RFC_MULTIPLES[0]
    for k in range(1, 16):
        enc = R.compress(R.pt_mul(k, R.BASEPOINT))
        assert enc.hex() == RFC_MULTIPLES[k], ("oracle vs RFC", k)
        assert sodium.scalarmult_base(na, k) == enc, ("libsodium vs RFC", k)
    bad_ok = []
    for h in BAD_ENCODINGS:
        b = bytes.fromhex(h)
        o, s = R.decompress(b) is None, not sodium.is_valid_point(na, b)
        assert o == s, ("oracle/libsodium disagree", h)
        if o:
            bad_ok.append(h)
    assert len(bad_ok) >= 12, len(bad_ok)
    for msg, exp in HASH_TO_GROUP:
        e = R.compress(R.hash_from_bytes_sha512(msg.encode()))
        assert e.hex() == exp and sodium.from_hash(na, hashlib.sha512(msg.encode()).digest()) == e
    json.dump({"source": "RFC 9496 Appendix A (asserted equal to oracle and libsodium 1.0.20 when generated)",
               "multiples_of_generator": RFC_MULTIPLES, "bad_encodings": bad_ok,
               "hash_to_group_sha512": [{"msg": m, "enc": e} for m, e in HASH_TO_GROUP]},
              open(os.path.join(OUT, "rfc9496.json"), "w"), indent=1)

    # ---- MSM known-answer tests ------------------------------------------------------------------------
    kats = []
    edge = [0, 1, R.L - 1, R.L // 2, R.L // 2 + 1, 2**128 - 1, R.L - (2**128 - 1), 2**252, 8, 2**200]
    for n in [1, 2, 3, 6, 7, 12, 36, 100, 189, 190, 191, 255, 256, 300, 499, 500, 600]:
        seed = b"msm-kat-%d" % n
        pts = seeded_points(n, seed)
        ks = seeded_scalars(n, seed)
        for j, e in enumerate(edge):
            if j < n and n in (36, 300):
                ks[j] = e
        dpts = [R.decompress(p) for p in pts]
        res = M.optional_multiscalar_mul(ks, dpts)         # dalek's algorithm choice (Straus < 190 <= Pippenger)
        enc = R.compress(res)
        assert enc == R.compress(M.naive_msm(ks, dpts))
        if n <= 36:
            assert enc == sodium.msm(na, ks, pts), ("libsodium", n)
            assert enc == R.compress(M.straus_ct(ks, dpts))
        kats.append({"n": n, "scalars": hexs(S.to_bytes(k) for k in ks), "points": hexs(pts), "expected": enc.hex(),
                     "checked_by": "oracle dalek-algorithm + naive" + (" + libsodium + straus_ct" if n <= 36 else "")})
    # 128-bit weights (the -rho rows of batch_verifier.rs:183) and a sum that is the identity
    n = 64
    pts = seeded_points(n, b"msm-kat-rho")
    ks = [R.L - r for r in seeded_scalars(n, b"rho", bits=128)]
    enc = R.compress(M.naive_msm(ks, [R.decompress(p) for p in pts]))
    kats.append({"n": n, "scalars": hexs(S.to_bytes(k) for k in ks), "points": hexs(pts), "expected": enc.hex(),
                 "checked_by": "oracle naive", "note": "coefficients -rho with 128-bit rho"})
    pts2 = pts[:8] + pts[:8]
    ks2 = seeded_scalars(8, b"cancel")
    ks2 = ks2 + [R.L - k for k in ks2]
    enc = R.compress(M.naive_msm(ks2, [R.decompress(p) for p in pts2]))
    assert enc == bytes(32)
    kats.append({"n": 16, "scalars": hexs(S.to_bytes(k) for k in ks2), "points": hexs(pts2), "expected": enc.hex(),
                 "checked_by": "oracle naive", "note": "sum is the identity"})
    json.dump({"kats": kats}, open(os.path.join(OUT, "msm_kat.json"), "w"))

    # seeded large cases: inputs are regenerated from the seed by tests/util_data.py, only the answer is stored
    from tests import util_data as U
    big = []
    for n, K in [(4096, 64), (20000, 128)]:
        base = U.base_points(K)
        sc = U.random_scalars(n, seed=n)
        big.append({"n": n, "K": K, "seed": n, "expected": U.tiled_expected(sc, base).hex()})
        if n == 4096:  # cross-check the tiling reduction itself against a straight sum
            ks = U.scalars_to_ints(sc)
            pts = [R.decompress(base[i % K]) for i in range(n)]
            assert R.compress(M.pippenger(ks, pts)).hex() == big[-1]["expected"]
    json.dump({"cases": big}, open(os.path.join(OUT, "msm_seeded.json"), "w"), indent=1)

    # ---- Merlin ----------------------------------------------------------------------------------------
    t = merlin.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    simple = t.challenge_bytes(b"challenge", 32).hex()
    t = merlin.Transcript(b"test protocol")
    t.append_message(b"step1", b"some data")
    for _ in range(32):
        ch = t.challenge_bytes(b"challenge", 32)
        t.append_message(b"bigdata", b"\x63" * 1024)
        t.append_message(b"challengedata", ch)
    assert ch.hex() == "a8c933f54fae76e3f9bea93648c1308e7dfa2152dd51674ff3ca438351cf003c"  # published vector
    json.dump({"simple": simple, "complex": ch.hex(),
               "note": "complex = merlin's published cross-implementation conformance vector"},
              open(os.path.join(OUT, "merlin.json"), "w"), indent=1)

    # ---- toolbox known answers (DLEQ + CMZ), injected randomness -----------------------------------------
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    x = 89327492234
    A, B = R.pt_mul(x, G), R.pt_mul(x, H)
    rng = T.SeededRng(b"golden-dleq")
    tr = merlin.Transcript(b"DLEQTest")
    pr = T.Prover(b"DLEQProof", tr)
    vx = pr.allocate_scalar(b"x", x)
    vG, eG = pr.allocate_point(b"G", G)
    vH, eH = pr.allocate_point(b"H", H)
    vA, eA = pr.allocate_point(b"A", A)
    vB, eB = pr.allocate_point(b"B", B)
    T.dleq_statement(pr, vx, vA, vB, vG, vH)
    chal, resp, coms, blind = pr._prove_impl(rng)
    dleq = {"x": x, "G": eG.hex(), "H": eH.hex(), "A": eA.hex(), "B": eB.hex(), "rng_seed": "golden-dleq",
            "challenge": S.to_bytes(chal).hex(), "responses": hexs(S.to_bytes(r) for r in resp),
            "commitments": hexs(coms), "blindings": hexs(S.to_bytes(b) for b in blind)}
    # batch of 4 macro-form DLEQ proofs (tests/zkp.rs:115-175) with the coefficient vector of the batch MSM
    rngb = T.SeededRng(b"golden-dleq-batch")
    proofs, encs = [], {k: [] for k in "ABH"}
    for j in range(4):
        xj = x + j
        pts = dict(A=R.pt_mul(xj, G), B=R.pt_mul(xj, H), H=H, G=G)
        p, e = T.DLEQ.prove_batchable(merlin.Transcript(b"DLEQBatchTest"), dict(x=xj), pts, rngb)
        proofs.append(p)
        for k in "ABH":
            encs[k].append(e[k])
    encs["G"] = e["G"]
    bv = T.DLEQ.build_batch_verifier(4, [merlin.Transcript(b"DLEQBatchTest") for _ in range(4)], encs)
    rngv = T.SeededRng(b"golden-dleq-batch-verify")
    scal, pts_enc = bv.batch_coeffs(proofs, rngv)
    chk = M.optional_multiscalar_mul(scal, [R.decompress(q) for q in pts_enc])
    assert R.is_identity(chk)
    dleq_batch = {"proofs": [{"commitments": hexs(p.commitments), "responses": hexs(S.to_bytes(r) for r in p.responses)}
                             for p in proofs],
                  "A": hexs(encs["A"]), "B": hexs(encs["B"]), "H": hexs(encs["H"]), "G": encs["G"].hex(),
                  "verify_rng_seed": "golden-dleq-batch-verify",
                  "msm_scalars": hexs(S.to_bytes(s) for s in scal), "msm_points": hexs(pts_enc),
                  "msm_result": R.compress(chk).hex()}
    json.dump({"dleq_compact": dleq, "dleq_batch": dleq_batch}, open(os.path.join(OUT, "toolbox_kat.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
