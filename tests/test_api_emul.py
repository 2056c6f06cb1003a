"""The whole C ABI of zkp_b200/csrc/api.cu on the CPU: the kernel launches are rewritten into host calls
(tests/host_emul/make_api_emul.py), the CUDA runtime is a host stand-in (cudart_shim.cpp), the kernels run through
cuda_shim.h.  What this checks without a GPU is everything api.cu itself does -- argument handling, workspace sizing,
the chunk pipelines of the host-input paths, the plans handed to the kernels, the launch sequences, the status codes --
against the oracle.  The emulation library is test infrastructure: nothing in zkp_b200/ can load it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import cref, merlin as OM, msm as M, ristretto as R, scalar as S, toolbox as OT
from tests import util_data as U
from zkp_b200 import native

# Every emulated MSM costs ~10 s (a 1024-thread block is 1024 OS threads): the default suite keeps one case per entry
# point, ZKP_SLOW_TESTS=1 adds the rest (all of them are also -m gpu tests on the device).
SLOW = os.environ.get("ZKP_SLOW_TESTS") == "1"
HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "host_emul")
ROOT = os.path.dirname(HERE)
vp, sz, i32, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_int64


class Desc(ctypes.Structure):
    _fields_ = [("m", i32), ("ni", i32), ("nc", i32), ("k", i32), ("labels", ctypes.c_char_p), ("lhs", vp), ("cons_off", vp),
                ("term_scalar", vp), ("term_point", vp)]


@pytest.fixture(scope="module")
def api():
    if not U.can_spawn_threads():
        pytest.skip("this environment does not allow ~1100 threads per process (needed to emulate 1024-thread blocks)")
    out = os.path.join(EMUL, "libapi_emul.so")
    csrc = os.path.join(ROOT, "zkp_b200", "csrc")
    deps = [os.path.join(EMUL, f) for f in ("make_api_emul.py", "cuda_shim.h", "cudart_shim.cpp")]
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".hpp"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["python", os.path.join(EMUL, "make_api_emul.py")])
        # -Bsymbolic: the library's cuda* calls must reach ITS runtime stand-in even when a real libcudart is in the process
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-Wl,-Bsymbolic",
                               "-DZKP_HOST_EMUL", "-DZKP_ABLATIONS",
                               "-I/usr/local/cuda/include", "-I" + csrc, "-I" + EMUL,
                               os.path.join(EMUL, "api_emul_generated.cpp"), os.path.join(EMUL, "cudart_shim.cpp"),
                               os.path.join(csrc, "host", "merlin.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    for name in native.SYMBOLS:                       # the emulated library exports the whole ABI as well
        assert hasattr(lib, name), name
    P = ctypes.POINTER
    lib.zkp_ctx_create.argtypes = [P(vp), i32]
    lib.zkp_ctx_destroy.argtypes = [vp]
    lib.zkp_ctx_destroy.restype = None
    lib.zkp_ctx_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.zkp_msm_vartime.argtypes = [vp, vp, vp, sz, vp, P(i32), P(i64)]
    lib.zkp_batch_verify.argtypes = [vp, vp, vp, sz, vp, vp, sz, sz, P(i32), P(i64)]
    lib.zkp_batch_verify_partial.argtypes = [vp, vp, vp, sz, vp, vp, sz, sz, vp, P(i64)]
    lib.zkp_partials_verdict.argtypes = [vp, vp, sz, P(i32), vp]
    lib.zkp_msm_vartime_batched.argtypes = [vp, vp, vp, vp, sz, vp, vp]
    lib.zkp_msm_ct_batched.argtypes = [vp, vp, vp, i32, vp, sz, vp]
    lib.zkp_decompress_batch.argtypes = [vp, vp, sz, vp, vp]
    lib.zkp_compress_batch.argtypes = [vp, vp, sz, vp]
    lib.zkp_batch_verify_proofs.argtypes = [vp, P(Desc), vp, sz, vp, vp, vp, vp, vp, P(i32), P(i64), vp, vp]
    lib.zkp_prove_batch.argtypes = [vp, P(Desc), vp, sz, vp, vp, vp, vp, vp, vp, vp]
    lib.zkp_selftest_hash.argtypes = [vp, vp]
    ctx = vp()
    assert lib.zkp_ctx_create(ctypes.byref(ctx), 0) == 0
    lib.ctx = ctx
    # window 12: 22 windows -> 22 blocks of the (slow to emulate) 1024-thread scan per MSM
    assert lib.zkp_ctx_set_option(ctx, b"window", 12) == 0
    yield lib
    lib.zkp_ctx_destroy(ctx)


def _p(a):
    return a.ctypes.data_as(vp)


def _msm(api, sc, pts):
    sc, pts = np.ascontiguousarray(sc, np.uint8).reshape(-1, 32), np.ascontiguousarray(pts, np.uint8).reshape(-1, 32)
    out = np.zeros(32, np.uint8)
    ident, bad = i32(0), i64(-1)
    rc = api.zkp_msm_vartime(api.ctx, _p(sc), _p(pts), sc.shape[0], _p(out), ctypes.byref(ident), ctypes.byref(bad))
    return rc, out.tobytes(), ident.value, bad.value


def test_msm_vartime_through_the_emulated_abi(api):
    """zkp_msm_vartime with the host-input chunk pipeline cut into many chunks (1024 terms each, alternating streams, phase
    boundary inside the batch) and as a single chunk: the bytes of the C port; the empty sum; the index of a bad point."""
    base = U.base_points(64)
    n = 3001
    sc = U.random_scalars(n, seed=8)
    pts = np.frombuffer(b"".join(base[(5 * i) % 64] for i in range(n)), np.uint8).reshape(-1, 32)
    want = cref.msm_vartime(sc, pts)
    assert api.zkp_ctx_set_option(api.ctx, b"chunk_terms", 1024) == 0
    try:
        assert _msm(api, sc, pts) == (0, want, 0, -1)
        if SLOW:
            bad = pts.copy()
            bad[2222] = 0xFF
            assert _msm(api, sc, bad)[0::3] == (native.ZKP_ERR_POINT, 2222)
    finally:
        api.zkp_ctx_set_option(api.ctx, b"chunk_terms", 1 << 19)
    if SLOW:
        assert _msm(api, sc[:700], pts[:700])[1] == cref.msm_vartime(sc[:700], pts[:700])
    assert _msm(api, sc[:0], pts[:0]) == (0, bytes(32), 1, -1)


def test_batch_verify_and_partial_sums_through_the_emulated_abi(api):
    """zkp_batch_verify on a cancelling instance laid out as static ++ row-major instance terms (accept), with one bit
    flipped (reject); the same batch cut into two shards through zkp_batch_verify_partial + zkp_partials_verdict."""
    base = U.base_points(64)
    num_s, rows, batch = 3, 4, 50
    n = num_s + rows * batch
    assert api.zkp_ctx_set_option(api.ctx, b"small_max", 0) == 0      # this test is about the pipeline's launch sequence
    sc = U.random_scalars(n, seed=12)
    pts = np.frombuffer(b"".join(base[(3 * i + 1) % 64] for i in range(n)), np.uint8).reshape(-1, 32).copy()
    # make the last term cancel the others:  sum_{i<n-1} s_i P_i + (l - 1) * R = 0  with R the sum so far
    part = cref.msm_vartime(sc[:n - 1], pts[:n - 1])
    pts[n - 1] = np.frombuffer(part, np.uint8)
    sc[n - 1] = np.frombuffer((S.L - 1).to_bytes(32, "little"), np.uint8)
    acc, bad = i32(-1), i64(0)
    call = lambda s_: api.zkp_batch_verify(api.ctx, _p(s_[:num_s]), _p(pts[:num_s]), num_s, _p(s_[num_s:]), _p(pts[num_s:]), rows,
                                           batch, ctypes.byref(acc), ctypes.byref(bad))
    try:
        assert call(sc) == 0 and acc.value == 1
        flipped = sc.copy()
        flipped[17, 0] ^= 1
        assert call(flipped) == 0 and acc.value == 0
    finally:
        api.zkp_ctx_set_option(api.ctx, b"small_max", 1024)
    if not SLOW:
        return
    api.zkp_ctx_set_option(api.ctx, b"small_max", 0)
    # two shards: columns [0, 20) and [20, 50) of every row; the static terms go to the first shard
    inst_s = sc[num_s:].reshape(rows, batch, 32)
    inst_p = pts[num_s:].reshape(rows, batch, 32)
    partials = np.zeros((2, 20), np.uint64)
    for g, (lo, hi) in enumerate(((0, 20), (20, 50))):
        s_ = np.ascontiguousarray(inst_s[:, lo:hi]).reshape(-1, 32)
        p_ = np.ascontiguousarray(inst_p[:, lo:hi]).reshape(-1, 32)
        ns = num_s if g == 0 else 0
        assert api.zkp_batch_verify_partial(api.ctx, _p(sc[:num_s]), _p(pts[:num_s]), ns, _p(s_), _p(p_), rows, hi - lo,
                                            _p(partials[g]), ctypes.byref(bad)) == 0
    enc = np.zeros(32, np.uint8)
    api.zkp_ctx_set_option(api.ctx, b"small_max", 1024)
    assert api.zkp_partials_verdict(api.ctx, _p(partials), 2, ctypes.byref(acc), _p(enc)) == 0
    assert acc.value == 1 and enc.tobytes() == bytes(32)


def _flat(ost):
    names = ost.instance + ost.common
    lhs = np.array([names.index(l) for l, _ in ost.constraints], dtype=np.int32)
    off = np.cumsum([0] + [len(r) for _, r in ost.constraints]).astype(np.int32)
    ts = np.array([ost.secrets.index(s) for _, r in ost.constraints for s, _ in r], dtype=np.int32)
    tp = np.array([names.index(q) for _, r in ost.constraints for _, q in r], dtype=np.int32)
    d = Desc(len(ost.secrets), len(ost.instance), len(ost.common), len(ost.constraints),
             b"".join(x.encode() + b"\0" for x in names), lhs.ctypes.data, off.ctypes.data, ts.ctypes.data, tp.ctypes.data)
    return d, (lhs, off, ts, tp)


def _prefix(ost, tlabel):
    t = OM.Transcript(tlabel)
    OT.domain_sep(t, ost.label)
    for s_ in ost.secrets:
        OT.append_scalar_var(t, s_.encode())
    st = t.strobe
    state = bytes(st.state)
    return np.array([int.from_bytes(state[4 * i:4 * i + 4], "little") for i in range(50)] + [st.pos, st.pos_begin, st.cur_flags],
                    dtype=np.uint32)


class OneShot:
    def __init__(self, b): self.b = b
    def bytes(self, n): return self.b


def _cmz(N, seed):
    rng = OT.SeededRng(seed)
    ost = OT.CMZ10
    common = {n: R.from_uniform_bytes(rng.bytes(64)) for n in ost.common}
    secs, ptss = [], []
    for _ in range(N):
        sec = {n: int.from_bytes(rng.bytes(64), "little") % S.L for n in ost.secrets}
        Pp, Q = R.from_uniform_bytes(rng.bytes(64)), R.from_uniform_bytes(rng.bytes(64))
        pts = dict(common)
        pts["P"], pts["Q"] = Pp, Q
        for i in range(1, 11):
            pts["C_%d" % i] = M.naive_msm([sec["m_%d" % i], sec["z_%d" % i]], [Pp, pts["A"]])
        pts["V"] = M.naive_msm([sec["m_%d" % i] for i in range(1, 11)] + [sec["minus_z_Q"]],
                               [pts["X_%d" % i] for i in range(1, 11)] + [Q])
        secs.append(sec)
        ptss.append(pts)
    return ost, secs, ptss, [rng.bytes(32) for _ in range(N)]


def _limbs(pt):
    return [((c % R.P) >> (51 * j)) & ((1 << 51) - 1) for c in pt for j in range(5)]


def test_prove_batch_and_verify_from_proof_bytes_through_the_emulated_abi(api):
    """zkp_prove_batch (Straus and comb paths, shared and per-proof tables, slices) gives the oracle prover's proofs for
    CMZ'13; zkp_batch_verify_proofs accepts them, hands out the oracle BatchVerifier's MSM inputs, rejects a tampered
    response, and reports an identity commitment / a non-canonical response as the reference does (VerificationFailure)."""
    N = 3
    ost, secs, ptss, entropy = _cmz(N, b"api-emul")
    names = ost.instance + ost.common
    m, p, k, ni = len(ost.secrets), len(names), len(ost.constraints), len(ost.instance)
    d, keep = _flat(ost)
    prefix = _prefix(ost, b"CMZ")
    want = [ost.prove_batchable(OM.Transcript(b"CMZ"), secs[j], ptss[j], OneShot(entropy[j])) for j in range(N)]
    sec = np.frombuffer(b"".join(S.to_bytes(s[n]) for s in secs for n in ost.secrets), np.uint8).copy()
    lim = np.array([[_limbs(pp[n]) for n in names] for pp in ptss], dtype=np.uint64)
    ent = np.frombuffer(b"".join(entropy), np.uint8).copy()

    def prove():
        enc, com, resp = np.zeros((N, p, 32), np.uint8), np.zeros((N, k, 32), np.uint8), np.zeros((N, m, 32), np.uint8)
        rc = api.zkp_prove_batch(api.ctx, ctypes.byref(d), _p(prefix), N, _p(sec), _p(lim), _p(ent), _p(enc), _p(com), _p(resp), None)
        return rc, enc, com, resp

    try:
        # prove_comb: 0 = Straus tables, 1 = combs scanned from global memory, 2 = combs staged in shared memory (default)
        # pipe = 1: slices of one proof alternating between the two workspaces / streams of the pipelined prover
        for comb, share, chunk, pipe in ((0, 1, 1 << 17, 0), (1, 1, 1 << 17, 0), (1, 0, 1 << 17, 0), (1, 1, 2, 0), (0, 0, 2, 0),
                                         (2, 1, 1 << 17, 0), (2, 0, 2, 0), (2, 1, 1 << 17, 1)):
            for key, v in ((b"prove_comb", comb), (b"share_static_tables", share), (b"prove_chunk", chunk),
                           (b"prove_pipe_chunk", pipe)):
                assert api.zkp_ctx_set_option(api.ctx, key, v) == 0
            rc, enc, com, resp = prove()
            assert rc == 0, (comb, share, chunk)
            for j in range(N):
                proof, oenc = want[j]
                assert [bytes(c) for c in com[j]] == proof.commitments, (comb, share, chunk, j)
                assert [bytes(r) for r in resp[j]] == [S.to_bytes(r) for r in proof.responses]
                assert [bytes(e) for e in enc[j]] == [oenc[n] for n in names]
    finally:
        for key, v in ((b"prove_comb", 2), (b"share_static_tables", 1), (b"prove_chunk", 1 << 17), (b"prove_pipe_chunk", 1 << 14)):
            api.zkp_ctx_set_option(api.ctx, key, v)
    # ---- verification of those proofs from their bytes ----
    seed = bytes(range(32))
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    comm = np.ascontiguousarray(enc[0, ni:])
    n = len(ost.common) + (ni + k) * N
    co, po = np.zeros((n, 32), np.uint8), np.zeros((n, 32), np.uint8)
    acc, bad = i32(-1), i64(0)
    sd = (ctypes.c_uint8 * 32)(*seed)

    def verify(com_, resp_, want_inputs=False):
        return api.zkp_batch_verify_proofs(api.ctx, ctypes.byref(d), _p(prefix), N, _p(inst), _p(comm), _p(com_), _p(resp_), sd,
                                           ctypes.byref(acc), ctypes.byref(bad), _p(co) if want_inputs else None,
                                           _p(po) if want_inputs else None)

    assert verify(com, resp, True) == 0 and acc.value == 1
    oencs = {nm: [bytes(enc[j, i]) for j in range(N)] for i, nm in enumerate(ost.instance)}
    for i, nm in enumerate(ost.common):
        oencs[nm] = bytes(comm[i])
    oproofs = [OT.BatchableProof([bytes(c) for c in com[j]], [int.from_bytes(bytes(r), "little") for r in resp[j]]) for j in range(N)]
    bv = ost.build_batch_verifier(N, [OM.Transcript(b"CMZ") for _ in range(N)], oencs)
    oscal, opts = bv.batch_coeffs(oproofs, OT.PerProofRng(seed))
    assert [bytes(x) for x in po] == opts and [bytes(x) for x in co] == [S.to_bytes(s) for s in oscal]
    tampered = resp.copy()
    tampered[1, 4, 0] ^= 1
    assert verify(com, tampered) == 0 and acc.value == 0
    if not SLOW:
        return
    ident = com.copy()
    ident[2, 3] = 0
    assert verify(ident, resp) == native.ZKP_ERR_POINT
    noncanon = resp.copy()
    noncanon[0, 2] = 0xFF
    assert verify(com, noncanon) == native.ZKP_ERR_SCALAR


def test_small_msms_codec_and_selftests_through_the_emulated_abi(api):
    """zkp_msm_vartime_batched / zkp_msm_ct_batched (both schedules: one thread and four lanes per MSM) over the golden KATs,
    zkp_decompress_batch / zkp_compress_batch round trip with an invalid encoding, the device Merlin conformance vector."""
    kats = [k_ for k_ in U.golden("msm_kat.json")["kats"] if k_["n"] <= 36]
    sc = np.frombuffer(b"".join(bytes.fromhex(x) for k_ in kats for x in k_["scalars"]), np.uint8).reshape(-1, 32)
    pt = np.frombuffer(b"".join(bytes.fromhex(x) for k_ in kats for x in k_["points"]), np.uint8).reshape(-1, 32)
    off = np.cumsum([0] + [k_["n"] for k_ in kats]).astype(np.uint64)
    M_ = len(kats)
    expected = b"".join(bytes.fromhex(k_["expected"]) for k_ in kats)
    try:
        for coop_max in (8192, 0):                     # four lanes per MSM / one thread per MSM
            assert api.zkp_ctx_set_option(api.ctx, b"coop_max_msms", coop_max) == 0
            out, valid = np.zeros((M_, 32), np.uint8), np.zeros(M_, np.uint8)
            assert api.zkp_msm_vartime_batched(api.ctx, _p(sc), _p(pt), _p(off), M_, _p(out), _p(valid)) == 0
            assert valid.all() and out.tobytes() == expected, coop_max
            out2 = np.zeros((M_, 32), np.uint8)
            assert api.zkp_msm_ct_batched(api.ctx, _p(sc), _p(pt), native.ZKP_POINTS_COMPRESSED, _p(off), M_, _p(out2)) == 0
            assert out2.tobytes() == expected, coop_max
    finally:
        api.zkp_ctx_set_option(api.ctx, b"coop_max_msms", 8192)
    # one small MSM through zkp_msm_vartime: the four-lane Straus path (k_single_msm_vt / k_single_finish) instead of the
    # sort pipeline -- one term per group, and several terms per group; the index of a bad point
    try:
        for groups in (1024, 3):
            assert api.zkp_ctx_set_option(api.ctx, b"small_groups", groups) == 0
            for k_ in kats[:4] + kats[-1:]:
                s_ = np.frombuffer(b"".join(bytes.fromhex(x) for x in k_["scalars"]), np.uint8).reshape(-1, 32)
                p_ = np.frombuffer(b"".join(bytes.fromhex(x) for x in k_["points"]), np.uint8).reshape(-1, 32)
                rc, enc, ident, bad = _msm(api, s_, p_)
                assert (rc, enc.hex(), bad) == (0, k_["expected"], -1), (groups, k_["n"])
                assert ident == (k_["expected"] == "00" * 32)
        p_bad = p_.copy()
        p_bad[len(p_bad) // 2] = 0xFF
        assert _msm(api, s_, p_bad)[0::3] == (native.ZKP_ERR_POINT, len(p_bad) // 2)
    finally:
        api.zkp_ctx_set_option(api.ctx, b"small_groups", 1024)
    encs = np.frombuffer(b"".join(U.base_points(8)) + b"\xff" * 32, np.uint8).reshape(-1, 32)
    limbs, valid = np.zeros((9, 20), np.uint64), np.zeros(9, np.uint8)
    assert api.zkp_decompress_batch(api.ctx, _p(encs), 9, _p(limbs), _p(valid)) == 0
    assert list(valid) == [1] * 8 + [0]
    back = np.zeros((8, 32), np.uint8)
    assert api.zkp_compress_batch(api.ctx, _p(limbs), 8, _p(back)) == 0
    assert back.tobytes() == encs[:8].tobytes()
    buf = ctypes.create_string_buffer(32)
    assert api.zkp_selftest_hash(api.ctx, buf) == 0
    assert buf.raw.hex() == U.golden("merlin.json")["complex"]


def test_verify_from_proof_bytes_over_several_slabs_through_the_emulated_abi(api):
    """zkp_batch_verify_proofs with more proofs than one slab holds (1030 DLEQ proofs, 1024-proof slabs): the copies run
    ahead, the front-end kernel runs slab by slab on its own stream as a resident grid (2 emulated SMs x 2 blocks: every block
    loops over two 128-proof groups), the split of the rows between the two ingestion phases is chosen per slab -- here the
    stand-in runtime reports every earlier slab as finished, so slab 1 is decompressed whole in phase 1 -- or alternates
    (-1) or is fixed, phase 2 runs once per run of equally split slabs.  Every variant builds the MSM inputs of the
    single-slab call, accepts, and rejects a tampered proof of the LAST slab."""
    ost = OT.DLEQ
    G = R.BASEPOINT
    H = R.hash_from_bytes_sha512(R.compress(G))
    rng = OT.SeededRng(b"slabs")
    D, N = 5, 1030
    proofs, encs = [], []
    for j in range(D):
        x = 7700000001 + 13 * j
        pts = {"A": R.pt_mul(x, G), "B": R.pt_mul(x, H), "H": H, "G": G}
        proof, enc = ost.prove_batchable(OM.Transcript(b"DLEQSlabs"), {"x": x}, pts, OneShot(rng.bytes(32)))
        proofs.append(proof)
        encs.append(enc)
    names = ost.instance + ost.common
    ni, nc, k, m = len(ost.instance), len(ost.common), len(ost.constraints), len(ost.secrets)
    d, keep = _flat(ost)
    prefix = _prefix(ost, b"DLEQSlabs")
    inst = np.frombuffer(b"".join(encs[j % D][n] for n in ost.instance for j in range(N)), np.uint8).reshape(ni, N, 32).copy()
    comm = np.frombuffer(b"".join(encs[0][n] for n in ost.common), np.uint8).reshape(nc, 32).copy()
    com = np.frombuffer(b"".join(c for j in range(N) for c in proofs[j % D].commitments), np.uint8).reshape(N, k, 32).copy()
    resp = np.frombuffer(b"".join(S.to_bytes(r) for j in range(N) for r in proofs[j % D].responses), np.uint8).reshape(N, m, 32).copy()
    n = nc + (ni + k) * N
    sd = (ctypes.c_uint8 * 32)(*bytes(range(3, 35)))
    acc, bad = i32(-1), i64(0)

    def verify(resp_, want=True):
        co, po = np.zeros((n, 32), np.uint8), np.zeros((n, 32), np.uint8)
        rc = api.zkp_batch_verify_proofs(api.ctx, ctypes.byref(d), _p(prefix), N, _p(inst), _p(comm), _p(com), _p(resp_), sd,
                                         ctypes.byref(acc), ctypes.byref(bad), _p(co) if want else None, _p(po) if want else None)
        return rc, acc.value, co, po

    tampered = resp.copy()
    tampered[N - 2, 0, 5] ^= 0x40
    try:
        assert api.zkp_ctx_set_option(api.ctx, b"bv_chunk_terms", 1 << 30) == 0     # one slab
        rc, a, co0, po0 = verify(resp)
        assert (rc, a) == (0, 1)
        assert api.zkp_ctx_set_option(api.ctx, b"bv_chunk_terms", 1024) == 0        # slabs of 1024 + 6 proofs
        cases = (0, -1, 1, 3, 4) if SLOW else (0, -1, 3)
        for rows1 in cases:
            assert api.zkp_ctx_set_option(api.ctx, b"bv_phase1_rows", rows1) == 0
            rc, a, co, po = verify(resp)
            assert (rc, a) == (0, 1), rows1
            assert (co == co0).all() and (po == po0).all(), rows1
            rc, a, _, _ = verify(tampered, want=False)
            assert (rc, a) == (0, 0), rows1
    finally:
        for key, v in ((b"bv_chunk_terms", 3 << 17), (b"bv_phase1_rows", 0)):
            api.zkp_ctx_set_option(api.ctx, key, v)
