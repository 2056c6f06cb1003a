"""GPU parity tests: the CUDA path through the C ABI against the oracle / golden fixtures (bit exact)."""
import numpy as np
import pytest

from oracle import msm as M, ristretto as R
from tests import util_data as U

pytestmark = pytest.mark.gpu


def _ablations(engine):
    """True if the library was built with -DZKP_ABLATIONS (python -m zkp_b200.build --ablations): the measured-negative
    paths (second-stream sort, ramped chunk schedule, byte-wise STROBE front end) are not in the product library."""
    from zkp_b200 import EngineError
    try:
        engine.set_option("overlap", 1)
    except EngineError:
        return False
    engine.set_option("overlap", 0)
    return True


def _h(xs):
    return np.frombuffer(b"".join(bytes.fromhex(x) for x in xs), dtype=np.uint8).reshape(-1, 32)


def test_rfc9496_multiples_via_msm(engine):
    g = U.golden("rfc9496.json")
    B = bytes.fromhex(g["multiples_of_generator"][1])
    for k, exp in enumerate(g["multiples_of_generator"]):
        enc, ident, _ = engine.msm_vartime(k.to_bytes(32, "little"), B)
        assert enc.hex() == exp
        assert ident == (k == 0)


def test_decompress_compress_roundtrip_and_bad_encodings(engine):
    g = U.golden("rfc9496.json")
    good = [bytes.fromhex(x) for x in g["multiples_of_generator"]] + U.base_points(40)
    bad = [bytes.fromhex(x) for x in g["bad_encodings"]]
    limbs, valid = engine.decompress_batch(good + bad)
    assert valid[:len(good)].all() and not valid[len(good):].any()
    back = engine.compress_batch(limbs[:len(good)])
    assert [bytes(b) for b in back] == good
    # limbs agree with the oracle's affine coordinates
    for i, e in enumerate(good[:20]):
        x, y, z, t = R.decompress(e)
        got = [sum(int(limbs[i, c, j]) << (51 * j) for j in range(5)) for c in range(4)]
        assert got == [x, y, 1, t]


def test_msm_kats(engine):
    for kat in U.golden("msm_kat.json")["kats"]:
        enc, ident, _ = engine.msm_vartime(_h(kat["scalars"]), _h(kat["points"]))
        assert enc.hex() == kat["expected"], kat["n"]
        assert ident == (kat["expected"] == "00" * 32)


def test_small_n_dispatch_matches_the_pipeline(engine):
    """Few terms take the four-lane Straus path (k_single_msm_vt: the reference's n < 190 dispatch,
    /root/reference/src/toolbox/verifier.rs:162-166), many take the Pippenger pipeline; every KAT gives the same bytes
    through both (small_max = 0 forces the pipeline, a large small_max forces the small path, small_groups cuts the terms
    over fewer groups), and the status / first-bad-index reporting is the same."""
    kats = U.golden("msm_kat.json")["kats"]
    g = U.golden("rfc9496.json")
    bad_pt = np.frombuffer(bytes.fromhex(g["bad_encodings"][5]), dtype=np.uint8)
    from zkp_b200 import EngineError
    try:
        for small_max, groups in ((0, 1024), (1 << 20, 1024), (1 << 20, 7), (1 << 20, 1)):
            engine.set_option("small_max", small_max)
            engine.set_option("small_groups", groups)
            for kat in kats:
                enc, ident, _ = engine.msm_vartime(_h(kat["scalars"]), _h(kat["points"]))
                assert enc.hex() == kat["expected"], (small_max, groups, kat["n"])
                assert ident == (kat["expected"] == "00" * 32)
            kat = [k for k in kats if k["n"] >= 36][0]
            sc, pts = _h(kat["scalars"]).copy(), _h(kat["points"]).copy()
            for idx in (0, 17, kat["n"] - 1):
                p2 = pts.copy()
                p2[idx] = bad_pt
                enc, _, first_bad = engine.msm_vartime(sc, p2)
                assert enc is None and first_bad == idx, (small_max, groups, idx)
                s2 = sc.copy()
                s2[idx] = 0xFF
                with pytest.raises(EngineError) as ei:
                    engine.msm_vartime(s2, pts)
                assert ei.value.code == 3
            # the batch entry point and the partial-sum entry point (single-verdict mode) over a small batch
            kb = U.golden("toolbox_kat.json")["dleq_batch"]
            scal, bpts = _h(kb["msm_scalars"]), _h(kb["msm_points"])
            rows = (scal.shape[0] - 1) // 4
            ok, rc = engine.batch_verify(scal[:1], bpts[:1], scal[1:], bpts[1:], rows, 4)
            assert ok and rc == 0, (small_max, groups)
            part = engine.batch_verify_partial(scal[:1], bpts[:1], scal[1:], bpts[1:], rows, 4)
            acc, enc = engine.partials_verdict(part.reshape(1, 20))
            assert acc and enc == bytes(32)
    finally:
        engine.set_option("small_max", 1024)
        engine.set_option("small_groups", 1024)


@pytest.mark.parametrize("window", [4, 5, 7, 8, 11, 13, 16])
def test_msm_all_windows(engine, window):
    kat = [k for k in U.golden("msm_kat.json")["kats"] if k["n"] == 300][0]
    engine.set_option("window", window)
    try:
        for lanes in (1, 4, 32):
            engine.set_option("lanes", lanes)
            enc, _, _ = engine.msm_vartime(_h(kat["scalars"]), _h(kat["points"]))
            assert enc.hex() == kat["expected"], (window, lanes)
    finally:
        engine.set_option("window", 0)
        engine.set_option("lanes", 0)


def test_msm_seeded_large(engine):
    for case in U.golden("msm_seeded.json")["cases"]:
        base = U.base_points(case["K"])
        sc = U.random_scalars(case["n"], seed=case["seed"])
        pts = np.frombuffer(b"".join(base[i % case["K"]] for i in range(case["n"])), dtype=np.uint8).reshape(-1, 32)
        enc, _, _ = engine.msm_vartime(sc, pts)
        assert enc.hex() == case["expected"]


def test_msm_edge_cases(engine):
    pts = U.base_points(8)
    # empty
    enc, ident, _ = engine.msm_vartime(b"", b"")
    assert enc == bytes(32) and ident
    # invalid point -> None with index
    g = U.golden("rfc9496.json")
    bad = bytes.fromhex(g["bad_encodings"][5])
    sc = U.random_scalars(8, seed=3)
    p = list(pts)
    p[5] = bad
    enc, ident, first_bad = engine.msm_vartime(sc, p)
    assert enc is None and first_bad == 5
    # all scalars equal (one bucket per window takes everything), and all-zero scalars
    n = 777
    one = np.tile(np.frombuffer((R.L - 5).to_bytes(32, "little"), dtype=np.uint8), (n, 1))
    P = [pts[i % 8] for i in range(n)]
    exp = R.compress(M.naive_msm([(R.L - 5) * (n // 8 + (1 if i < n % 8 else 0)) for i in range(8)],
                                 [R.decompress(e) for e in pts]))
    enc, _, _ = engine.msm_vartime(one, P)
    assert enc == exp
    enc, ident, _ = engine.msm_vartime(np.zeros((n, 32), dtype=np.uint8), P)
    assert enc == bytes(32) and ident
    # non-canonical scalar is rejected
    from zkp_b200 import EngineError
    s2 = sc.copy()
    s2[2] = np.frombuffer(R.L.to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(EngineError) as ei:
        engine.msm_vartime(s2, pts)
    assert ei.value.code == 3


def test_msm_linearity_full_size(engine):
    """size-independent property at a BASELINE-sized input: MSM(a*s, P) == a * MSM(s, P) via tiling."""
    n, K = 1 << 18, 256
    base = U.base_points(K)
    sc = U.random_scalars(n, seed=99)
    pts = np.frombuffer(b"".join(base[i % K] for i in range(n)), dtype=np.uint8).reshape(-1, 32)
    enc, _, _ = engine.msm_vartime(sc, pts)
    assert enc == U.tiled_expected(sc, base)


def test_msm_bench_scale_properties(engine):
    """At the size class of the bench (2^22 terms: window 17+, two-phase ingestion over two chunks, multi-block item scan,
    length-sorted accumulation) the oracle is too slow, so size-independent properties decide: the tiling identity against
    a 256-term oracle MSM, additivity MSM(all) = MSM(first part) + MSM(rest) with the sum taken by the engine itself, a
    cancelling instance (s and l - s on the same point) that must give the identity, and one flipped bit that must not."""
    n, K = 1 << 22, 256
    base = U.base_points(K)
    sc = U.random_scalars(n, seed=4242)
    pts = np.frombuffer(b"".join(base[i % K] for i in range(K)) * (n // K), dtype=np.uint8).reshape(-1, 32)
    enc, ident, _ = engine.msm_vartime(sc, pts)
    assert enc == U.tiled_expected(sc, base) and not ident
    cut = 1234567
    e1, _, _ = engine.msm_vartime(sc[:cut], pts[:cut])
    e2, _, _ = engine.msm_vartime(sc[cut:], pts[cut:])
    one = np.zeros((2, 32), np.uint8)
    one[:, 0] = 1
    esum, _, _ = engine.msm_vartime(one, np.frombuffer(e1 + e2, dtype=np.uint8).reshape(2, 32))
    assert esum == enc
    # cancelling pairs: term i and term i + n/2 carry s and l - s on the same point
    half = n // 2
    L = R.L
    s_lo = sc[:half].copy()
    lw = np.frombuffer(L.to_bytes(32, "little"), dtype=np.uint64)
    a = s_lo.view(np.uint64).reshape(half, 4)
    neg = np.empty_like(a)
    borrow = np.zeros(half, dtype=np.uint64)
    for w in range(4):
        d = lw[w] - a[:, w]
        b1 = (lw[w] < a[:, w]).astype(np.uint64)
        d2 = d - borrow
        b2 = (d < borrow).astype(np.uint64)
        neg[:, w] = d2
        borrow = b1 | b2
    zero = (a == 0).all(axis=1)
    neg[zero] = 0
    both = np.concatenate([s_lo, neg.view(np.uint8).reshape(half, 32)])
    enc0, ident0, _ = engine.msm_vartime(both, pts)          # pts[i + half] == pts[i] because half is a multiple of K
    assert ident0 and enc0 == bytes(32)
    both[777, 0] ^= 1
    enc1, ident1, _ = engine.msm_vartime(both, pts)
    assert not ident1


_HEADLINE = {}


def _headline_case(n, K=256):
    """n terms with the scalar mix of the bench (bench.make_cmz_batch: 13 of 24 rows full-size, 11 of 24 rows l - rho with
    128-bit rho, i.e. what batch_verifier.rs:183 produces and the engine's sign fold turns back into 128-bit digits) over
    K tiled base points, and the exact expected encoding from the K-term reduction."""
    if n not in _HEADLINE:
        import bench
        rng = np.random.default_rng(20260 + n % 1000)
        sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        sc[:, 31] &= 0x0F                                         # < 2^252 < l
        small = (np.arange(n) % 24) >= 13                        # the commitment rows of a CMZ batch
        rho = sc[small].copy()
        rho[:, 16:] = 0
        sc[small] = bench._sub_from_l(rho.view(np.uint64).reshape(-1, 4)).view(np.uint8).reshape(-1, 32)
        base = U.base_points(K)
        reps = -(-n // K)
        pts = np.frombuffer(b"".join(base) * reps, dtype=np.uint8).reshape(-1, 32)[:n]
        _HEADLINE[n] = (sc, pts, U.tiled_expected(sc, base))
    return _HEADLINE[n]


@pytest.mark.parametrize("window", [17, 18, 19, 20, 22, 24])
def test_msm_headline_windows_bit_exact(engine, window):
    """The windows the bench runs at (c = 19 by the automatic choice at 5 * 10^7 terms; 17 / 18 at smaller sizes; 20..24 by
    option) checked bit for bit at n = 2^22 + 12 terms against the tiling oracle: B up to 2^23 buckets per window, the
    multi-block item scan with > 256 tiles, bucket-reduction depth 4, the sign-folded 128-bit coefficient mix of a CMZ
    batch (/root/reference/src/toolbox/batch_verifier.rs:173-230)."""
    sc, pts, exp = _headline_case((1 << 22) + 12)
    engine.set_option("window", window)
    try:
        enc, ident, _ = engine.msm_vartime(sc, pts)
        assert engine.stage_ms()["window"] == window
    finally:
        engine.set_option("window", 0)
    assert enc == exp and not ident, window


def test_msm_2_pow_24_terms_at_the_automatic_window(engine):
    """One 2^24-term MSM (a third of the bench's per-GPU MSM) at the window the engine picks itself, bit for bit against
    the tiling oracle; and through the batch entry point with 12 static terms in front (the bench's e2e call)."""
    n = 1 << 24
    sc, pts, exp = _headline_case(n)
    enc, ident, _ = engine.msm_vartime(sc, pts)
    assert engine.stage_ms()["window"] >= 17
    assert enc == exp and not ident
    # zkp_batch_verify: same terms as 12 static + 24 rows x batch (the result is not the identity: reject, status ok)
    batch = (n - 12) // 24
    m = 12 + 24 * batch
    ok, rc = engine.batch_verify(sc[:12], pts[:12], sc[12:m], pts[12:m], 24, batch)
    assert not ok and rc == 0


def test_small_batched_vartime(engine):
    kats = [k for k in U.golden("msm_kat.json")["kats"] if k["n"] <= 36]
    scal = np.concatenate([_h(k["scalars"]) for k in kats])
    pts = np.concatenate([_h(k["points"]) for k in kats])
    off = np.cumsum([0] + [k["n"] for k in kats]).astype(np.uint64)
    out, valid = engine.msm_vartime_batched(scal, pts, off)
    assert valid.all()
    assert [bytes(o).hex() for o in out] == [k["expected"] for k in kats]
    # one thread per MSM (throughput schedule) and four lanes per MSM (latency schedule) give the same bytes
    for coop in (0, 1 << 20):
        engine.set_option("coop_max_msms", coop)
        try:
            o2, v2 = engine.msm_vartime_batched(scal, pts, off)
        finally:
            engine.set_option("coop_max_msms", 8192)
        assert v2.all() and (o2 == out).all(), coop
    # one bad point invalidates only its MSM
    g = U.golden("rfc9496.json")
    pts2 = pts.copy()
    pts2[int(off[3])] = np.frombuffer(bytes.fromhex(g["bad_encodings"][0]), dtype=np.uint8)
    out2, valid2 = engine.msm_vartime_batched(scal, pts2, off)
    assert not valid2[3] and valid2.sum() == len(kats) - 1
    assert bytes(out2[2]).hex() == kats[2]["expected"]


def test_small_batched_ct(engine):
    kats = [k for k in U.golden("msm_kat.json")["kats"] if k["n"] <= 36]
    scal = np.concatenate([_h(k["scalars"]) for k in kats])
    pts = np.concatenate([_h(k["points"]) for k in kats])
    off = np.cumsum([0] + [k["n"] for k in kats]).astype(np.uint64)
    out = engine.msm_ct_batched(scal, pts, off)
    assert [bytes(o).hex() for o in out] == [k["expected"] for k in kats]
    limbs, valid = engine.decompress_batch(pts)
    assert valid.all()
    out = engine.msm_ct_batched(scal, limbs, off, limbs=True)
    assert [bytes(o).hex() for o in out] == [k["expected"] for k in kats]
    # throughput schedule (one thread per MSM) and latency schedule (four lanes per MSM): same bytes
    for coop in (0, 1 << 20):
        engine.set_option("coop_max_msms", coop)
        try:
            o2 = engine.msm_ct_batched(scal, pts, off)
        finally:
            engine.set_option("coop_max_msms", 8192)
        assert [bytes(o).hex() for o in o2] == [k["expected"] for k in kats], coop


def test_batch_verify_dleq_golden(engine):
    kb = U.golden("toolbox_kat.json")["dleq_batch"]
    scal, pts = _h(kb["msm_scalars"]), _h(kb["msm_points"])
    num_s, batch = 1, 4
    rows = (scal.shape[0] - num_s) // batch
    ok, rc = engine.batch_verify(scal[:num_s], pts[:num_s], scal[num_s:], pts[num_s:], rows, batch)
    assert ok and rc == 0
    bad = scal.copy()
    bad[3, 0] ^= 1
    ok, rc = engine.batch_verify(bad[:num_s], pts[:num_s], bad[num_s:], pts[num_s:], rows, batch)
    assert not ok and rc == 0


def test_host_pipeline_chunking_overlap_and_profile_options(engine):
    """The chunked H2D pipeline (several chunks, ragged tail), the second-stream sort and the profiling mode all
    give the same bytes."""
    case = U.golden("msm_seeded.json")["cases"][1]
    base = U.base_points(case["K"])
    sc = U.random_scalars(case["n"], seed=case["seed"])
    pts = np.frombuffer(b"".join(base[i % case["K"]] for i in range(case["n"])), dtype=np.uint8).reshape(-1, 32)
    try:
        cases = [{"chunk_terms": 4096}, {"profile": 1}, {"chunk_terms": 1 << 16}]
        if _ablations(engine):
            cases += [{"chunk_terms": 4096, "overlap": 1}, {"overlap": 1}]
        for opts in cases:
            for k, v in opts.items():
                engine.set_option(k, v)
            enc, _, _ = engine.msm_vartime(sc, pts)
            assert enc.hex() == case["expected"], opts
            if "profile" in opts:
                st = engine.stage_ms()
                assert st["decompress"] > 0 and st["accumulate"] > 0 and 4 <= st["window"] <= 24
            engine.set_option("profile", 0)
            engine.set_option("overlap", 0)
            engine.set_option("chunk_terms", 1 << 19)
    finally:
        engine.set_option("profile", 0)
        engine.set_option("overlap", 0)
        engine.set_option("chunk_terms", 1 << 19)


def test_two_phase_ingestion_and_item_balance_options(engine):
    """The two-phase ingestion (histogram / scatter riding under the two halves of the decompression: k_ingest2) and the
    size-ordered accumulation give the same bytes as the separate passes, for odd and even chunk counts, a single
    chunk, odd term counts, and through the device-resident entry point (zkp_msm_vartime_dev)."""
    import torch
    case = U.golden("msm_seeded.json")["cases"][1]
    base = U.base_points(case["K"])
    sc = U.random_scalars(case["n"], seed=case["seed"])
    pts = np.frombuffer(b"".join(base[i % case["K"]] for i in range(case["n"])), dtype=np.uint8).reshape(-1, 32)
    n = case["n"]
    from oracle import msm as M
    odd = 4097
    exp_odd = M.msm_bytes([sc[i].tobytes() for i in range(odd)], [pts[i].tobytes() for i in range(odd)])
    try:
        for fused in (1, 0):
            for balance in (1, 0):
                engine.set_option("fused_sort", fused)
                engine.set_option("balance", balance)
                for chunk in (1 << 21, 4096, 5000, n // 3 + 1, n):       # 1 chunk, many, ragged, 3 chunks (odd), exactly 1
                    engine.set_option("chunk_terms", max(chunk, 1024))
                    # uniform chunks on one / two streams (and the ramped schedule where the ablations are built in)
                    for ramp, dual in (((1, 0), (0, 0), (0, 1), (1, 1)) if _ablations(engine) else ((0, 0), (0, 1))):
                        engine.set_option("ramp_chunks", ramp)
                        engine.set_option("dual_stream", dual)
                        enc, _, _ = engine.msm_vartime(sc, pts)
                        assert enc.hex() == case["expected"], (fused, balance, chunk, ramp, dual)
                        for pct in (1, 30, 100):                         # phase boundary at either end and off-centre
                            engine.set_option("phase1_percent", pct)
                            enc, _, _ = engine.msm_vartime(sc, pts)
                            assert enc.hex() == case["expected"], (fused, balance, chunk, ramp, dual, pct)
                        engine.set_option("phase1_percent", 50)
                    engine.set_option("ramp_chunks", 0)
                    engine.set_option("dual_stream", 1)
                engine.set_option("chunk_terms", 1024)
                enc, _, _ = engine.msm_vartime(sc[:odd], pts[:odd])
                assert enc == exp_odd, (fused, balance)
                # device-resident entry point
                d_sc = torch.from_numpy(sc).cuda()
                d_pt = torch.from_numpy(pts.copy()).cuda()
                d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
                for cnt, exp in ((1, None), (odd, exp_odd), (n, bytes.fromhex(case["expected"]))):
                    engine.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), cnt, d_res.data_ptr())
                    engine.synchronize()
                    r = d_res.cpu().numpy()
                    assert int(np.frombuffer(r[32:36].tobytes(), dtype=np.int32)[0]) == 0
                    if exp is not None:
                        assert r[:32].tobytes() == exp, (fused, balance, cnt)
                    else:
                        assert r[:32].tobytes() == M.msm_bytes([sc[0].tobytes()], [pts[0].tobytes()])
                if fused:
                    live = engine.live_ms()
                    assert live["ingest_phase1"] > 0 and live["ingest_phase2"] > 0 and live["accumulate"] > 0
        # an invalid point / a non-canonical scalar in either half is reported with its index by the fused path
        engine.set_option("fused_sort", 1)
        engine.set_option("chunk_terms", 4096)
        from zkp_b200 import EngineError
        for dual in (0, 1):
            engine.set_option("dual_stream", dual)
            for idx in (3, n - 2):
                bad = pts.copy()
                bad[idx] = 0xFF
                enc, _, first_bad = engine.msm_vartime(sc, bad)
                assert enc is None and first_bad == idx
                bs = sc.copy()
                bs[idx] = 0xFF
                with pytest.raises(EngineError) as ei:
                    engine.msm_vartime(bs, pts)
                assert ei.value.code == 3
    finally:
        engine.set_option("fused_sort", 1)
        engine.set_option("balance", 1)
        engine.set_option("ramp_chunks", 0)
        engine.set_option("dual_stream", 1)
        engine.set_option("phase1_percent", 50)
        engine.set_option("chunk_terms", 1 << 19)


def test_maximum_sizes_are_refused_not_attempted(engine):
    """Term indices travel in 31 bits of the sorted entries, so one call takes fewer than 2^31 - 1 terms: larger requests
    come back as ZKP_ERR_SIZE before any allocation or launch (device-resident and batch entry points), and the batch
    entry points refuse rows * batch products that overflow."""
    import ctypes
    import torch
    from zkp_b200 import native
    lib, ctx = engine._lib, engine._ctx
    d = torch.zeros(256, dtype=torch.uint8, device="cuda")
    before = engine.launch_count
    for n in (2**31 - 1, 2**31, 2**40):
        rc = lib.zkp_msm_vartime_dev(ctx, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(d.data_ptr()), n,
                                     ctypes.c_void_p(d.data_ptr()))
        assert rc == native.ZKP_ERR_SIZE, n
    acc = ctypes.c_int32(7)
    buf = np.zeros((4, 32), np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.zkp_batch_verify(ctx, None, None, 0, p(buf), p(buf), 2**30, 2**30, ctypes.byref(acc), None) == native.ZKP_ERR_SIZE
    assert lib.zkp_batch_verify(ctx, None, None, 0, p(buf), p(buf), 2**62, 8, ctypes.byref(acc), None) == native.ZKP_ERR_SIZE
    assert lib.zkp_batch_verify(ctx, None, None, 0, p(buf), p(buf), 2**16, 2**16, ctypes.byref(acc), None) == native.ZKP_ERR_SIZE
    assert acc.value == 0
    assert engine.launch_count == before
    # the engine still works afterwards
    enc, ident, _ = engine.msm_vartime(b"", b"")
    assert enc == bytes(32) and ident


def test_abi_misuse_is_reported_not_crashed(engine):
    import ctypes
    from zkp_b200 import native
    lib, ctx = engine._lib, engine._ctx
    out = (ctypes.c_uint8 * 32)()
    assert lib.zkp_msm_vartime(ctx, None, None, 5, out, None, None) == native.ZKP_ERR_SIZE
    assert lib.zkp_msm_vartime(None, None, None, 0, out, None, None) == native.ZKP_ERR_SIZE
    assert lib.zkp_ctx_set_option(ctx, b"window", 99) == native.ZKP_ERR_SIZE
    assert lib.zkp_ctx_set_option(ctx, b"no-such-option", 1) == native.ZKP_ERR_SIZE
    assert lib.zkp_msm_vartime_dev(ctx, ctypes.c_void_p(8), ctypes.c_void_p(16), 4, ctypes.c_void_p(32)) == native.ZKP_ERR_SIZE
    bad_off = np.array([0, 5, 3], dtype=np.uint64)      # decreasing offsets
    s = np.zeros((5, 32), np.uint8)
    o = np.zeros((2, 32), np.uint8)
    v = np.zeros(2, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.zkp_msm_vartime_batched(ctx, p(s), p(s), p(bad_off), 2, p(o), p(v)) == native.ZKP_ERR_SIZE
    acc = ctypes.c_int32(7)
    assert lib.zkp_batch_verify(ctx, None, None, 0, None, None, 0, 0, ctypes.byref(acc), None) == 0 and acc.value == 1
