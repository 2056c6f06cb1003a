"""The legs of bench.py beyond the headline line: the other BASELINE.json configs at the current HEAD, run by every rank
under torchrun (weak scaling unless said otherwise), reported in the `configs` block of the one JSON line.

  cmz_prove          configs[1]: zkp_prove_batch over 2^16 CMZ'13 proofs per GPU (/root/reference/src/toolbox/prover.rs:76-112)
  dleq_batch_verify  configs[2]: BatchVerifier::verify_batchable over 2^20 DLEQ proofs per GPU, one 2 + 4 * 2^20 term MSM
                     (/root/reference/benches/dleq.rs:188-241): device-resident, from host coefficient buffers, from proofs
  raw_msm_sweep      configs[4]: ONE MSM of 2^8 .. 2^22 random terms; at N > 1 it is cut over the ranks (strong scaling): every
                     rank sums its slice, the 160-byte partial sums cross NCCL in one all-gather on the compute stream,
                     every rank adds them (zkp_msm_vartime_partial_dev / zkp_partials_verdict_dev)
  single_verdict     ONE CMZ batch with ONE accept bit cut over the ranks (exact semantics of batch_verifier.rs:219-234):
                     the shards alone are not the identity, their sum is; a tampered shard voids the verdict
  h2d_ceiling        the end-to-end leg's host-to-device copies alone (all ranks at once): the ceiling of e2e at N GPUs
All inputs are made with the engine itself (no oracle on any timed or product path)."""
import time

import numpy as np
import torch
import torch.distributed as dist

from tools.workloads import BASE, L, mults_of_base, rand_scalars
from zkp_b200 import toolbox as PT

# RistrettoPoint::hash_from_bytes::<Sha512>(G.compress()) of benches/dleq.rs:52 (value pinned by tests/test_oracle.py)
DLEQ_H = np.frombuffer(bytes.fromhex("90ca11cd6c6227cb0abc39e2710c444ae6617ea81898e716353f3410d9656605"), dtype=np.uint8)


def _pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def _max_ranks(x, world):
    t = torch.tensor(list(x), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def _read(d_res):
    r = d_res.cpu().numpy()
    status, ident = np.frombuffer(r[32:40].tobytes(), dtype=np.int32)
    return r[:32].tobytes(), int(status), int(ident)


def _timed_dev(stream, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _timed_host(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


def h2d_ceiling(pairs, steps, barrier, world):
    """pairs = [(device tensor, pinned host tensor)]: the copies of one e2e step, nothing else."""
    nbytes = sum(h.numel() * h.element_size() for _, h in pairs)

    def step():
        for d, h in pairs:
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
    step()
    barrier()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    barrier()
    ms = (time.perf_counter() - t) / steps * 1e3
    ms = _max_ranks([ms], world)[0]
    return {"ms_per_step": ms, "bytes_per_rank_per_step": int(nbytes), "GBs_per_rank": nbytes / ms / 1e6,
            "GBs_all_ranks": world * nbytes / ms / 1e6,
            "what": "cudaMemcpyAsync of the e2e leg's pinned buffers only, all ranks at once, max over ranks"}


def cmz_prove(eng, st, sec, limbs, entropy, fe_rates, reps, barrier, world):
    """zkp_prove_batch from pinned host buffers: witnesses + limb-form points in, encodings + commitments + responses out."""
    N = sec.shape[0]
    sec_p, limbs_p, ent_p = _pin(sec), _pin(limbs), _pin(entropy)
    outs = tuple(torch.zeros(shp, dtype=torch.uint8).pin_memory().numpy()
                 for shp in ((N, st.p, 32), (N, st.k, 32), (N, st.m, 32)))
    call = lambda: st.prove_many_device(eng, b"CMZ", sec_p, limbs_p, ent_p, out=outs)
    call()
    barrier()
    ms = _timed_host(call, reps, warm=1)
    barrier()
    ms = _max_ranks([ms], world)[0]
    # arithmetic of one proof on the comb path: per-proof combs of P and Q, 16 units of <= 2 terms with 64
    # doublings each, 11 additions of projective and 20 of affine table entries per column, 25 + 11 compressions
    sq, mul = fe_rates
    dbl, add_p, add_a, enc = (4, 4), (0, 8), (0, 7), (255, 22)
    ops = [(2 * 192 + 16 * 64, dbl), (11 * 64 + 2 * 10, add_p), (20 * 64, add_a), (25 + 11, enc)]
    s_cnt = sum(c * o[0] for c, o in ops)
    m_cnt = sum(c * o[1] for c, o in ops)
    ideal_ms = N * (s_cnt / sq + m_cnt / mul) * 1e3
    enc_out, com_out, resp_out = (np.array(o) for o in outs)
    return ({"proofs_per_gpu": N, "ms_per_call": ms, "value": world * N / ms * 1e3, "unit": "proofs/s",
             "api": "zkp_prove_batch via zkph_prove_many_device (C ABI), pinned host buffers, H2D and D2H inside",
             "h2d_bytes": int(sec_p.nbytes + limbs_p.nbytes + ent_p.nbytes), "d2h_bytes": int(sum(o.nbytes for o in outs)),
             "integer_pipe": {"squarings_per_proof": s_cnt, "multiplications_per_proof": m_cnt,
                              "ms_at_calibrated_field_rates": ideal_ms, "frac_of_calibrated": ideal_ms / ms,
                              "what": "field operations of the comb prover per proof at the register-resident rates of this "
                                      "run, against the whole call (copies and the constant-time table scans included)"}},
            (enc_out, com_out, resp_out))


def dleq_statement():
    """The statement of benches/dleq.rs:37-47 (A = x G, B = x H; A, B per proof, G, H static): 2 + 4 N MSM terms."""
    return PT.Statement("dleq", "DLEQProof", ["x"], ["A", "B"], ["G", "H"], [("A", [("x", "G")]), ("B", [("x", "H")])])


def dleq_batch_verify(eng, stream, log2, steps, rank, barrier, world):
    N = 1 << log2
    st = dleq_statement()
    xs = np.zeros((N, 1, 32), np.uint8)
    x0 = 89327492234 + rank * N                                                        # benches/dleq.rs:198
    xs[:, 0, :8] = (np.arange(N, dtype=np.uint64) + np.uint64(x0)).view(np.uint8).reshape(N, 8)
    GH = np.stack([BASE, DLEQ_H])
    AB, valid = eng.msm_vartime_batched(np.repeat(xs.reshape(N, 32), 2, axis=0), np.tile(GH, (N, 1)),
                                        np.arange(2 * N + 1, dtype=np.uint64))
    assert valid.all()
    enc = np.empty((N, 4, 32), np.uint8)                                              # A, B, G, H
    enc[:, :2], enc[:, 2], enc[:, 3] = AB.reshape(N, 2, 32), BASE, DLEQ_H
    limbs, valid = eng.decompress_batch(enc.reshape(-1, 32))
    assert valid.all()
    entropy = np.random.default_rng(900 + rank).integers(0, 256, size=(N, 32), dtype=np.uint8)
    sec_p, limbs_p, ent_p = _pin(xs), _pin(limbs.reshape(N, 4, 20)), _pin(entropy)
    outs = tuple(torch.zeros(shp, dtype=torch.uint8).pin_memory().numpy() for shp in ((N, 4, 32), (N, 2, 32), (N, 1, 32)))
    prove = lambda: st.prove_many_device(eng, b"DLEQBatchTest", sec_p, limbs_p, ent_p, out=outs)
    prove()
    ms_prove = _timed_host(prove, 2, warm=0)
    enc2, com, resp = outs
    assert (enc2 == enc).all()
    inst = _pin(enc2[:, :2].transpose(1, 0, 2))
    common = np.ascontiguousarray(enc2[0, 2:])
    com_h, resp_h = _pin(com), _pin(resp)
    seed = bytes(range(32))
    co, po = st.batch_verify_device(eng, com_h, resp_h, b"DLEQBatchTest", inst, common, seed, want_msm_inputs=True)
    bad = resp_h.copy()
    bad[N // 3, 0, 0] ^= 1
    try:
        st.batch_verify_device(eng, com_h, bad, b"DLEQBatchTest", inst, common, seed)
        raise SystemExit("tampered DLEQ proof accepted")
    except PT.VerificationFailure:
        pass
    n = co.shape[0]
    assert n == 2 + 4 * N
    # device-resident MSM
    h_sc, h_pt = torch.from_numpy(co).pin_memory(), torch.from_numpy(po).pin_memory()
    d_sc, d_pt = h_sc.cuda(non_blocking=True), h_pt.cuda(non_blocking=True)
    d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    eng.set_stream(stream.cuda_stream)
    barrier()
    ms_dev = _timed_dev(stream, lambda: eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), n, d_res.data_ptr()), steps)
    assert _read(d_res)[1:] == (0, 1)
    window = int(eng.stage_ms()["window"])
    # host coefficient buffers through zkp_batch_verify
    sc_np, pt_np = h_sc.numpy(), h_pt.numpy()

    def e2e():
        ok, rc = eng.batch_verify(sc_np[:2], pt_np[:2], sc_np[2:], pt_np[2:], 4, N)
        assert ok and rc == 0
    barrier()
    ms_e2e = _timed_host(e2e, steps)
    eng.set_stream(None)
    barrier()
    ms_proofs = _timed_host(lambda: st.batch_verify_device(eng, com_h, resp_h, b"DLEQBatchTest", inst, common, seed), steps)
    barrier()
    ms_dev, ms_e2e, ms_proofs, ms_prove = _max_ranks([ms_dev, ms_e2e, ms_proofs, ms_prove], world)
    rate = lambda ms: world * N / ms * 1e3
    return {"proofs_per_gpu": N, "msm_terms_per_gpu": int(n), "window": window, "bytes_per_proof": 256,
            "value": rate(ms_dev), "unit": "proofs/s", "ms_per_step": ms_dev,
            "e2e": {"value": rate(ms_e2e), "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n * 64),
                    "api": "zkp_batch_verify (C ABI), pinned host buffers"},
            "e2e_from_proofs": {"value": rate(ms_proofs), "ms_per_step": ms_proofs,
                                "h2d_bytes_per_step": int(inst.nbytes + com_h.nbytes + resp_h.nbytes + common.nbytes),
                                "api": "zkp_batch_verify_proofs via zkph_batch_verify_device: transcripts, challenges, weights, "
                                       "fold and MSM on the GPU"},
            "prove": {"value": rate(ms_prove), "ms_per_call": ms_prove, "api": "zkp_prove_batch, pinned host buffers"},
            "whole_path_hbm_GBs": N * 256 / ms_dev / 1e6,
            "data": "2^%d real DLEQ proofs per GPU made by zkp_prove_batch (x_j = 89327492234 + j), accept checked, tampered "
                    "response rejected" % log2}


def raw_msm_sweep(eng, stream, sizes_log2, rank, world, cpu_port_max_log2, host_threads):
    """ONE MSM of 2^lg uniformly random terms (full-size scalars, points from a pool of 2^16 multiples of the basepoint).
    world = 1: zkp_msm_vartime_dev, and the C port of the reference's CPU algorithms on the same input beside it (bytes
    compared).  world > 1: the terms are cut over the ranks; partial sums, one NCCL all-gather, verdict on every rank."""
    rng = np.random.default_rng(2026)                      # the same stream on every rank: one global instance
    K = 1 << 16
    pool = mults_of_base(eng, rand_scalars(rng, (K,)))
    d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    d_part = torch.zeros(160, dtype=torch.uint8, device="cuda")
    d_all = torch.zeros(160 * world, dtype=torch.uint8, device="cuda")
    cref = None
    if world == 1 and cpu_port_max_log2 >= 8:
        from oracle import cref                           # baseline leg only (like cpu_baseline): never the product path
    eng.set_stream(stream.cuda_stream)
    rows = []
    for lg in sizes_log2:
        n = 1 << lg
        sc = rand_scalars(rng, (n,))
        pt = pool[rng.integers(0, K, size=n)]
        lo, hi = rank * n // world, (rank + 1) * n // world
        d_sc = torch.from_numpy(sc[lo:hi]).cuda()
        d_pt = torch.from_numpy(np.ascontiguousarray(pt[lo:hi])).cuda()
        cnt = hi - lo
        if world == 1:
            fn = lambda: eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), cnt, d_res.data_ptr())
        else:
            def fn():
                eng.msm_vartime_partial_dev(d_sc.data_ptr(), d_pt.data_ptr(), cnt, d_res.data_ptr(), d_part.data_ptr())
                dist.all_gather_into_tensor(d_all, d_part)       # on the current (= the engine's) stream
                eng.partials_verdict_dev(d_all.data_ptr(), world, d_res.data_ptr())
        if world > 1:
            dist.barrier()
        ms = _timed_dev(stream, fn, 5)
        ms = _max_ranks([ms], world)[0]
        enc, status, _ = _read(d_res)
        assert status == 0
        row = {"log2_n": lg, "gpu_ms": ms, "gpu_terms_per_s": n / ms * 1e3}
        if world > 1:   # every rank holds the same sum
            encs = [None] * world
            dist.all_gather_object(encs, enc.hex())
            assert len(set(encs)) == 1
            row["sum_encoding"] = enc.hex()[:16]
        if cref is not None and lg <= cpu_port_max_log2:
            t = time.perf_counter()
            cpu_enc = cref.msm_vartime(sc, pt, threads=host_threads)
            dt = time.perf_counter() - t
            assert cpu_enc == enc, "GPU and CPU port disagree at n = 2^%d" % lg
            row.update(cpu_port_ms=dt * 1e3, cpu_terms_per_s=n / dt, cpu_threads=host_threads, bytes_equal=True)
        rows.append(row)
    eng.set_stream(None)
    return rows


def single_verdict(eng, stream, d_scal, d_pts, n_terms, num_s, rank, world, steps):
    """ONE batch over all ranks: rank r's static coefficient 0 gets delta_r with sum(delta) = 0 mod l, and static point 0
    is the basepoint on every rank, so a shard's MSM is delta_r * B (not the identity) while the sum over the shards is.
    Per step: the shard's MSM, k_ext_to_limbs, NCCL all-gather of 160 bytes per rank on the compute stream, the sum and
    the identity test on every rank."""
    deltas = [r + 1 for r in range(world - 1)]
    deltas.append((L - sum(deltas)) % L)
    sc = d_scal.clone()
    pts = d_pts.clone()
    sc[0] = torch.from_numpy(np.frombuffer(int(deltas[rank]).to_bytes(32, "little"), dtype=np.uint8).copy()).cuda()
    pts[0] = torch.from_numpy(BASE.copy()).cuda()
    d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    d_own = torch.zeros(64, dtype=torch.uint8, device="cuda")
    d_part = torch.zeros(160, dtype=torch.uint8, device="cuda")
    d_all = torch.zeros(160 * world, dtype=torch.uint8, device="cuda")
    eng.set_stream(stream.cuda_stream)

    def gather_and_verdict():
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_part)
        else:
            d_all.copy_(d_part)
        eng.partials_verdict_dev(d_all.data_ptr(), world, d_res.data_ptr())

    def step():
        eng.msm_vartime_partial_dev(sc.data_ptr(), pts.data_ptr(), n_terms, d_own.data_ptr(), d_part.data_ptr())
        gather_and_verdict()
    if world > 1:
        dist.barrier()
    ms = _timed_dev(stream, step, steps)
    own_identity = _read(d_own)[2]
    _, status, accept = _read(d_res)
    gather_us = _timed_dev(stream, gather_and_verdict, 20) * 1e3
    # a tampered shard (rank world - 1 flips one coefficient bit) must void the one verdict on every rank
    if rank == world - 1:
        sc[num_s + 7, 0] ^= 1
    step()
    torch.cuda.synchronize()
    tampered_accept = _read(d_res)[2]
    eng.set_stream(None)
    ms, gather_us = _max_ranks([ms, gather_us], world)
    flags = torch.tensor([own_identity, accept, tampered_accept, status], dtype=torch.int32, device="cuda")
    allf = [torch.zeros_like(flags) for _ in range(world)]
    if world > 1:
        dist.all_gather(allf, flags)
    else:
        allf = [flags]
    allf = [[int(v) for v in f.tolist()] for f in allf]
    ok = all(f[1] == 1 and f[2] == 0 and f[3] == 0 for f in allf) and (world == 1 or all(f[0] == 0 for f in allf))
    assert ok, "single-verdict mode failed: %r" % (allf,)
    return {"mode": "single_verdict", "ranks": world, "ms_per_step": ms, "gather_plus_verdict_us": gather_us,
            "shard_alone_is_identity": [f[0] for f in allf], "accept": [f[1] for f in allf],
            "accept_with_one_tampered_shard": [f[2] for f in allf], "bytes_gathered_per_rank": 160,
            "what": "one batch, one verdict: zkp_msm_vartime_partial_dev -> all_gather_into_tensor (NCCL, compute stream) -> "
                    "zkp_partials_verdict_dev"}
