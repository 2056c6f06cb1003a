"""Measurements for the other BASELINE.json configs (the headline config lives in bench.py):
  configs[0]  DLEQ 1 prove + 1 verify through the host mirror (plumbing; latency)
  configs[1]  CMZ'13 cred_show_10: prove 2^16 proofs (constant-time batched MSM path)
  configs[2]  DLEQ: BatchVerifier::verify_batchable over 2^20 real proofs (one Pippenger MSM)
  configs[4]  raw ristretto255 MSM sweep 2^8 .. 2^22 (device-resident) next to the C port on the host cores
Every proof set is REAL: instances are built with the engine, proofs with prove_many, and verification must accept
(and reject after tampering).  Usage: python tools/bench_configs.py [--log2-cmz 16] [--log2-dleq 20] [--sweep-max 22]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine, toolbox as PT  # noqa: E402

from tools.workloads import BASE, cmz_instances, mults_of_base, rand_scalars, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-cmz", type=int, default=16)
    ap.add_argument("--log2-dleq", type=int, default=20)
    ap.add_argument("--sweep-max", type=int, default=22)
    ap.add_argument("--cpu-port", action="store_true",
                    help="also time the C port of the reference's CPU algorithms (oracle/_ref) beside the raw MSM sweep, as "
                         "bench.py's cpu_baseline leg does; it doubles as a byte-for-byte check of the GPU results")
    ap.add_argument("--prove-comb", action="store_true",
                    help="batch proving through the comb path (engine option prove_comb; off by default until measured)")
    ap.add_argument("--out", default="gpurun_out/configs.json")
    args = ap.parse_args()
    eng = Engine(0)
    if args.prove_comb:
        eng.set_option("prove_comb", 1)
    rng = np.random.default_rng(2026)
    threads = os.cpu_count() or 1
    res = {"host_threads": threads, "gpu": torch.cuda.get_device_name(0), "prove_comb": bool(args.prove_comb)}

    # ---- configs[0]: DLEQ 1 prove + 1 verify (latency through the host mirror) -----------------------------
    st = PT.dleq_statement()
    x = (89327492234).to_bytes(32, "little")
    G = BASE
    H = mults_of_base(eng, rand_scalars(rng, (1,)))[0]
    AB, _ = eng.msm_vartime_batched(np.frombuffer(x * 2, np.uint8).reshape(2, 32), np.stack([G, H]),
                                    np.arange(3, dtype=np.uint64))
    encs = np.stack([AB[0], AB[1], H, G])
    limbs, _ = eng.decompress_batch(encs)
    for _ in range(3):
        (chal, resp), enc = st.prove_compact(eng, b"DLEQTest", np.frombuffer(x, np.uint8), limbs, b"seed")
    t_p, ((chal, resp), enc) = timed(lambda: st.prove_compact(eng, b"DLEQTest", np.frombuffer(x, np.uint8), limbs, b"seed"), 5)
    t_v, _ = timed(lambda: st.verify_compact(eng, (chal, resp), b"DLEQTest", enc), 5)
    res["config0_dleq_single"] = {"prove_ms": t_p * 1e3, "verify_compact_ms": t_v * 1e3}
    print("config0", res["config0_dleq_single"], flush=True)

    # ---- configs[1]: CMZ prove 2^16 ----------------------------------------------------------------------------
    N = 1 << args.log2_cmz
    st, sec, limbs, enc_expected = cmz_instances(eng, N, rng)
    entropy = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
    st.prove_many(eng, b"CMZ", sec[:256], limbs[:256], entropy[:256], threads=threads)     # warm-up
    t_prove, (enc, com, resp) = timed(lambda: st.prove_many(eng, b"CMZ", sec, limbs, entropy, threads=threads))
    assert (enc == enc_expected).all()
    # the same proofs with the per-proof transcript / nonce / response work on the GPU (zkp_prove_batch)
    st.prove_many_device(eng, b"CMZ", sec[:256], limbs[:256], entropy[:256])               # warm-up
    t_prove_dev, (enc_d, com_d, resp_d) = timed(lambda: st.prove_many_device(eng, b"CMZ", sec, limbs, entropy), 2)
    assert (enc_d == enc).all() and (com_d == com).all() and (resp_d == resp).all()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    sec_p, limbs_p, ent_p = pin(sec), pin(limbs), pin(entropy)
    outs = tuple(torch.zeros(shp, dtype=torch.uint8).pin_memory().numpy() for shp in ((N, 25, 32), (N, 11, 32), (N, 21, 32)))
    t_prove_pin, (enc_d, com_d, resp_d) = timed(
        lambda: st.prove_many_device(eng, b"CMZ", sec_p, limbs_p, ent_p, out=outs), 3)
    assert (enc_d == enc).all() and (com_d == com).all() and (resp_d == resp).all()
    # device part alone: the N*11 constant-time MSMs + compressions from prepared inputs
    k_terms = [2] * 10 + [11]
    sc_idx = [i for c in range(10) for i in (c, 10 + c)] + list(range(10)) + [20]
    pt_idx = [i for c in range(10) for i in (10, 23)] + list(range(13, 23)) + [11]
    bl = rand_scalars(rng, (N, 21))
    ct_sc = bl[:, sc_idx].reshape(-1, 32)
    ct_pt = limbs[:, pt_idx].reshape(-1, 20)
    off = np.concatenate([[0], np.cumsum(np.tile(np.array(k_terms, dtype=np.uint64), N))]).astype(np.uint64)
    eng.msm_ct_batched(ct_sc[:3100], ct_pt[:3100], off[:1101], limbs=True)
    t_ct, _ = timed(lambda: eng.msm_ct_batched(ct_sc, ct_pt, off, limbs=True))
    t_cmp, _ = timed(lambda: eng.compress_batch(limbs.reshape(-1, 20)))
    res["config1_cmz_prove"] = {"proofs": N, "prove_many_s": t_prove, "proofs_per_s": N / t_prove,
                                "prove_many_device_s": t_prove_dev, "proofs_per_s_device_front_end": N / t_prove_dev,
                                "prove_many_device_pinned_s": t_prove_pin,
                                "proofs_per_s_device_front_end_pinned": N / t_prove_pin,
                                "device_ct_msm_plus_compress_s": t_ct, "ct_msms_per_s": N * 11 / t_ct,
                                "device_compress_25_points_per_proof_s": t_cmp,
                                "note": "prove_many = batched compress + host Merlin/blindings (%d threads) + one "
                                        "zkp_msm_ct_batched call + host challenges/responses" % threads}
    print("config1", res["config1_cmz_prove"], flush=True)
    # the proofs verify (real data for a CMZ batch verification through the whole host path)
    ni = 13
    inst = np.ascontiguousarray(enc[:, :ni].transpose(1, 0, 2))
    t_bv, hs = timed(lambda: st.batch_verify(eng, com, resp, b"CMZ", inst, enc[0, ni:], b"rho", threads=threads))
    bad = resp.copy()
    bad[N // 2, 3, 0] ^= 1
    try:
        st.batch_verify(eng, com, bad, b"CMZ", inst, enc[0, ni:], b"rho", threads=threads)
        raise SystemExit("tampered CMZ batch accepted")
    except PT.VerificationFailure:
        pass
    seed = bytes(range(32))
    st.batch_verify_device(eng, com[:64], resp[:64], b"CMZ", np.ascontiguousarray(inst[:, :64]), enc[0, ni:], seed)
    t_dev, _ = timed(lambda: st.batch_verify_device(eng, com, resp, b"CMZ", inst, enc[0, ni:], seed), 2)
    try:
        st.batch_verify_device(eng, com, bad, b"CMZ", inst, enc[0, ni:], seed)
        raise SystemExit("tampered CMZ batch accepted by the device front end")
    except PT.VerificationFailure:
        pass
    res["cmz_batch_verify_real_proofs"] = {"proofs": N, "total_s": t_bv, "host_hash_and_fold_s": hs,
                                           "proofs_per_s_end_to_end_incl_host": N / t_bv,
                                           "device_front_end_total_s": t_dev,
                                           "proofs_per_s_device_front_end": N / t_dev}
    print("cmz batch verify (real proofs)", res["cmz_batch_verify_real_proofs"], flush=True)

    # ---- configs[3], bitmap mode: the batch cut into sub-batches, one accept bit each (the bits are what the ranks
    # all-gather); one tampered proof must clear exactly its sub-batch's bit ---------------------------------------
    K_sub = 16
    S_sub = N // K_sub
    bad_j = 5 * S_sub + 17
    bad = resp.copy()
    bad[bad_j, 2, 0] ^= 1

    def bitmap(responses):
        bits = []
        for q in range(K_sub):
            lo, hi = q * S_sub, (q + 1) * S_sub
            try:
                st.batch_verify_device(eng, com[lo:hi], responses[lo:hi], b"CMZ", np.ascontiguousarray(inst[:, lo:hi]),
                                       enc[0, ni:], seed)
                bits.append(1)
            except PT.VerificationFailure:
                bits.append(0)
        return bits
    t_bits, bits_ok = timed(lambda: bitmap(resp), 2)
    bits_bad = bitmap(bad)
    assert bits_ok == [1] * K_sub and bits_bad == [1] * 5 + [0] + [1] * (K_sub - 6), (bits_ok, bits_bad)
    res["config3_bitmap_mode"] = {"proofs": N, "sub_batches": K_sub, "proofs_per_sub_batch": S_sub, "total_s": t_bits,
                                  "proofs_per_s": N / t_bits, "bitmap_valid": bits_ok, "bitmap_one_tampered_proof": bits_bad}
    print("config3-bitmap", res["config3_bitmap_mode"], flush=True)

    # ---- configs[2]: DLEQ batch verify 2^20 -------------------------------------------------------------------
    N = 1 << args.log2_dleq
    st = PT.dleq_statement()
    xs = np.zeros((N, 1, 32), np.uint8)
    for j, v in enumerate(range(89327492234, 89327492234 + N)):                      # benches/dleq.rs:198
        xs[j, 0, :8] = np.frombuffer(int(v).to_bytes(8, "little"), np.uint8)
    GH = np.stack([G, H])
    ab_sc = np.repeat(xs.reshape(N, 32), 2, axis=0)
    ab_pt = np.tile(GH, (N, 1))
    AB, valid = eng.msm_vartime_batched(ab_sc, ab_pt, np.arange(2 * N + 1, dtype=np.uint64))
    AB = AB.reshape(N, 2, 32)
    enc = np.empty((N, 4, 32), np.uint8)                                             # A, B, H, G
    enc[:, 0], enc[:, 1], enc[:, 2], enc[:, 3] = AB[:, 0], AB[:, 1], H, G
    limbs, valid = eng.decompress_batch(enc.reshape(-1, 32))
    entropy = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
    t_prove, (enc2, com, resp) = timed(lambda: st.prove_many(eng, b"DLEQBatchTest", xs, limbs.reshape(N, 4, 20), entropy,
                                                              threads=threads))
    inst = np.ascontiguousarray(enc2[:, :3].transpose(1, 0, 2))
    co, po, hs = st.batch_verify(eng, com, resp, b"DLEQBatchTest", inst, enc2[0, 3:], b"rho", threads=threads,
                                 want_msm_inputs=True)
    t_bv, hs = timed(lambda: st.batch_verify(eng, com, resp, b"DLEQBatchTest", inst, enc2[0, 3:], b"rho", threads=threads))
    # the MSM alone, device-resident
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    d_sc, d_pt = torch.from_numpy(co).cuda(), torch.from_numpy(po).cuda()
    d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), co.shape[0], d_res.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), co.shape[0], d_res.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    r = d_res.cpu().numpy()
    assert tuple(np.frombuffer(r[32:40].tobytes(), dtype=np.int32)) == (0, 1)
    st.batch_verify_device(eng, com[:64], resp[:64], b"DLEQBatchTest", np.ascontiguousarray(inst[:, :64]), enc2[0, 3:], bytes(32))
    t_dev, _ = timed(lambda: st.batch_verify_device(eng, com, resp, b"DLEQBatchTest", inst, enc2[0, 3:], bytes(range(32))), 2)
    res["config2_dleq_batch_verify"] = {"proofs": N, "msm_terms": int(co.shape[0]), "msm_device_ms": ms,
                                        "device_front_end_total_s": t_dev, "proofs_per_s_device_front_end": N / t_dev,
                                        "proofs_per_s_msm_device": N / (ms * 1e-3),
                                        "whole_host_path_s": t_bv, "host_hash_and_fold_s": hs,
                                        "proofs_per_s_incl_host": N / t_bv, "prove_many_s": t_prove}
    print("config2", res["config2_dleq_batch_verify"], flush=True)

    # ---- configs[4]: raw MSM sweep ------------------------------------------------------------------------------
    cref = None
    if args.cpu_port:        # baseline leg only (like bench.py's cpu_baseline): never on the product path
        from oracle import cref
    K = 1 << 16
    pool = mults_of_base(eng, rand_scalars(rng, (K,)))
    sweep = []
    for lg in range(8, args.sweep_max + 1, 2):
        n = 1 << lg
        sc = rand_scalars(rng, (n,))
        pt = pool[rng.integers(0, K, size=n)]
        d_sc, d_pt = torch.from_numpy(sc).cuda(), torch.from_numpy(pt).cuda()
        for _ in range(2):
            eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), n, d_res.data_ptr())
        torch.cuda.synchronize()
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            eng.msm_vartime_dev(d_sc.data_ptr(), d_pt.data_ptr(), n, d_res.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gpu_enc = d_res.cpu().numpy()[:32].tobytes()
        row = {"log2_n": lg, "gpu_ms": ms, "gpu_terms_per_s": n / (ms * 1e-3)}
        if cref is not None and lg <= 18:
            t = time.perf_counter()
            cpu_enc = cref.msm_vartime(sc, pt, threads=threads)
            dt = time.perf_counter() - t
            assert cpu_enc == gpu_enc, "GPU and CPU port disagree at n=2^%d" % lg
            row.update(cpu_port_ms=dt * 1e3, cpu_terms_per_s=n / dt, cpu_threads=threads)
        sweep.append(row)
        print("sweep", row, flush=True)
    res["config4_raw_msm_sweep"] = sweep
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
