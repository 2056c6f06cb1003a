"""The synthetic CMZ batch of bench.py is a VALID instance: its MSM is the identity (checked with the C oracle),
its coefficients have the reference's distribution, and a flipped bit breaks it."""
import numpy as np

import bench
from oracle import cref, ristretto as R


def _instance(N):
    sc, K, half = bench.make_cmz_batch(N, seed=5)
    sc = bench.finish_cancellation(sc, N, half)
    pts = bench._host_points(K, distinct=K)
    pidx = (np.arange(N) % half)[None, :] + (np.arange(bench.ROWS) & 1)[:, None] * half
    return sc.view(np.uint8).reshape(bench.ROWS * N, 32), pts[pidx.reshape(-1)]


def test_bench_instance_sums_to_identity_and_has_reference_distribution():
    N = 64
    scal, pts = _instance(N)
    vals = [int.from_bytes(scal[i].tobytes(), "little") for i in range(scal.shape[0])]
    assert all(v < R.L for v in vals)
    rows = np.array(vals, dtype=object).reshape(bench.ROWS, N)
    # commitment rows are -rho with 128-bit rho (batch_verifier.rs:183); instance rows are full size
    assert all((R.L - v) % R.L < 2**128 for v in rows[bench.FULL_ROWS:].reshape(-1))
    assert sum(v.bit_length() > 200 for v in rows[:bench.FULL_ROWS].reshape(-1)) > 0.9 * bench.FULL_ROWS * N
    assert cref.msm_vartime(scal, pts, threads=2) == bytes(32)
    bad = scal.copy()
    bad[5, 0] ^= 1
    assert cref.msm_vartime(bad, pts, threads=2) != bytes(32)


def test_reference_arm_line(capsys, monkeypatch):
    import json
    import sys
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1",
                                      "--ref-sample-log2", "8"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "proofs/s"
    # the arm says that a step is a bounded sample of the workload, and how large
    assert line["same_workload_size"] is False and line["sample_proofs_per_step"] == 256
