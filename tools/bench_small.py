"""Latency of ONE small MSM through zkp_msm_vartime (host buffers in, result out): the four-lane Straus path
(k_single_msm_vt) against the Pippenger pipeline, n = 1 .. 8192, to place the dispatch threshold `small_max`.
Usage: python tools/bench_small.py [--out gpurun_out/small.json]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine  # noqa: E402
from tools.workloads import mults_of_base, rand_scalars  # noqa: E402

out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "gpurun_out/small.json"
eng = Engine(0)
rng = np.random.default_rng(3)
pool = mults_of_base(eng, rand_scalars(rng, (8192,)))
rows = []
for n in (1, 6, 36, 128, 256, 512, 1024, 2048, 4096, 8192):
    sc = rand_scalars(rng, (n,))
    pt = np.ascontiguousarray(pool[:n])
    row = {"n": n}
    ref = None
    for name, small_max, groups in (("pipeline", 0, 1024), ("small_g1024", 1 << 20, 1024), ("small_g256", 1 << 20, 256),
                                    ("small_g4096", 1 << 20, 4096)):
        eng.set_option("small_max", small_max)
        eng.set_option("small_groups", groups)
        for _ in range(3):
            enc, _, _ = eng.msm_vartime(sc, pt)
        ts = []
        for _ in range(10):
            t = time.perf_counter()
            enc, _, _ = eng.msm_vartime(sc, pt)
            ts.append(time.perf_counter() - t)
        ref = ref or enc
        assert enc == ref, (n, name)
        row[name + "_ms"] = float(np.median(ts)) * 1e3
    rows.append(row)
    print(row, flush=True)
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"), indent=1)
