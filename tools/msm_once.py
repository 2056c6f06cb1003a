"""One raw MSM of 2^LOG2 random terms through zkp_msm_vartime_dev-equivalent host entry (for ncu launch lists).
Usage: python tools/msm_once.py [log2_n] [repeats]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine  # noqa: E402
from tools.workloads import rand_scalars, mults_of_base  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
eng = Engine(0)
rng = np.random.default_rng(3)
pts = mults_of_base(eng, rand_scalars(rng, (min(n, 4096),)))
pts = np.ascontiguousarray(np.tile(pts, (n // pts.shape[0] + 1, 1))[:n])
sc = rand_scalars(rng, (n,))
for _ in range(reps):
    enc, ident, _ = eng.msm_vartime(sc, pts)
print("done", n, enc.hex()[:16])
