"""Deterministic workload for ncu captures of the batch prover (zkp_prove_batch): 2^LOG2 CMZ cred_show_10 proofs through
the Straus path and through the comb path.  Usage: python tools/ncu_prover.py [--log2 14] [--comb-only] [--straus-only]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine  # noqa: E402
from tools.workloads import cmz_instances  # noqa: E402

lg = int(sys.argv[sys.argv.index("--log2") + 1]) if "--log2" in sys.argv else 14
N = 1 << lg
eng = Engine(0)
rng = np.random.default_rng(11)
st, sec, limbs, enc = cmz_instances(eng, N, rng)
entropy = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
outs = {}
for comb in (0, 1, 2):        # Straus tables; combs from global memory; combs staged in shared memory (the default)
    if (comb and "--straus-only" in sys.argv) or (not comb and "--comb-only" in sys.argv):
        continue
    if comb == 1 and "--cta-only" in sys.argv:
        continue
    eng.set_option("prove_comb", comb)
    outs[comb] = st.prove_many_device(eng, b"CMZ", sec, limbs, entropy)
ref = next(iter(outs.values()))
for comb, o in outs.items():
    assert all((a == b).all() for a, b in zip(ref, o)), "prover paths disagree (prove_comb = %d)" % comb
print("done", N, sorted(outs))
