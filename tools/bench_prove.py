"""Tuning sweep of zkp_prove_batch (2^LOG2 CMZ cred_show_10 proofs, pinned host buffers): the prover paths and the slice size of the copy/compute pipeline.  Usage: python tools/bench_prove.py
[--log2 16] [--out gpurun_out/prove.json] [--quick]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine  # noqa: E402
from tools.workloads import cmz_instances, timed  # noqa: E402

lg = int(sys.argv[sys.argv.index("--log2") + 1]) if "--log2" in sys.argv else 16
out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "gpurun_out/prove.json"
N = 1 << lg
eng = Engine(0)
rng = np.random.default_rng(11)
st, sec, limbs, enc = cmz_instances(eng, N, rng)
entropy = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
sec_p, limbs_p, ent_p = pin(sec), pin(limbs), pin(entropy)
outs = tuple(torch.zeros(shp, dtype=torch.uint8).pin_memory().numpy() for shp in ((N, 25, 32), (N, 11, 32), (N, 21, 32)))
ref = None
rows = []
DEFAULTS = {"prove_comb": 2, "prove_pipe_chunk": 1 << 14, "prove_piece": 2}
cases = [{"prove_comb": 0, "prove_pipe_chunk": 0}, {"prove_comb": 1, "prove_pipe_chunk": 0}, {"prove_comb": 2, "prove_pipe_chunk": 0}]
for pc in (1 << 12, 1 << 13, 1 << 14, 1 << 15):
    cases.append({"prove_pipe_chunk": pc})
cases += [{"prove_piece": 3, "prove_pipe_chunk": 0}, {"prove_piece": 1, "prove_pipe_chunk": 0}]
if "--quick" in sys.argv:
    cases = [{"prove_comb": 1, "prove_pipe_chunk": 0}, {"prove_pipe_chunk": 0}, {}, {"prove_pipe_chunk": 1 << 15}]
for case in cases:
    opts = dict(DEFAULTS)
    opts.update(case)
    for k, v in opts.items():
        eng.set_option(k, v)
    call = lambda: st.prove_many_device(eng, b"CMZ", sec_p, limbs_p, ent_p, out=outs)
    call()
    t, res = timed(call, 4)
    got = tuple(np.array(o) for o in res)
    if ref is None:
        ref = got
    assert all((a == b).all() for a, b in zip(ref, got)), case
    rows.append({"options": case, "ms": t * 1e3, "proofs_per_s": N / t})
    print(rows[-1], flush=True)
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump({"proofs": N, "rows": rows}, open(out, "w"), indent=1)
