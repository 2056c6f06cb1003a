"""Development aid: the event timeline of ONE zkp_batch_verify_proofs call at the bench size (ZKP_BV_TIMELINE=1 makes the
library print when every copy, front-end kernel and ingestion launch ended).  Usage:
  python tools/bv_timeline.py [--log2-proofs 21] [--opt key=value ...]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.workloads import cmz_instances  # noqa: E402
from zkp_b200 import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-proofs", type=int, default=21)
    ap.add_argument("--real-log2", type=int, default=16)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--pageable", action="store_true", help="host buffers in pageable memory (a slow host link)")
    args = ap.parse_args()
    eng = Engine(0)
    N, n_real = 1 << args.log2_proofs, 1 << min(args.real_log2, args.log2_proofs)
    st, sec, limbs, enc = cmz_instances(eng, n_real, np.random.default_rng(500))
    ent = np.random.default_rng(600).integers(0, 256, size=(n_real, 32), dtype=np.uint8)
    enc_p, com, resp = st.prove_many_device(eng, b"CMZ", sec, limbs, ent)
    reps = N // n_real
    pin = (lambda a: np.ascontiguousarray(a)) if args.pageable else (
        lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy())
    inst_h = pin(np.tile(np.ascontiguousarray(enc_p[:, :13].transpose(1, 0, 2)), (1, reps, 1)))
    com_h, resp_h = pin(np.tile(com, (reps, 1, 1))), pin(np.tile(resp, (reps, 1, 1)))
    common_h = np.ascontiguousarray(enc_p[0, 13:])
    seed = bytes(range(32))
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    for _ in range(3):
        st.batch_verify_device(eng, com_h, resp_h, b"CMZ", inst_h, common_h, seed)
    t0 = time.perf_counter()
    for _ in range(3):
        st.batch_verify_device(eng, com_h, resp_h, b"CMZ", inst_h, common_h, seed)
    print("opts", args.opt, "wall ms per call", round((time.perf_counter() - t0) / 3 * 1e3, 2), flush=True)
    os.environ["ZKP_BV_TIMELINE"] = "1"
    st.batch_verify_device(eng, com_h, resp_h, b"CMZ", inst_h, common_h, seed)


if __name__ == "__main__":
    main()
