"""Single-verdict mode over real NCCL ranks (run under torchrun; world size 1 works too): ONE batch, ONE weight stream, ONE
accept bit (/root/reference/src/toolbox/batch_verifier.rs:219-234) with the MSM cut over the GPUs.  Rank r holds the
terms (s_i, P_i) of its slice and the negated terms (l - s_j, P_j) of the next rank's slice: every shard's sum is a
non-trivial point, the sum over all shards is the identity.  Checked through both surfaces:
  device:  zkp_msm_vartime_partial_dev -> all_gather_into_tensor (compute stream) -> zkp_partials_verdict_dev
  host:    zkp_batch_verify_partial -> parallel.gather_partial_sums -> zkp_partials_verdict
at a size that takes the Pippenger pipeline and at one that takes the small-MSM path; a tampered shard must void the
verdict on every rank.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zkp_b200 import Engine, parallel  # noqa: E402
from tools.workloads import L, mults_of_base, rand_scalars  # noqa: E402


def neg_mod_l(s):
    out = np.empty_like(s)
    for i in range(s.shape[0]):
        v = int.from_bytes(s[i].tobytes(), "little")
        out[i] = np.frombuffer(((L - v) % L).to_bytes(32, "little"), dtype=np.uint8)
    return out


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    report = {"world": world, "cases": []}
    for per_rank in (96, 6000):                      # small-MSM path (<= 1024 terms per shard) and the sort pipeline
        rng = np.random.default_rng(77)              # the same global instance on every rank
        pool = mults_of_base(eng, rand_scalars(rng, (64,)))
        sc_all = rand_scalars(rng, (world, per_rank))
        pi_all = rng.integers(0, 64, size=(world, per_rank))
        nxt = (rank + 1) % world
        sc = np.concatenate([sc_all[rank], neg_mod_l(sc_all[nxt])])
        pt = np.concatenate([pool[pi_all[rank]], pool[pi_all[nxt]]])
        n = sc.shape[0]
        d_sc, d_pt = torch.from_numpy(sc).cuda(), torch.from_numpy(np.ascontiguousarray(pt)).cuda()
        d_own = torch.zeros(64, dtype=torch.uint8, device="cuda")
        d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
        d_part = torch.zeros(160, dtype=torch.uint8, device="cuda")
        d_all = torch.zeros(160 * world, dtype=torch.uint8, device="cuda")

        def device_verdict(scalars):
            eng.msm_vartime_partial_dev(scalars.data_ptr(), d_pt.data_ptr(), n, d_own.data_ptr(), d_part.data_ptr())
            dist.all_gather_into_tensor(d_all, d_part)
            eng.partials_verdict_dev(d_all.data_ptr(), world, d_res.data_ptr())
            torch.cuda.synchronize()
            own = np.frombuffer(d_own.cpu().numpy()[32:40].tobytes(), dtype=np.int32)
            res = np.frombuffer(d_res.cpu().numpy()[32:40].tobytes(), dtype=np.int32)
            return int(own[0]), int(own[1]), int(res[0]), int(res[1])
        own_status, own_ident, status, accept = device_verdict(d_sc)
        bad = d_sc.clone()
        if rank == world - 1:
            bad[3, 0] ^= 1
        _, _, _, accept_bad = device_verdict(bad)
        # the host surface: static part empty, the shard as a 1-row "instance matrix"
        eng.set_stream(None)
        part = eng.batch_verify_partial(sc[:0], pt[:0], sc, pt, 1, n)
        allp = parallel.gather_partial_sums(part, device="cuda")
        h_accept, h_enc = eng.partials_verdict(allp)
        eng.set_stream(stream.cuda_stream)
        flags = torch.tensor([own_status, own_ident, status, accept, accept_bad, int(h_accept)], dtype=torch.int32, device="cuda")
        out = [torch.zeros_like(flags) for _ in range(world)]
        dist.all_gather(out, flags)
        rows = [[int(v) for v in t.tolist()] for t in out]
        ok = all(r[0] == 0 and r[2] == 0 and r[3] == 1 and r[4] == 0 and r[5] == 1 for r in rows)
        ok = ok and (world == 1 or all(r[1] == 0 for r in rows)) and h_enc == bytes(32)
        report["cases"].append({"terms_per_rank": n, "ok": bool(ok), "per_rank": rows})
    report["ok"] = all(c["ok"] for c in report["cases"])
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()
