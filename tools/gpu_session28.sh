#!/bin/bash
# round 2, GPU session 28 (8 GPUs, HEAD): BASELINE configs[3] exactly -- 2^24 CMZ proofs over 8 x B200 -- through bench.py under torchrun
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/s28_gpus.txt; nproc >> $O/s28_gpus.txt
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 \
  bench.py --gpus $N --steps 5 --warmup 3 > $O/s28_bench_n$N.json 2> $O/s28_bench_n$N.err; echo "bench n$N rc=$?"
tail -3 $O/s28_bench_n$N.err
python - $N <<'P'
import json, sys
try:
    d = json.loads(open("gpurun_out/s28_bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
          "from_proofs", d.get("e2e_from_proofs", {}).get("ms_per_step"))
    for k, v in d.get("configs", {}).items():
        print(k, json.dumps(v)[:500])
except Exception as e:
    print("bench parse failed", e)
P
