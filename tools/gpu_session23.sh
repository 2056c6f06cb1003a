#!/bin/bash
# round 2, GPU session 23: occupancy points of k_accumulate (16 / 20 / 24 warps per SM) with k_ingest2 at 28 warps
set -u
O=gpurun_out
for a in 4 5 6; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg --ingest-variant 4 --accumulate-variant $a > $O/s23_a$a.json 2> $O/s23_a$a.err
  python - $a <<'P'
import json, sys
d = json.loads(open("gpurun_out/s23_a%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("accumulate variant", sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"],
      d["roofline"]["integer_pipe"]["k_accumulate_ms"])
P
done
timeout 600 python -m pytest tests/test_gpu_msm.py -m gpu -q 2>&1 | tail -2
