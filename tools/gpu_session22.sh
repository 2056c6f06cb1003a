#!/bin/bash
# round 2, GPU session 22: occupancy points of k_ingest2 beyond 24 warps/SM (72 and 64 registers)
set -u
O=gpurun_out
for v in 2 3 4 5; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg --ingest-variant $v > $O/s22_v$v.json 2> $O/s22_v$v.err
  python - $v <<'P'
import json, sys
d = json.loads(open("gpurun_out/s22_v%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("variant", sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"])
P
done
