#!/bin/bash
# round 2, GPU session 8: occupancy points of k_ingest2 after the carry-folding field arithmetic
set -u
O=gpurun_out
mkdir -p $O
for v in 0 1 2 3; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg --ingest-variant $v > $O/s8_bench_v$v.json 2> $O/s8_bench_v$v.err
  python - $v <<'P'
import json, sys
d = json.loads(open("gpurun_out/s8_bench_v%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("variant", sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"])
P
done
