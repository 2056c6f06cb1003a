#!/bin/bash
# round 2, GPU session 29: k_accumulate<5> with the next point's line prefetched into L2 / L1 (no register held)
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
set -u
O=gpurun_out
for a in 5 51 52 5 51; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-configs --no-proofs-leg --accumulate-variant $a > $O/s29_a$a.json 2> $O/s29_a$a.err
  python - $a <<'P'
import json, sys
d = json.loads(open("gpurun_out/s29_a%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("accumulate variant", sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["integer_pipe"]["k_accumulate_ms"])
P
done
timeout 600 python -m pytest tests/test_gpu_msm.py -m gpu -q 2>&1 | tail -2
