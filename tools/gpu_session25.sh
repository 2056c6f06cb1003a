#!/bin/bash
# round 2, GPU session 25: share of the points in phase 1 of the host-input pipeline (phase 1 is copy-bound at 55 GB/s)
set -u
O=gpurun_out
for f in 50 52 54 56 58; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-configs --no-proofs-leg --phase1-percent $f > $O/s25_f$f.json 2> $O/s25_f$f.err
  python - $f <<'P'
import json, sys
d = json.loads(open("gpurun_out/s25_f%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("phase1_percent", sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2))
P
done
