#!/bin/bash
# round 2, GPU session 31: H2D chunk size of the host-input pipeline below 2^19 terms
set -u
O=gpurun_out
for c in 18 17 16 18 17; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-configs --no-proofs-leg --chunk-terms-log2 $c > $O/s31_$c.json 2> $O/s31_$c.err
  python - $c <<'P'
import json, sys
d = json.loads(open("gpurun_out/s31_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("chunk_terms 2^%s" % sys.argv[1], "| ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2))
P
done
