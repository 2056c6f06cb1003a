#!/bin/bash
# round 2, GPU session 30: window width, work-item length and H2D chunk size re-checked with the kernels of HEAD
set -u
O=gpurun_out
i=0
for o in "--window 18" "--window 20" "--window 21" "--chunk 256" "--chunk 512" "--chunk-terms-log2 18" "--chunk-terms-log2 20"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 6 --warmup 3 --no-configs --no-proofs-leg $o > $O/s30_$i.json 2> $O/s30_$i.err
  python - $i "$o" <<'P'
import json, sys
try:
    d = json.loads(open("gpurun_out/s30_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "| ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"], d["roofline"]["integer_pipe"]["k_accumulate_ms"])
except Exception as e:
    print(sys.argv[2], "failed", e)
P
done
