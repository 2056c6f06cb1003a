#!/bin/bash
# round 2, GPU session 32: slab size / resident front-end grid of the from-proofs leg and the DLEQ window, at HEAD
set -u
for o in "bv_chunk_terms=262144" "bv_chunk_terms=393216" "bv_chunk_terms=524288" "bv_chunk_terms=786432" "bv_chunk_terms=524288 bv_prep_blocks=3 bv_prep_smem_kb=40" "bv_chunk_terms=393216 bv_prep_blocks=3 bv_prep_smem_kb=40"; do
  a=""; for kv in $o; do a="$a --opt $kv"; done
  python tools/bv_timeline.py $a 2>&1 | grep "wall ms"
done
for w in 16 18; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-proofs-leg --window $w --sweep-max-log2 8 > gpurun_out/s32_w$w.json 2> gpurun_out/s32_w$w.err
  python - $w <<'P'
import json, sys
d = json.loads(open("gpurun_out/s32_w%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
dl = d["configs"]["dleq_batch_verify"]
print("window", sys.argv[1], "dleq ms", round(dl["ms_per_step"], 3), "e2e", round(dl["e2e"]["ms_per_step"], 3))
P
done
