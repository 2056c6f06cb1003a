#!/bin/bash
# round 2, GPU session 10: from-proofs path -- front-end kernel on its own high-priority stream next to the decompression,
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
# uneven split of the rows between the two ingestion phases (three term ranges per thread)
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s10_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s10_pytest.log
tail -4 $O/s10_pytest.log
i=0
while read -r prep rows chunk cap; do
  i=$((i+1))
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --bv-prep-stream $prep --bv-phase1-rows $rows \
      --bv-chunk-terms-log2 $chunk --bv-prep-smem-kb $cap > $O/s10_b$i.json 2> $O/s10_b$i.err
  python - $i $prep $rows $chunk $cap <<'P'
import json, sys
try:
    d = json.loads(open("gpurun_out/s10_b%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("prep", sys.argv[2], "rows1", sys.argv[3], "chunk", sys.argv[4], "cap", sys.argv[5], "| ms", round(d["ms_per_step"], 2),
          "e2e", round(d["e2e"]["ms_per_step"], 2), "proofs", round(d["e2e_from_proofs"]["ms_per_step"], 2))
except Exception as e:
    print("run", sys.argv[1:], "failed", e)
P
done <<'L'
0 12 21 64
1 12 21 64
1 15 21 64
1 14 21 64
1 16 21 64
1 15 20 64
1 15 19 64
1 15 21 0
1 15 21 100
0 15 21 64
L
