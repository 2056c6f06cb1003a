#!/bin/bash
# round 2, GPU session 13: shifting digit extractor + batched scatter atomics in k_ingest2; L2 fetch granularity
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s13_pytest.log
tail -3 $O/s13_pytest.log
i=0
for o in "--scatter-batch 1" "--scatter-batch 0" "--scatter-batch 1 --l2-fetch 32" "--scatter-batch 1 --l2-fetch 64" "--scatter-batch 1 --l2-fetch 128"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg $o > $O/s13_b$i.json 2> $O/s13_b$i.err
  python - $i "$o" <<'P'
import json, sys
try:
    d = json.loads(open("gpurun_out/s13_b%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "| ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"],
          d["roofline"]["integer_pipe"]["k_accumulate_ms"], d["roofline"]["stage_ms_unfused_profile_mode"])
except Exception as e:
    print("run", sys.argv[1:], "failed", e)
P
done
for g in 32 64 128; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"k_accumulate$|k_ingest2" -s 9 -c 3 --csv --log-file $O/s13_ncu_l2_$g.csv \
    python bench.py --steps 1 --warmup 3 --no-configs --no-proofs-leg --l2-fetch $g > /dev/null 2>&1
  grep -v "^==" $O/s13_ncu_l2_$g.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tail -9
done
