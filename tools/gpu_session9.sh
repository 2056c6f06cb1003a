#!/bin/bash
# round 2, GPU session 9: decoder's inverse square root started from v itself (2 S + 3 M fewer per point), slab rows of the
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
# from-proofs path as one 2-D launch (bv_merge_rows): whole GPU suite, field rates, full bench line, merge on/off
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s9_pytest.log
tail -4 $O/s9_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/s9_bench.json 2> $O/s9_bench.err; echo "bench rc=$?"
for mr in 0 1; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --bv-merge-rows $mr > $O/s9_bench_mr$mr.json 2> $O/s9_bench_mr$mr.err
done
python - <<'P'
import json
def L(p): return json.loads(open(p).read().strip().splitlines()[-1])
d = L("gpurun_out/s9_bench.json")
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "proofs", d["e2e_from_proofs"]["ms_per_step"],
      d["roofline"]["kernel_ms_each"], d["roofline"]["integer_pipe"]["k_ingest2_frac_of_calibrated"],
      d["roofline"]["integer_pipe"]["k_accumulate_ms"])
print("prove", d["configs"]["cmz_prove"]["ms_per_call"], "dleq", d["configs"]["dleq_batch_verify"]["ms_per_step"])
for mr in (0, 1):
    e = L("gpurun_out/s9_bench_mr%d.json" % mr)
    print("merge_rows", mr, "ms", e["ms_per_step"], "e2e", e["e2e"]["ms_per_step"], "proofs", e["e2e_from_proofs"]["ms_per_step"])
P
