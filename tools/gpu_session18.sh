#!/bin/bash
# round 2, GPU session 18: Keccak on 32-bit halves (122 LOP3 + 58 SHF per round), bucket work items planned under phase 2;
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
# compute-sanitizer memcheck over the kernels this round touched
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s18_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s18_pytest.log
tail -3 $O/s18_pytest.log
for po in 1 0 1 0; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-configs --no-proofs-leg --plan-overlap $po > $O/s18_po$po.json 2> $O/s18_po$po.err
  python - $po <<'P'
import json, sys
d = json.loads(open("gpurun_out/s18_po%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("plan_overlap", sys.argv[1], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["roofline"]["kernel_ms_each"],
      d["roofline"]["integer_pipe"]["k_accumulate_ms"])
P
done
python tools/bv_timeline.py 2>&1 | grep "wall ms\|prepare\[0\]\|copy\[0\]\|scan\|ingest2\|finish"
timeout 600 python tools/bench_prove.py --quick --out $O/s18_prove.json > $O/s18_prove.log 2>&1; tail -3 $O/s18_prove.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_toolbox.py -m gpu -q -x \
  -k "slab_pipeline or front_end_matches_oracle or edge_shapes or device_merlin" > $O/s18_memcheck_toolbox.log 2>&1
tail -4 $O/s18_memcheck_toolbox.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -m gpu -q -x > $O/s18_memcheck_msm.log 2>&1
tail -4 $O/s18_memcheck_msm.log
