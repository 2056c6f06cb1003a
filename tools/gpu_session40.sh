#!/bin/bash
# round 2, GPU session 40: HEAD, closing check of the round -- smoke, GPU suite, the full bench line
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q > $O/s40_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s40_pytest.log
tail -3 $O/s40_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/s40_bench.json 2> $O/s40_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/s40_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "proofs", d["e2e_from_proofs"]["ms_per_step"],
      d["roofline"]["kernel_ms_each"], d["roofline"]["frac"], d["roofline"]["traffic_each"],
      d["roofline"]["integer_pipe"]["k_ingest2_frac_of_calibrated"], d["roofline"]["integer_pipe"]["k_accumulate_frac_of_calibrated"])
print("prove", d["configs"]["cmz_prove"]["ms_per_call"], d["configs"]["cmz_prove"]["integer_pipe"]["frac_of_calibrated"])
dl = d["configs"]["dleq_batch_verify"]; print("dleq", dl["ms_per_step"], dl["e2e"]["ms_per_step"], dl["e2e_from_proofs"]["ms_per_step"], dl["prove"]["ms_per_call"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
P
python - <<'P'
import json
d = json.loads(open("gpurun_out/s40_bench.json").read().strip().splitlines()[-1])
print([(r["log2_n"], round(r["gpu_ms"], 3)) for r in d["configs"]["raw_msm_sweep"]])
P
