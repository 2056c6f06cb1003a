#!/bin/bash
# round 2, GPU session 7: after the carry-folding field arithmetic: whole GPU suite, prover sweep, full bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s7_pytest.log
tail -4 $O/s7_pytest.log
timeout 600 python tools/bench_prove.py --quick --out $O/s7_prove.json > $O/s7_prove.log 2>&1; tail -6 $O/s7_prove.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/s7_bench.json 2> $O/s7_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/s7_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "proofs", d["e2e_from_proofs"]["ms_per_step"],
      d["roofline"]["kernel_ms_each"], d["roofline"]["integer_pipe"])
print("prove", d["configs"]["cmz_prove"]["ms_per_call"], "dleq", d["configs"]["dleq_batch_verify"]["ms_per_step"])
P
