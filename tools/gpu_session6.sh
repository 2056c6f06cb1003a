#!/bin/bash
# round 2, GPU session 6: carry-folding field multiplication / squaring against the round-1 schedule
set -u
O=gpurun_out
mkdir -p $O
python - > $O/s6_field.txt 2>&1 <<'P'
import sys
sys.path.insert(0, ".")
from zkp_b200 import Engine
eng = Engine(0)
names = {9: "fe_sq  (vt tails, carry-folding)", 17: "fe_sq  (vt tails, round-1 schedule)", 8: "fe_mul (vt tails, carry-folding)",
         18: "fe_mul (vt tails, round-1 schedule)", 1: "fe_sq  (ct tails, carry-folding)", 0: "fe_mul (ct tails, carry-folding)",
         11: "mixed addition (vt)", 10: "mixed addition (ct)"}
for rep in range(2):
    for k in (9, 17, 8, 18, 1, 0, 11, 10):
        print("%-40s %.4e ops/s" % (names[k], eng.bench_field(k, 4096 if k < 10 or k > 11 else 1024)), flush=True)
P
cat $O/s6_field.txt
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_toolbox.py -m gpu -q -x > $O/s6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s6_pytest.log
tail -4 $O/s6_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg > $O/s6_bench.json 2> $O/s6_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/s6_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_each"], d["roofline"]["integer_pipe"]["k_accumulate_ms"])
P
