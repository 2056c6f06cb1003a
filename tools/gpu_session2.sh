#!/bin/bash
# round 2, GPU session 2: the CTA-staged comb prover, the small-MSM dispatch, the device-resident partial sums, bench.py
# with its `configs` block.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s2_pytest.log
tail -25 $O/s2_pytest.log
timeout 300 python tools/bench_small.py --out $O/s2_small.json > $O/s2_small.log 2>&1; tail -12 $O/s2_small.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/s2_bench.json 2> $O/s2_bench.err; echo "bench rc=$?"
tail -5 $O/s2_bench.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/s2_bench.json").read())
    print("value", d["value"], "e2e", d["e2e"]["value"], "from_proofs", d.get("e2e_from_proofs", {}).get("ms_per_step"))
    for k, v in d.get("configs", {}).items():
        print(k, json.dumps(v)[:700])
except Exception as e:
    print("bench parse failed", e)
P
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_comb_msm_cta|k_small_msm_comb|k_comb_recode_il|k_build_combs" -f -o $O/s2_prover \
  python tools/ncu_prover.py --log2 14 --comb-only > $O/s2_ncu_prover.log 2>&1
ncu -i $O/s2_prover.ncu-rep --page raw --csv > $O/s2_prover_raw.csv 2>/dev/null
ncu -i $O/s2_prover.ncu-rep --page source --csv -k regex:"k_comb_msm_cta" > $O/s2_cta_source.csv 2>/dev/null
ls -la $O/s2_prover.ncu-rep
for f in $O/s2_prover.ncu-rep; do
  if [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ]; then rm -f $f; fi
done
timeout 300 compute-sanitizer --tool memcheck python tools/ncu_prover.py --log2 8 --comb-only > $O/s2_memcheck.log 2>&1; tail -3 $O/s2_memcheck.log
du -sh $O
