#!/bin/bash
# round 2, GPU session 27: HEAD (final of the round) -- smoke, GPU suite, the full bench line
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q > $O/s27_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s27_pytest.log
tail -3 $O/s27_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/s27_bench.json 2> $O/s27_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/s27_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "proofs", d["e2e_from_proofs"]["ms_per_step"],
      d["roofline"]["kernel_ms_each"], d["roofline"]["frac"], d["roofline"]["traffic_each"],
      d["roofline"]["integer_pipe"]["k_ingest2_frac_of_calibrated"], d["roofline"]["integer_pipe"]["k_accumulate_frac_of_calibrated"])
print("prove", d["configs"]["cmz_prove"]["ms_per_call"], d["configs"]["cmz_prove"]["integer_pipe"]["frac_of_calibrated"])
dl = d["configs"]["dleq_batch_verify"]; print("dleq", dl["ms_per_step"], dl["e2e"]["ms_per_step"], dl["e2e_from_proofs"]["ms_per_step"], dl["prove"]["ms_per_call"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
P
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/s27_ref.json 2> $O/s27_ref.err; echo "ref rc=$?"; cut -c1-200 $O/s27_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/s27_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-proofs-leg --no-configs > /dev/null 2>&1
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_ingest2|k_accumulate" -s 6 -c 3 -f -o $O/s27_ingest \
  python bench.py --steps 1 --warmup 3 --no-proofs-leg --no-configs > $O/s27_ncu_ingest.log 2>&1
ncu -i $O/s27_ingest.ncu-rep --page raw --csv > $O/s27_ingest_raw.csv 2>/dev/null
rm -f $O/s27_ingest.ncu-rep
