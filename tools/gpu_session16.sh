#!/bin/bash
# round 2, GPU session 16: state of HEAD -- GPU suite, full bench line, reference arm, from-proofs timeline, ncu --set full of the
# dominant kernels with the pipe counters, ncu launch list of the bench command
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/s16_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s16_pytest.log
tail -3 $O/s16_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/s16_bench.json 2> $O/s16_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/s16_ref.json 2> $O/s16_ref.err; echo "ref rc=$?"; cut -c1-300 $O/s16_ref.json
python tools/bv_timeline.py > $O/s16_timeline.txt 2>&1; grep "wall" $O/s16_timeline.txt
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_ingest2|k_accumulate$" -s 6 -c 3 -f -o $O/s16_ingest \
  python bench.py --steps 1 --warmup 3 --no-proofs-leg --no-configs > $O/s16_ncu_ingest.log 2>&1
ncu -i $O/s16_ingest.ncu-rep --page raw --csv > $O/s16_ingest_raw.csv 2>/dev/null
rm -f $O/s16_ingest.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/s16_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-proofs-leg --no-configs > /dev/null 2>&1
wc -l $O/s16_launches.csv
python - <<'P'
import json
d = json.loads(open("gpurun_out/s16_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "proofs", d["e2e_from_proofs"]["ms_per_step"],
      d["roofline"]["kernel_ms_each"], d["roofline"]["integer_pipe"]["k_ingest2_frac_of_calibrated"],
      d["roofline"]["integer_pipe"]["k_accumulate_frac_of_calibrated"])
print("prove", d["configs"]["cmz_prove"]["ms_per_call"], "dleq", json.dumps(d["configs"]["dleq_batch_verify"])[:400])
print("sweep", json.dumps(d["configs"]["raw_msm_sweep"])[:600])
P
