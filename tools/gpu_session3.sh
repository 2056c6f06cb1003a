#!/bin/bash
# round 2, GPU session 3: the pipelined prover, register-resident shared combs (shuffle lookup), warp stagger sweep
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/s3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s3_pytest.log
tail -15 $O/s3_pytest.log
timeout 600 python tools/bench_prove.py --out $O/s3_prove.json > $O/s3_prove.log 2>&1; tail -25 $O/s3_prove.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s3_prove_launches.csv \
  python tools/ncu_prover.py --log2 16 --comb-only --cta-only > $O/s3_ncu_launches.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s3_prove_launches.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if v > 50000:
        print(r[ki][:60], round(v / 1e6, 3), "ms")
P
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"k_comb_msm_cta" -c 1 -f -o $O/s3_cta \
  python tools/ncu_prover.py --log2 14 --comb-only --cta-only > $O/s3_ncu_cta.log 2>&1
ncu -i $O/s3_cta.ncu-rep --page raw --csv > $O/s3_cta_raw.csv 2>/dev/null
rm -f $O/s3_cta.ncu-rep
du -sh $O
