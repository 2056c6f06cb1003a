#!/bin/bash
# round 2, GPU session 4: fused scan + wave-aligned slices of the prover; small-path defaults
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/s4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s4_pytest.log
tail -8 $O/s4_pytest.log
timeout 600 python tools/bench_prove.py --quick --out $O/s4_prove.json > $O/s4_prove.log 2>&1; tail -12 $O/s4_prove.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s4_prove_launches.csv \
  python tools/ncu_prover.py --log2 16 --comb-only --cta-only > $O/s4_ncu_launches.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s4_prove_launches.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if v > 50000 and ("k_pv" in r[ki] or "comb" in r[ki] or "compress_limbs_sh" in r[ki]):
        print(r[ki][:60], round(v / 1e6, 3), "ms")
P
du -sh $O
