#!/bin/bash
# round 2, GPU session 15: device scalar arithmetic on the multiply-add chains; front-end kernel ncu after the change
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/s15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s15_pytest.log
tail -3 $O/s15_pytest.log
python tools/bv_timeline.py 2>&1 | grep "wall ms\|prepare\[0\]\|copy\[0\]\|scan\|ingest2\|finish"
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_bv_prepare2" -s 3 -c 1 -f -o $O/s15_prep \
  python tools/bv_timeline.py --log2-proofs 18 --opt bv_prep_stream=0 --opt bv_chunk_terms=1073741824 > $O/s15_ncu.log 2>&1
ncu -i $O/s15_prep.ncu-rep --page raw --csv > $O/s15_prep_raw.csv 2>/dev/null
ncu -i $O/s15_prep.ncu-rep --page source --csv > $O/s15_prep_source.csv 2>/dev/null
rm -f $O/s15_prep.ncu-rep
timeout 600 python tools/bench_prove.py --quick --out $O/s15_prove.json > $O/s15_prove.log 2>&1; tail -3 $O/s15_prove.log
