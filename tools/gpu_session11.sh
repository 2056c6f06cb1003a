#!/bin/bash
# round 2, GPU session 11: ncu --set full of the front-end kernel of zkp_batch_verify_proofs (k_bv_prepare2)
set -u
O=gpurun_out
mkdir -p $O
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_bv_prepare2" -s 3 -c 1 -f -o $O/s11_prep \
  python tools/bv_timeline.py --log2-proofs 18 --opt bv_prep_stream=0 --opt bv_chunk_terms=1073741824 > $O/s11_ncu.log 2>&1
tail -3 $O/s11_ncu.log
ncu -i $O/s11_prep.ncu-rep --page raw --csv > $O/s11_prep_raw.csv 2>/dev/null
ncu -i $O/s11_prep.ncu-rep --page source --csv > $O/s11_prep_source.csv 2>/dev/null
ls -la $O/s11_prep.ncu-rep
if [ $(stat -c %s $O/s11_prep.ncu-rep) -gt 25000000 ]; then rm -f $O/s11_prep.ncu-rep; fi
