#!/bin/bash
# round 2, GPU session 1: parity tests (comb prover, headline windows), prover timings, ncu --set full of the prover
# kernels and of k_ingest2 / k_accumulate with the pipe and stall counters.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_smi.txt 2>&1
nproc >> $O/s1_smi.txt
timeout 900 python -m pytest tests -m gpu -q > $O/s1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s1_pytest.log
tail -5 $O/s1_pytest.log
timeout 600 python tools/bench_configs.py --out $O/s1_configs_straus.json --sweep-max 8 --log2-dleq 12 > $O/s1_configs_straus.log 2>&1
timeout 600 python tools/bench_configs.py --prove-comb --out $O/s1_configs_comb.json --sweep-max 8 --log2-dleq 12 > $O/s1_configs_comb.log 2>&1
grep -h "config1" $O/s1_configs_straus.log $O/s1_configs_comb.log | cut -c1-400
EXTRA=smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_small_msm_ct|k_small_msm_comb|k_build_tables|k_build_combs|k_comb_recode" -f -o $O/s1_prover \
  python tools/ncu_prover.py --log2 14 > $O/s1_ncu_prover.log 2>&1
ncu -i $O/s1_prover.ncu-rep --page raw --csv > $O/s1_prover_raw.csv 2>/dev/null
ncu -i $O/s1_prover.ncu-rep --page details --csv > $O/s1_prover_details.csv 2>/dev/null
ls -la $O/s1_prover.ncu-rep
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on \
  -k regex:"k_ingest2|k_accumulate$" -s 6 -c 3 -f -o $O/s1_ingest \
  python bench.py --steps 1 --warmup 3 --no-proofs-leg > $O/s1_ncu_ingest.log 2>&1
ncu -i $O/s1_ingest.ncu-rep --page raw --csv > $O/s1_ingest_raw.csv 2>/dev/null
ncu -i $O/s1_ingest.ncu-rep --page details --csv > $O/s1_ingest_details.csv 2>/dev/null
ncu -i $O/s1_ingest.ncu-rep --page source --csv -k regex:"k_ingest2" --launch-skip 0 --launch-count 1 > $O/s1_ingest_source.csv 2>/dev/null
ls -la $O/s1_ingest.ncu-rep
# keep the merged output small: the reports themselves only if they are small
for f in $O/s1_prover.ncu-rep $O/s1_ingest.ncu-rep; do
  if [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ]; then rm -f $f; fi
done
timeout 600 python bench.py --steps 5 --warmup 3 > $O/s1_bench.json 2> $O/s1_bench.err
cut -c1-600 $O/s1_bench.json
du -sh $O
