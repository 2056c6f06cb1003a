#!/bin/bash
# round 2, GPU session 14: affine Niels points in 128-byte slots (one line per gathered point)
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/s14_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s14_pytest.log
tail -3 $O/s14_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-proofs-leg > $O/s14_b1.json 2> $O/s14_b1.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/s14_b1.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["kernel_ms_each"],
      d["roofline"]["integer_pipe"]["k_accumulate_ms"], d["roofline"]["stage_ms_unfused_profile_mode"])
P
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:"k_accumulate$|k_ingest2" -s 9 -c 3 --csv --log-file $O/s14_ncu.csv \
  python bench.py --steps 1 --warmup 3 --no-configs --no-proofs-leg > /dev/null 2>&1
grep -v "^==" $O/s14_ncu.csv | awk -F'","' '{print substr($5,1,24), $(NF-2), $(NF)}' | tail -9
