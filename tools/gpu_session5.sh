#!/bin/bash
# round 2, GPU session 5 (2 GPUs): NCCL single-verdict test at world size 2, bench.py under torchrun at N = 2
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/s5_gpus.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_toolbox.py -m gpu -q -x > $O/s5_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s5_pytest.log
tail -8 $O/s5_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > $O/s5_bench_n2.json 2> $O/s5_bench_n2.err; echo "bench n2 rc=$?"
tail -5 $O/s5_bench_n2.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/s5_bench_n2.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "from_proofs", d.get("e2e_from_proofs", {}).get("ms_per_step"))
    for k, v in d.get("configs", {}).items():
        print(k, json.dumps(v)[:600])
except Exception as e:
    print("bench parse failed", e)
P
timeout 600 python bench.py --steps 5 --warmup 3 > $O/s5_bench_n1.json 2> $O/s5_bench_n1.err; echo "bench n1 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s5_ref.json 2> $O/s5_ref.err; echo "ref rc=$?"; cut -c1-400 $O/s5_ref.json
du -sh $O
