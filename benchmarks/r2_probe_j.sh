#!/bin/bash
# round-2 probe J (2 GPUs): C-ABI NCCL all-reduce, strong-scaling record, reference arm under torchrun, --indiv over 2 ranks
O=gpurun_out/r2j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tests/multi_gpu_allreduce_check.py > $O/allreduce_check.txt 2>&1; grep -E "rank|Error|error" $O/allreduce_check.txt | head
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 300 $O/bench_n2.err
timeout 300 $TR --master-port 29543 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/ref_n1.json 2> $O/ref_n1.err
mkdir -p /tmp/mg && python tests/multi_gpu_scoring_check.py prepare /tmp/mg
FADTK_SYNTHETIC=1 timeout 200 python -m fadtk_b200 vggish /tmp/mg/base.npz /tmp/mg/ev /tmp/mg/one.csv --indiv > $O/indiv_one.log 2>&1
FADTK_SYNTHETIC=1 timeout 200 $TR --master-port 29544 -m fadtk_b200 vggish /tmp/mg/base.npz /tmp/mg/ev /tmp/mg/two.csv --indiv > $O/indiv_two.log 2>&1
python tests/multi_gpu_scoring_check.py compare /tmp/mg 2>&1 | tee $O/indiv_compare.txt
python - <<'PY'
import json
for f in ("bench_n2","ref_n2","ref_n1"):
    try:
        j=json.loads(open(f"gpurun_out/r2j/{f}.json").read().strip().splitlines()[-1])
        print(f, j.get("n_gpus"), round(j["value"]), j.get("ms_per_step"), (j.get("cpu_baseline") or {}).get("cores"), (j.get("cpu_baseline") or {}).get("host"), j.get("strong_scaling"), (j.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
