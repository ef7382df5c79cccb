#!/bin/bash
# round-2 probe W (2 GPUs): the final library under torchrun - weak + strong scaling line, reference arm (rank 0 only), C-ABI all-reduce check
O=gpurun_out/r2w; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $T bench.py --gpus 2 --no-cpu-baseline > $O/bench_vggish_n2.json 2> $O/bench_vggish_n2.err; tail -c 300 $O/bench_vggish_n2.err
timeout 300 $T bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_reference_n2.json 2> $O/bench_reference_n2.err
timeout 200 $T tests/multi_gpu_allreduce_check.py > $O/allreduce_check.txt 2>&1; tail -3 $O/allreduce_check.txt
python - <<'PY'
import json
for f in ("bench_vggish_n2","bench_reference_n2"):
    try:
        j=json.loads(open(f"gpurun_out/r2w/{f}.json").read().strip().splitlines()[-1])
        print(f, j.get("n_gpus"), round(j["value"]), round(j["ms_per_step"],1), (j.get("e2e") or {}).get("value"), j.get("strong_scaling"))
    except Exception as e: print(f, "ERR", e)
PY
