#!/bin/bash
# round-2 probe X (N GPUs, N = $1): the final library under torchrun - weak + strong scaling line
N=$1; O=gpurun_out/r2x; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 500 $T bench.py --gpus $N --no-cpu-baseline > $O/bench_vggish_n$N.json 2> $O/bench_vggish_n$N.err; tail -c 300 $O/bench_vggish_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
j=json.loads(open(f"gpurun_out/r2x/bench_vggish_n{n}.json").read().strip().splitlines()[-1])
print(j.get("n_gpus"), round(j["value"]), round(j["ms_per_step"],1), (j.get("e2e") or {}).get("value"), j.get("strong_scaling"))
PY
