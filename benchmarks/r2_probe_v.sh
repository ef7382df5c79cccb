#!/bin/bash
# round-2 probe V: are single launches of the GEMM kernel ever slow?  per-launch timings (30 per shape and mode), the shipped
# wait loop (try_wait with a suspend hint) against the same loop without the hint, two passes each on one box; attention spill fix
O=gpurun_out/r2v; mkdir -p $O
V=$PWD/fadtk_b200/csrc/variant_nohint.so
for pass in 1 2; do
  LINEAR_SHAPES_REPS=30 timeout 200 python benchmarks/linear_shapes.py 2>/dev/null | tail -1 > $O/shapes_hint_$pass.json
  LINEAR_SHAPES_REPS=30 FADTK_B200_LIB=$V timeout 200 python benchmarks/linear_shapes.py 2>/dev/null | tail -1 > $O/shapes_nohint_$pass.json
done
python - <<'PY'
import json
for f in ("shapes_hint_1","shapes_nohint_1","shapes_hint_2","shapes_nohint_2"):
    print(f)
    for r in json.load(open(f"gpurun_out/r2v/{f}.json")):
        print("   ", r["layer"][:28].ljust(28), " ".join(f"{k}:{r[k]['min_ms']:.3f}/{r[k]['ms']:.3f}/{r[k]['max_ms']:.3f}" for k in ("split","split_pair","fp16")))
PY
B="--no-cpu-baseline --no-e2e --no-strong --files-clips 0"
for m in whisper-small w2v2-base; do
  timeout 200 python bench.py --model $m $B > $O/bench_$m.json 2>/dev/null
  FADTK_B200_LIB=$V timeout 200 python bench.py --model $m $B > $O/bench_${m}_nohint.json 2>/dev/null
done
timeout 200 python bench.py $B > $O/bench_vggish.json 2>/dev/null
FADTK_B200_LIB=$V timeout 200 python bench.py $B > $O/bench_vggish_nohint.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v/bench_*.json")):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], round(j["ms_per_step"],2), round(j["value"]))
PY
timeout 600 python -m pytest tests/test_whisper.py tests/test_w2v.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest.txt
