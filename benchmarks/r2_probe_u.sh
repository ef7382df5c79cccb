#!/bin/bash
# round-2 probe U: epilogue changes (bias requested ahead of the accumulator wait, staging by st.shared/ld.shared, no spilled loop scalars)
# against the previous commit's build on the same box
O=gpurun_out/r2u; mkdir -p $O
B="--no-cpu-baseline --no-e2e --no-strong --files-clips 0"
V=$PWD/fadtk_b200/csrc/variant_prev.so
run() { # name lib model
  FADTK_B200_LIB=$2 timeout 300 python bench.py --model $3 $B > $O/bench_$1.json 2> $O/bench_$1.err
  python - "$O/bench_$1.json" "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(j["ms_per_step"],2), round(j["value"]), j["roofline"].get("frac"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run clap_new "" clap-laion-audio
run clap_prev $V clap-laion-audio
run vggish_new "" vggish
run vggish_prev $V vggish
run whisper_new "" whisper-small
run whisper_prev $V whisper-small
run w2v_new "" w2v2-base
run w2v_prev $V w2v2-base
run encodec_new "" encodec-emb
run encodec_prev $V encodec-emb
timeout 300 python benchmarks/linear_shapes.py 2> $O/linear_shapes.err | tail -1 > $O/linear_shapes.json
python - <<'PY'
import json
for r in json.load(open("gpurun_out/r2u/linear_shapes.json")):
    print(r["layer"], r["K"], r["N"], {k: r[k]["ms"] for k in ("split", "split_pair", "fp16")})
PY
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "FAD gpu|passed|failed|error|Error" | tail -20 | tee $O/pytest.txt
