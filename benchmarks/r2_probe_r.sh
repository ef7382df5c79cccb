#!/bin/bash
# round-2 probe R: lean mbarrier wait loop vs the clock-watchdog loop (same box, two builds), attention without per-key predicates
O=gpurun_out/r2r; mkdir -p $O
B="--no-cpu-baseline --no-e2e --no-strong --files-clips 0"
V=$PWD/fadtk_b200/csrc/variant_clockwait.so
run() { # name lib model
  FADTK_B200_LIB=$2 timeout 300 python bench.py --model $3 $B > $O/bench_$1.json 2> $O/bench_$1.err
  python - "$O/bench_$1.json" "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(j["ms_per_step"],2), round(j["value"]), j["roofline"].get("frac"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run vggish_lean "" vggish
run vggish_clock $V vggish
run vggish_lean2 "" vggish
run whisper_lean "" whisper-small
run whisper_clock $V whisper-small
run clap_lean "" clap-laion-audio
run clap_clock $V clap-laion-audio
run w2v_lean "" w2v2-base
run w2v_clock $V w2v2-base
run encodec_lean "" encodec-emb
run encodec_clock $V encodec-emb
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "FAD gpu|passed|failed|error|Error" | tail -20 | tee $O/pytest.txt
