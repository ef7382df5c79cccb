#!/bin/bash
# round-2 probe Q: CTA pairs on the Linear layers (min K sweep), ncu of the tcgen05 attention kernel, parity with pairs on
O=gpurun_out/r2q; mkdir -p $O
B="--no-cpu-baseline --no-e2e --no-strong --files-clips 0"
run() { # name env model
  FADTK_PAIR_LINEAR=$2 timeout 300 python bench.py --model $3 $B > $O/bench_$1.json 2> $O/bench_$1.err
  python - "$O/bench_$1.json" "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(j["ms_per_step"],2), round(j["value"]), j.get("parity_sample"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run whisper_p0 0 whisper-small
run whisper_p512 512 whisper-small
run whisper_p2048 2048 whisper-small
run w2v_p0 0 w2v2-base
run w2v_p512 512 w2v2-base
run clap_p0 0 clap-laion-audio
run clap_p384 384 clap-laion-audio
run encodec_p0 0 encodec-emb
run encodec_p1024 1024 encodec-emb
FADTK_PAIR_LINEAR=192 timeout 900 python -m pytest tests/test_whisper.py tests/test_w2v.py tests/test_clap.py tests/test_encodec.py -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_pairs.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attention_umma -c 2 -o $O/ncu_attention_umma python bench.py --model whisper-small --clips 64 --baseline-clips 64 --steps 1 --warmup 0 $B > $O/ncu_attention_umma.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_attention_umma.ncu-rep
ls -la $O | head -40; du -sh $O
