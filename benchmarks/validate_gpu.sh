#!/bin/bash
# Full GPU validation of one box (end of round 2): parity tests, smoke, the bench line of every BASELINE configuration,
# the reference arm, scoring, per-shape GEMM rates, ncu of the final VGGish kernels.  Outputs under gpurun_out/final/.
# Usage (from the repo root, on a B200):  bash benchmarks/validate_gpu.sh
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "FAD gpu|rel |passed|failed|error|Error" | tail -30 > $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 400 python bench.py > $O/bench_vggish.json 2> $O/bench_vggish.err; tail -c 200 $O/bench_vggish.err
timeout 300 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 300 python bench.py --model clap-laion-audio --clips 6250 --no-cpu-baseline > $O/bench_clap_6250.json 2> $O/bench_clap_6250.err
timeout 300 python bench.py --model encodec-emb --clips 1250 --indiv --no-cpu-baseline > $O/bench_encodec_1250_indiv.json 2> $O/bench_encodec_1250_indiv.err
timeout 400 python bench.py --model whisper-small --clips 3125 --inf --no-cpu-baseline > $O/bench_whisper_3125_inf.json 2> $O/bench_whisper_3125_inf.err
for m in clap-laion-audio clap-laion-music whisper-small encodec-emb w2v2-base; do
    timeout 200 python bench.py --model $m --no-cpu-baseline > $O/bench_$m.json 2> $O/bench_$m.err
done
timeout 100 python benchmarks/scoring.py --mode indiv > $O/scoring_indiv.json 2> $O/scoring_indiv.err
timeout 100 python benchmarks/scoring.py --mode inf > $O/scoring_inf.json 2> $O/scoring_inf.err
timeout 200 python benchmarks/linear_shapes.py 2> $O/linear_shapes.err | tail -1 > $O/linear_shapes.json
timeout 400 python benchmarks/parity_large.py --model encodec-emb --clips 100 2>/dev/null | tail -1 > $O/parity_encodec_100.json
NCU="ncu --set full --clock-control none --import-source on"
B="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-strong --files-clips 0"
timeout 250 $NCU -k regex:conv_gemm -c 8 -o $O/ncu_vggish python bench.py --clips 1000 --baseline-clips 1000 --chunk-clips 1000 $B > $O/ncu_vggish.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_vggish.ncu-rep; rm -f $O/ncu_vggish.source.csv.gz
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_vggish.csv python bench.py --clips 2000 --baseline-clips 1000 --chunk-clips 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong --files-clips 0 > $O/launches.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json") + glob.glob("gpurun_out/final/scoring_*.json") + glob.glob("gpurun_out/final/parity_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], j.get("value"), j.get("ms_per_step"), (j.get("roofline") or {}).get("frac"),
              "e2e", (j.get("e2e") or {}).get("value"), "parity", (j.get("parity_sample") or {}).get("rel_err"), j.get("rel_err"))
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh $O
