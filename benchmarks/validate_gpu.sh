#!/bin/bash
# Full GPU validation of one box: parity tests, smoke, every bench line.  Outputs under gpurun_out/.
# Usage (from the repo root, on a B200):  bash benchmarks/validate_gpu.sh
mkdir -p gpurun_out
timeout 430 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
timeout 80 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 150 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
timeout 170 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2>/dev/null
timeout 70 python bench.py --model w2v2-base --no-cpu-baseline > gpurun_out/bench_w2v2.json 2>/dev/null
timeout 60 python benchmarks/scoring.py --mode indiv > gpurun_out/scoring_indiv.json 2>/dev/null
timeout 60 python benchmarks/scoring.py --mode inf > gpurun_out/scoring_inf.json 2>/dev/null
for m in encodec-emb clap-laion-audio whisper-small clap-laion-music; do
    timeout 80 python bench.py --model $m > gpurun_out/bench_$m.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench*.json") + glob.glob("gpurun_out/bench_*.json") + glob.glob("gpurun_out/scoring_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j.get("value"), j.get("ms_per_step"), (j.get("roofline") or {}).get("frac"),
              (j.get("e2e") or {}).get("value"), (j.get("parity_sample") or {}).get("rel_err"))
    except Exception as e:
        print(f, "ERR", e)
PY
