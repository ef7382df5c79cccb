#!/bin/bash
# round-2 probe C: pair=auto default, grouped fp8 low-part MMAs, pipelined plugin forward, BASELINE configs 2-4 lines
O=gpurun_out/r2c; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -6 > $O/pytest.txt; cat $O/pytest.txt
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 400 $O/bench_default.err
for g in 2 3; do
FADTK_WLO=fp8 FADTK_LO8_GROUP=$g timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_fp8_g$g.json 2> $O/bench_fp8_g$g.err
FADTK_PAIR=1 FADTK_WLO=fp8 FADTK_LO8_GROUP=$g timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_fp8_pairall_g$g.json 2> $O/bench_fp8_pairall_g$g.err
done
FADTK_PAIR=0 FADTK_WLO=fp8 FADTK_LO8_GROUP=2 timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_fp8_nopair_g2.json 2> $O/bench_fp8_nopair_g2.err
timeout 200 python benchmarks/profile_file_flow.py 4000 > $O/file_flow_profile.txt 2>&1
timeout 300 python bench.py --model clap-laion-audio --clips 6250 --no-cpu-baseline > $O/bench_clap_6250.json 2> $O/bench_clap_6250.err; tail -c 300 $O/bench_clap_6250.err
timeout 300 python bench.py --model encodec-emb --clips 1250 --indiv --no-cpu-baseline > $O/bench_encodec_1250_indiv.json 2> $O/bench_encodec_1250_indiv.err; tail -c 300 $O/bench_encodec_1250_indiv.err
timeout 400 python bench.py --model whisper-small --clips 3125 --inf --no-cpu-baseline > $O/bench_whisper_3125_inf.json 2> $O/bench_whisper_3125_inf.err; tail -c 300 $O/bench_whisper_3125_inf.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), (j.get("e2e") or {}).get("value"), (j.get("e2e_fused") or {}).get("value"), (j.get("e2e_files") or {}).get("value"), (j.get("parity_sample") or {}).get("rel_err"), j["clocks"]["sm_mhz"], j.get("scoring"))
        if j["roofline"].get("per_layer"): print({k: round(v["ms_per_launch"],3) for k,v in j["roofline"]["per_layer"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
head -60 $O/file_flow_profile.txt
du -sh $O
