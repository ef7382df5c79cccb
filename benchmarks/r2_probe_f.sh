#!/bin/bash
# round-2 probe F: stacked N = 256 hi|lo MMA in conv_gemm (all models), attention v2 (2 CTAs per SM), mirrored statistics in CUDA
O=gpurun_out/r2f; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_all.txt; cat $O/pytest_all.txt
FADTK_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "umma_layer" 2>&1 | tail -4 > $O/pytest_pairall.txt; cat $O/pytest_pairall.txt
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
FADTK_PAIR=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_pair_all.json 2> $O/bench_pair_all.err
FADTK_PAIR=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_nopair.json 2> $O/bench_nopair.err
timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_umma.json 2> $O/bench_whisper_umma.err
FADTK_ATTN=legacy timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_legacy.json 2> $O/bench_whisper_legacy.err
timeout 200 python bench.py --model w2v2-base --no-cpu-baseline --no-e2e > $O/bench_w2v_umma.json 2> $O/bench_w2v_umma.err
timeout 200 python bench.py --model clap-laion-audio --no-cpu-baseline --no-e2e > $O/bench_clap.json 2> $O/bench_clap.err
timeout 200 python bench.py --model encodec-emb --no-cpu-baseline --no-e2e > $O/bench_encodec.json 2> $O/bench_encodec.err
timeout 100 python benchmarks/scoring.py --mode indiv > $O/indiv.json 2> $O/indiv.err; cat $O/indiv.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), (j.get("e2e") or {}).get("value"), j["clocks"]["sm_mhz"], (j.get("parity_sample") or {}).get("rel_err"), {k:(round(v['ms_total'],1),v['launches']) for k,v in j['roofline']['other_kernels'].items() if k.startswith('clap')})
        if j["roofline"].get("per_layer"): print({k: round(v["ms_per_launch"],3) for k,v in j["roofline"]["per_layer"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh $O
