#!/bin/bash
# round-2 probe I: attention v4 (two threads per query row)
O=gpurun_out/r2i; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_whisper.py tests/test_w2v.py -m gpu -q -k "attention or whisper or w2v" 2>&1 | tail -4 | tee $O/pytest_attn.txt
timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_umma.json 2> $O/bench_whisper_umma.err
FADTK_ATTN=legacy timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_legacy.json 2> $O/bench_whisper_legacy.err
timeout 200 python bench.py --model w2v2-base --no-cpu-baseline --no-e2e > $O/bench_w2v_umma.json 2> $O/bench_w2v_umma.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_umma -s 2 -c 2 -o $O/ncu_attn python bench.py --model whisper-small --clips 64 --baseline-clips 64 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-strong --files-clips 0 > $O/ncu_attn.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_attn.ncu-rep; rm -f $O/*.source.csv.gz
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), {k:(round(v['ms_total'],1),v['launches']) for k,v in j['roofline']['other_kernels'].items() if k.startswith('clap')})
    except Exception as e:
        print(f, "ERR", e)
PY
