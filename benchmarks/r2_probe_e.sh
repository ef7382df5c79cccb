#!/bin/bash
# round-2 probe E: tcgen05 attention (stage test, model tests, whisper / w2v bench), umma microbenchmark debug
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "encoder_attention" 2>&1 | tail -12 > $O/pytest_attn.txt; cat $O/pytest_attn.txt
timeout 400 python -m pytest tests/test_whisper.py tests/test_w2v.py -m gpu -q 2>&1 | tail -8 > $O/pytest_models.txt; cat $O/pytest_models.txt
timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_umma.json 2> $O/bench_whisper_umma.err; tail -c 300 $O/bench_whisper_umma.err
FADTK_ATTN=legacy timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e > $O/bench_whisper_legacy.json 2> $O/bench_whisper_legacy.err
timeout 200 python bench.py --model w2v2-base --no-cpu-baseline --no-e2e > $O/bench_w2v_umma.json 2> $O/bench_w2v_umma.err
FADTK_ATTN=legacy timeout 200 python bench.py --model w2v2-base --no-cpu-baseline --no-e2e > $O/bench_w2v_legacy.json 2> $O/bench_w2v_legacy.err
for m in 5 0 1 4 2 3; do timeout 60 python benchmarks/umma_modes.py --modes $m --ksteps 100000 > $O/umma_mode$m.json 2> $O/umma_mode$m.err; echo "mode $m rc $?"; cat $O/umma_mode$m.json; tail -c 200 $O/umma_mode$m.err; done
timeout 120 compute-sanitizer --tool memcheck python benchmarks/umma_modes.py --modes 0 --ksteps 1600 > $O/sanitizer_mode0.txt 2>&1; grep -E "Invalid|Error|error|at |by thread" $O/sanitizer_mode0.txt | head -20
timeout 500 python benchmarks/parity_large.py --model clap-laion-audio --clips 200 > $O/parity_clap_200.json 2> $O/parity_clap_200.err; cat $O/parity_clap_200.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), {k:(round(v['ms_total'],1),v['launches']) for k,v in j['roofline']['other_kernels'].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh $O
