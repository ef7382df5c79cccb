#!/bin/bash
# round-2 probe B: whole GPU suite on the new host code; CTA-pair (cta_group::2) conv_gemm: parity + speed
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_all.txt; cat $O/pytest_all.txt
FADTK_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q -k "umma_layer or vggish or parity or config0 or directory" 2>&1 | tail -15 > $O/pytest_pair.txt; cat $O/pytest_pair.txt
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.err
FADTK_PAIR=1 timeout 200 python bench.py --no-cpu-baseline --files-clips 0 --steps 5 > $O/bench_pair.json 2> $O/bench_pair.err; tail -c 300 $O/bench_pair.err
FADTK_PAIR=1 FADTK_WLO=fp8 timeout 200 python bench.py --no-cpu-baseline --files-clips 0 --steps 5 > $O/bench_pair_fp8.json 2> $O/bench_pair_fp8.err; tail -c 300 $O/bench_pair_fp8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), j["roofline"]["frac"], (j.get("e2e") or {}).get("value"), (j.get("e2e_fused") or {}).get("value"), (j.get("e2e_files") or {}), (j.get("parity_sample") or {}).get("rel_err"), j["clocks"])
        print({k: round(v["ms_per_launch"],3) for k,v in j["roofline"]["per_layer"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh $O
