#!/bin/bash
# round-2 probe G: full GPU suite on the default build; same-box A/B of the stacked N = 256 variant
O=gpurun_out/r2g; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_all.txt; cat $O/pytest_all.txt
timeout 200 python -m pytest tests/test_w2v.py tests/test_whisper.py -m gpu -q -s -k "parity" 2>&1 | grep -E "FAD gpu|passed|failed" | tee $O/parity_models.txt
for rep in 1 2; do
timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_unstacked_$rep.json 2> $O/bench_unstacked_$rep.err
FADTK_STACK=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_stacked_$rep.json 2> $O/bench_stacked_$rep.err
done
FADTK_STACK=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "umma_layer or vggish" 2>&1 | tail -3 | tee $O/pytest_stacked.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), j["clocks"]["sm_mhz"], j["clocks"].get("power_w_median"))
        print({k: round(v["ms_per_launch"],3) for k,v in j["roofline"]["per_layer"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
