#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" 2>&1 | tail -2 | tee $O/pytest_attn.txt
timeout 300 python benchmarks/attn_parity_sweep.py 5 2>/dev/null | tail -1 | tee $O/sweep_umma.json
FADTK_ATTN=legacy timeout 300 python benchmarks/attn_parity_sweep.py 5 2>/dev/null | tail -1 | tee $O/sweep_legacy.json
timeout 200 python bench.py --model whisper-small --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('whisper', j['ms_per_step'], j['value'], j['roofline']['other_kernels'].get('clap_attn'))"
