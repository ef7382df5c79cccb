#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 400 python -m pytest tests/test_w2v.py -m gpu -q -s -k "parity" 2>&1 | grep -E "FAD gpu|passed|failed" | tee $O/parity_umma.txt
FADTK_ATTN=legacy timeout 400 python -m pytest tests/test_w2v.py -m gpu -q -s -k "parity" 2>&1 | grep -E "FAD gpu|passed|failed" | tee $O/parity_legacy.txt
