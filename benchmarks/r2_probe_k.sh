#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 200 python -m pytest tests/test_w2v.py tests/test_whisper.py -m gpu -q -s -k "parity" 2>&1 | grep -E "FAD gpu|passed|failed" | tee $O/parity_umma.txt
FADTK_ATTN=legacy timeout 200 python -m pytest tests/test_w2v.py tests/test_whisper.py -m gpu -q -s -k "parity" 2>&1 | grep -E "FAD gpu|passed|failed" | tee $O/parity_legacy.txt
timeout 200 python -m pytest tests/test_w2v.py -m gpu -q -s -k "hidden_states" 2>&1 | grep -E "rms rel|passed|failed" | tee $O/hidden_umma.txt
FADTK_ATTN=legacy timeout 200 python -m pytest tests/test_w2v.py -m gpu -q -s -k "hidden_states" 2>&1 | grep -E "rms rel|passed|failed" | tee $O/hidden_legacy.txt
