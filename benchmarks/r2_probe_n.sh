#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 300 python benchmarks/w2v_layer_bias.py 2>/dev/null | tail -1 | tee $O/w2v_layer_bias.json
timeout 600 python benchmarks/parity_large.py --model whisper-small --clips 40 2>/dev/null | tail -1 | tee $O/parity_whisper_small_40.json
timeout 400 python benchmarks/parity_large.py --model encodec-emb --clips 100 2>/dev/null | tail -1 | tee $O/parity_encodec_100.json
