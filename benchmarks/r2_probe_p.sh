#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 300 python benchmarks/gemm_bias_probe.py 2>/dev/null | tail -1 | tee $O/gemm_bias_probe.json
timeout 600 python benchmarks/w2v_fad_terms.py 32 2>/dev/null | tail -1 | tee $O/w2v_fad_terms.json
