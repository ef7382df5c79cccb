#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
B="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-strong --files-clips 0"
timeout 300 $NCU -k regex:conv_gemm -c 6 -o $O/ncu_clap_gemm python bench.py --model clap-laion-audio --clips 50 --baseline-clips 50 $B > $O/ncu_clap_gemm.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_clap_gemm.ncu-rep
ls -la $O; du -sh $O
