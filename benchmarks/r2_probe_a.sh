#!/bin/bash
# round-2 probe A: DMMA statistics / Newton-Schulz kernels (tests, timings, ncu) + conv_gemm WMODE diagnosis
O=gpurun_out/r2a; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt; cat $O/pytest.txt
timeout 240 python benchmarks/fp64_kernels.py > $O/fp64.json 2> $O/fp64.err; tail -c 400 $O/fp64.err
timeout 100 python benchmarks/scoring.py --mode indiv > $O/indiv.json 2> $O/indiv.err
timeout 100 python benchmarks/scoring.py --mode inf > $O/inf.json 2> $O/inf.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stats_dmma_kernel -c 8 -o $O/ncu_stats python benchmarks/fp64_kernels.py --reps 1 --what stats > $O/ncu_stats.log 2>&1
bash benchmarks/ncu_export.sh $O/ncu_stats.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -s 8 -c 4 -o $O/ncu_ns768 python benchmarks/fp64_kernels.py --reps 1 --what frechet --dims 768 > $O/ncu_ns768.log 2>&1
bash benchmarks/ncu_export.sh $O/ncu_ns768.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dgemm_strided -s 8 -c 3 -o $O/ncu_nsb python benchmarks/fp64_kernels.py --reps 1 --what batched --songs 2000 > $O/ncu_nsb.log 2>&1
bash benchmarks/ncu_export.sh $O/ncu_nsb.ncu-rep
FADTK_WLO=fp8 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 8 -o $O/ncu_wlo8 python bench.py --clips 1000 --baseline-clips 1000 --chunk-clips 1000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_wlo8.log 2>&1
bash benchmarks/ncu_export.sh $O/ncu_wlo8.ncu-rep
timeout 120 python bench.py --steps 5 --no-cpu-baseline --no-e2e > $O/bench_wlo16.json 2> $O/bench_wlo16.err
FADTK_WLO=fp8 timeout 120 python bench.py --steps 5 --no-cpu-baseline --no-e2e > $O/bench_wlo8.json 2> $O/bench_wlo8.err
du -sh $O; ls -la $O
