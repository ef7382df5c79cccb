#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 300 python benchmarks/gemm_bias_probe.py 2>/dev/null | tail -1 | tee $O/gemm_bias_probe.json
timeout 300 python benchmarks/w2v_layer_bias.py 2>/dev/null | tail -1 | tee $O/w2v_layer_bias.json
timeout 600 python -m pytest tests/test_w2v.py -m gpu -q -s -k parity 2>&1 | tail -15 | tee $O/w2v_parity.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest.txt
timeout 400 python bench.py > $O/bench_vggish.json 2> $O/bench_vggish.err; cat $O/bench_vggish.json
