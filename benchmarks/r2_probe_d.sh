#!/bin/bash
# round-2 probe D: tcgen05 issue-pattern microbenchmark, e2e after the host-path fixes, large-scale parity, ncu evidence
O=gpurun_out/r2d; mkdir -p $O
timeout 120 python benchmarks/umma_modes.py > $O/umma_modes.json 2> $O/umma_modes.err; cat $O/umma_modes.json; tail -c 300 $O/umma_modes.err
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
timeout 400 python benchmarks/parity_large.py --model vggish --clips 1000 > $O/parity_vggish_1000.json 2> $O/parity_vggish_1000.err; cat $O/parity_vggish_1000.json
timeout 500 python benchmarks/parity_large.py --model clap-laion-audio --clips 200 > $O/parity_clap_200.json 2> $O/parity_clap_200.err; cat $O/parity_clap_200.json
NCU="ncu --set full --clock-control none --import-source on"
B="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-strong --files-clips 0"
timeout 250 $NCU -k regex:conv_gemm -c 8 -o $O/ncu_vggish_pair python bench.py --clips 1000 --baseline-clips 1000 --chunk-clips 1000 $B > $O/ncu_vggish_pair.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_vggish_pair.ncu-rep
timeout 250 $NCU -k regex:'logmel_kernel|conv1_kernel' -c 2 -o $O/ncu_vggish_front python bench.py --clips 1000 --baseline-clips 1000 --chunk-clips 1000 $B > $O/ncu_vggish_front.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_vggish_front.ncu-rep
timeout 250 $NCU -k regex:'whisper_flash_attention|whisper_logmel|whisper_cross' -c 3 -o $O/ncu_whisper python bench.py --model whisper-small --clips 64 --baseline-clips 64 $B > $O/ncu_whisper.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_whisper.ncu-rep
timeout 250 $NCU -k regex:'lstm_cell|encodec_im2col' -c 4 -o $O/ncu_encodec python bench.py --model encodec-emb --clips 128 --baseline-clips 64 --chunk-clips 128 $B > $O/ncu_encodec.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_encodec.ncu-rep
timeout 250 $NCU -k regex:'w2v_conv0_apply|w2v_posconv|w2v_normalize' -c 3 -o $O/ncu_w2v python bench.py --model w2v2-base --clips 32 --baseline-clips 32 $B > $O/ncu_w2v.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_w2v.ncu-rep
timeout 250 $NCU -k regex:'clap_window_attention|clap_ln_kernel' -c 4 -o $O/ncu_clap python bench.py --model clap-laion-audio --clips 50 --baseline-clips 50 $B > $O/ncu_clap.log 2>&1; bash benchmarks/ncu_export.sh $O/ncu_clap.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_vggish.csv python bench.py --clips 2000 --baseline-clips 1000 --chunk-clips 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong --files-clips 0 > $O/launches.log 2>&1
rm -f $O/*.source.csv.gz
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2d/bench_default.json").read().strip().splitlines()[-1])
print("default", round(j["ms_per_step"],1), round(j["value"]), round(j["roofline"]["frac"],4), "e2e", j["e2e"]["value"], "fused", j["e2e_fused"]["value"], "files", j["e2e_files"], "parity", j["parity_sample"], "cpu", j["cpu_baseline"])
PY
du -sh $O; ls $O
