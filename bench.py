#!/usr/bin/env python
"""Benchmark of the hot path.  Default workload = BASELINE.json configs[1]: VGGish FAD on
10 000 x 10 s synthetic 16 kHz clips per GPU.  `--model clap-laion-audio` runs the configs[2]-style
workload (CLAP-LAION HTSAT-tiny, 10 s 48 kHz clips, 10 windows per clip) at a single-GPU size.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model M]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over the eval set: PCM16 -> front-end -> embedder ->
fp16 embeddings -> (n, sum, outer-product) statistics -> [all-reduce] -> Frechet distance against
fixed baseline statistics.  `value` is audio-seconds embedded per second over all ranks with the
PCM already resident in HBM; `e2e` is the same step fed from pinned HOST memory (H2D inside the
timed region, FAD scalar read back).  `--impl reference` times the reference's CPU implementation
of the path (torch-CPU fp32 restatement of the third-party model + reference-pinned numpy
statistics/Frechet, oracle/) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

import os as _os
_os.environ.setdefault("FADTK_SYNTHETIC", "1")      # benchmarks run the real architectures on seeded random weights (no checkpoints offline)
ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CLIP_SECONDS = 10.0
ROWS_PER_CLIP = 10
# tensor-core layers of VGGish (conv1's 7.1 MFLOP run on the CUDA cores): 2*M*N*K per example
UMMA_LAYER_FLOP = {
    "conv2": 2 * 48 * 32 * 128 * 576, "conv3_1": 2 * 24 * 16 * 256 * 1152, "conv3_2": 2 * 24 * 16 * 256 * 2304,
    "conv4_1": 2 * 12 * 8 * 512 * 2304, "conv4_2": 2 * 12 * 8 * 512 * 4608,
    "fc1": 2 * 12288 * 4096, "fc2": 2 * 4096 * 4096, "fc3": 2 * 4096 * 128,
}
# HTSAT-tiny GEMMs per 10-s window: 24 T C^2 per Swin block + 3 patch-merging reductions
def _htsat_gemm_flop(embed, depths):
    dims = [(d, 4096 >> (2 * i), embed << i) for i, d in enumerate(depths)]
    return sum(d * 24 * t * c * c for d, t, c in dims) + sum(2 * (t // 4) * 4 * c * 2 * c for _, t, c in dims[:3])


def _whisper_gemm_flop(d, n_enc, n_dec):
    """tensor-core GEMM FLOPs per clip (30-s padded input): conv stem (as 3-tap GEMMs), encoder layers,
    cross-attention K/V projections; the 2-token decoder GEMMs are negligible."""
    f = 4 * d
    stem = 2 * 3000 * d * 240 + 2 * 1500 * d * 3 * d
    enc = n_enc * 2 * 1500 * (4 * d * d + 2 * d * f)
    ckv = n_dec * 2 * 1500 * 2 * d * d
    return stem + enc + ckv


CLAP_GEMM_FLOP = _htsat_gemm_flop(96, (2, 2, 6, 2))
WHISPER_SMALL_GEMM_FLOP = _whisper_gemm_flop(768, 12, 12)


def _encodec_gemm_flop(T=240000):
    """algorithmic conv + LSTM FLOPs of the 24 kHz SEANet encoder per clip of T samples (2*M*N*K, no padding)"""
    fl, ch, t = 2 * T * 32 * 7, 32, T
    for r in (2, 4, 5, 8):
        fl += 2 * t * (ch // 2) * 3 * ch + 2 * t * ch * (ch // 2) + 2 * t * ch * ch     # conv3, conv1, shortcut
        t = -(-t // r)
        fl += 2 * t * 2 * ch * 2 * r * ch                                               # down conv
        ch *= 2
    fl += 2 * (2 * t * 2048 * 512 * 2)                                                 # LSTM: input + recurrent, 2 layers
    fl += 2 * t * 128 * 7 * 512
    return fl


ENCODEC_GEMM_FLOP = _encodec_gemm_flop()


def _w2v_gemm_flop(L=160000, d=768, layers=12, ffn=3072):
    """conv feature encoder + projection + positional conv + transformer GEMMs per clip of L samples"""
    t, fl, cin = L, 0, 1
    for k, s_ in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        t = (t - k) // s_ + 1
        fl += 2 * t * 512 * k * cin
        cin = 512
    fl += 2 * t * d * 512 + 2 * t * d * (d // 16) * 128
    return fl + layers * 2 * t * (4 * d * d + 2 * d * ffn)


W2V2_BASE_GEMM_FLOP = _w2v_gemm_flop()
CLAP_MUSIC_GEMM_FLOP = _htsat_gemm_flop(128, (2, 2, 12, 2))

MODELS = {
    "vggish": dict(sr=16000, clips=10000, baseline_clips=1000, chunk_clips=1000, d=128,
                   workload="VGGish FAD, {clips} x 10 s synthetic 16 kHz clips per GPU vs {base}-clip baseline (BASELINE.json configs[1])",
                   rows_flop=sum(UMMA_LAYER_FLOP.values())),
    "clap-laion-audio": dict(sr=48000, clips=500, baseline_clips=100, chunk_clips=50, d=512,
                             workload="clap-laion-audio (HTSAT-tiny) FAD, {clips} x 10 s synthetic 48 kHz clips per GPU vs {base}-clip "
                                      "baseline (BASELINE.json configs[2] at single-GPU size)",
                             rows_flop=CLAP_GEMM_FLOP),
    "clap-laion-music": dict(sr=48000, clips=250, baseline_clips=50, chunk_clips=25, d=512,
                             workload="clap-laion-music (HTSAT-base) FAD, {clips} x 10 s synthetic 48 kHz clips per GPU vs {base}-clip baseline",
                             rows_flop=CLAP_MUSIC_GEMM_FLOP),
    "encodec-emb": dict(sr=24000, clips=512, baseline_clips=64, chunk_clips=512, d=128,
                        workload="encodec-emb (24 kHz SEANet encoder) FAD, {clips} x 10 s synthetic 24 kHz clips per GPU (750 rows per clip) "
                                 "vs {base}-clip baseline (BASELINE.json configs[3] embedding stage)",
                        rows_flop=ENCODEC_GEMM_FLOP),
    "w2v2-base": dict(sr=16000, clips=256, baseline_clips=32, chunk_clips=32, d=768,
                      workload="w2v2-base (hidden_states[12]) FAD, {clips} x 10 s synthetic 16 kHz clips per GPU (499 rows per clip) vs {base}-clip baseline",
                      rows_flop=W2V2_BASE_GEMM_FLOP),
    "whisper-small": dict(sr=16000, clips=256, baseline_clips=64, chunk_clips=64, d=768,
                          workload="whisper-small FAD, {clips} x 10 s synthetic 16 kHz clips per GPU (each padded to 30 s, 2 rows per clip) "
                                   "vs {base}-clip baseline (BASELINE.json configs[4] embedding stage)",
                          rows_flop=WHISPER_SMALL_GEMM_FLOP),
}


def cpu_scoring_indiv(mu_b, cov_b, songs):
    """cpu_baseline leg of benchmarks/scoring.py --mode indiv: the reference arithmetic (oracle) per song.
    -> (scores, seconds)"""
    from oracle import fad_oracle as fo
    t0 = time.perf_counter()
    want = [fo.frechet_distance(mu_b, cov_b, *fo.embd_statistics(s)) for s in songs]
    return want, time.perf_counter() - t0


def cpu_scoring_inf(mu_b, cov_b, rows, steps, k):
    """cpu_baseline leg of benchmarks/scoring.py --mode inf: the first k bootstrap sizes on the oracle, consuming the
    global numpy RNG exactly like the reference (seeded 0 here).  -> (scores, seconds in gather+cov, seconds in Frechet, sizes)"""
    from oracle import fad_oracle as fo
    sizes = fo.inf_sample_sizes(len(rows), steps, 500)
    np.random.seed(0)
    pts, t_stats, t_fr = [], 0.0, 0.0
    for n in sizes[:k]:
        t0 = time.perf_counter()
        pick = np.random.choice(rows.shape[0], size=n, replace=True)
        st = fo.embd_statistics(rows[pick])
        t1 = time.perf_counter()
        pts.append(fo.frechet_distance(mu_b, cov_b, *st))
        t_stats += t1 - t0
        t_fr += time.perf_counter() - t1
    return pts, t_stats, t_fr, sizes


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1443.3), d.get("hbm_gbs", 6567.7), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.tmp.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit()]
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(rows[0][1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any("Active" in r[3 + i] and "Not" not in r[3 + i] for r in rows):
                out["reasons"].append(n)
        out["samples"] = len(rows)
        return out


def oracle_embed_fn(model: str, state):
    if model == "vggish":
        from oracle import vggish_oracle as vo
        return lambda pcm: vo.embed(vo.load_wav_semantics(pcm), state)
    if model == "encodec-emb":
        from oracle import encodec_oracle as eo
        return lambda pcm: eo.embed(pcm / 32768.0, state)
    if model == "w2v2-base":
        from oracle import w2v_oracle as wv
        hf, fe = wv.build(state, "w2v2")
        return lambda pcm: wv.embed(pcm / 32768.0, hf, fe, 12)
    if model.startswith("whisper-"):
        from fadtk_b200 import weights_whisper
        from oracle import whisper_oracle as wo
        hf, fe = wo.build(state, weights_whisper.SYNTH_START)
        return lambda pcm: wo.embed(pcm / 32768.0, hf, fe, weights_whisper.SYNTH_START)
    from oracle import clap_oracle as co
    return lambda pcm: co.embed(pcm / 32768.0, state)


def cpu_reference_leg(model, pcm_clips: np.ndarray, base_stats, state, budget_s: float = 15.0):
    """Reference CPU path on a bounded sample: per-clip loop (fad_batch.py semantics), fp32 torch
    restatement of the model, fp16 cache rounding, per-file statistics + Chan merge (utils.py:13-46),
    eig-route Frechet.  -> dict"""
    from oracle import fad_oracle as fo
    embed = oracle_embed_fn(model, state)
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    embs = []
    used = 0
    for i in range(pcm_clips.shape[0]):
        embs.append(embed(pcm_clips[i]))
        used += 1
        if time.perf_counter() - t0 > budget_s and used >= 4:
            break
    t_embed = time.perf_counter() - t0
    t1 = time.perf_counter()
    mu, cov = fo.online_statistics(embs)          # one clip = one file: utils.py:19-46 semantics
    t_stats = time.perf_counter() - t1
    t2 = time.perf_counter()
    fad = fo.frechet_distance(base_stats[0], base_stats[1], mu, cov)
    t_fr = time.perf_counter() - t2
    total = t_embed + t_stats + t_fr
    return {"value": used * CLIP_SECONDS / total, "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{used} of the eval clips ({used * CLIP_SECONDS:.0f} audio-s): embed {t_embed:.2f}s, "
                      f"stats {t_stats:.3f}s, frechet {t_fr:.3f}s; "
                      + ("transformers WhisperFeatureExtractor + WhisperModel on CPU (the reference's own dependency, "
                         "driven as model_loader.py:663-669) + numpy/scipy" if model.startswith("whisper-") else
                         f"oracle/ torch-CPU fp32 {model} + numpy/scipy (reference third-party model is not installable offline)"),
            "fad": float(fad), "clips": used, "seconds": total}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vggish", choices=list(MODELS))
    ap.add_argument("--clips", type=int, default=0, help="eval clips per GPU (0 = the model's default)")
    ap.add_argument("--baseline-clips", type=int, default=0)
    ap.add_argument("--chunk-clips", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    spec = MODELS[args.model]
    args.clips = args.clips or spec["clips"]
    args.baseline_clips = args.baseline_clips or spec["baseline_clips"]
    args.chunk_clips = args.chunk_clips or spec["chunk_clips"]
    sr = spec["sr"]
    clip_samples = int(sr * CLIP_SECONDS)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pcm_gb = args.clips * clip_samples * 2 / 1e9
    config = {"workload": spec["workload"].format(clips=args.clips, base=args.baseline_clips),
              "model": f"{args.model} (seeded synthetic weights, real architecture)", "clips_per_gpu": args.clips,
              "clip_seconds": CLIP_SECONDS, "chunk_clips": args.chunk_clips,
              "l2": f"inputs ({pcm_gb:.1f} GB PCM per GPU) exceed L2; no explicit flush", "parallelism": f"dp{world}"}

    from fadtk_b200 import synth, weights, weights_clap
    if args.model == "vggish":
        state = weights.synthetic_vggish_state(0)
    elif args.model == "encodec-emb":
        from fadtk_b200 import weights_encodec
        state = weights_encodec.synthetic_encodec_state(0)
    elif args.model == "w2v2-base":
        from fadtk_b200 import weights_w2v
        state = weights_w2v.synthetic_w2v_state(0)
    elif args.model.startswith("whisper-"):
        from fadtk_b200 import weights_whisper
        state = weights_whisper.synthetic_whisper_state(0, args.model.split("-", 1)[1])
    else:
        state = weights_clap.synthetic_clap_state(0, "base" if args.model == "clap-laion-music" else "tiny")

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        embed = oracle_embed_fn(args.model, state)
        n_base, n_eval = (16, 64) if args.model == "vggish" else (4, 16)
        base = np.concatenate([embed(synth.musiclike_clip(i, CLIP_SECONDS, sr, True)) for i in range(n_base)])
        base_stats = (base.astype(np.float64).mean(0), np.cov(base.astype(np.float64), rowvar=False))
        sample = np.stack([synth.musiclike_clip(i, CLIP_SECONDS, sr) for i in range(n_eval)])
        per_step = max(4.0, 40.0 / max(1, args.steps + args.warmup))
        for _ in range(args.warmup):
            cpu_reference_leg(args.model, sample, base_stats, state, budget_s=per_step)
        legs = [cpu_reference_leg(args.model, sample, base_stats, state, budget_s=per_step) for _ in range(args.steps)]
        secs = sum(l["seconds"] for l in legs)
        clips = sum(l["clips"] for l in legs)
        val = clips * CLIP_SECONDS / secs
        cb = dict(legs[-1]); cb["value"] = val
        cb.pop("fad"); cb.pop("clips"); cb.pop("seconds")
        print(json.dumps({"impl": "reference", "metric": "audio_seconds_embedded_per_second", "value": val,
                          "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------ our arm
    from fadtk_b200 import _native, dist
    from fadtk_b200.pipeline import EvalSetFAD
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    eng = _native.Engine(local_rank, max_examples=args.chunk_clips * ROWS_PER_CLIP)
    if args.model == "vggish":
        eng.vggish_load(weights.pack_vggish(state))
    elif args.model == "w2v2-base":
        eng.w2v_load(weights_w2v.config_of(state), weights_w2v.pack_w2v(state), args.chunk_clips, max_len=int(CLIP_SECONDS * sr))
    elif args.model == "encodec-emb":
        eng.encodec_load(weights_encodec.pack_encodec(state), max_chunk_samples=16 * int(CLIP_SECONDS * sr))
    elif args.model.startswith("whisper-"):
        eng.whisper_load(weights_whisper.config_of(state), weights_whisper.pack_whisper(state, weights_whisper.SYNTH_START),
                         max_clips=args.chunk_clips)
    else:
        eng.clap_load(weights_clap.pack_clap(state), max_chunks=args.chunk_clips * ROWS_PER_CLIP)

    # baseline statistics (identical on every rank), outside the timed region
    d = spec["d"]
    zero_mu, eye = torch.zeros(d, dtype=torch.float64), torch.eye(d, dtype=torch.float64)
    helper = EvalSetFAD(eng, zero_mu, eye, clip_samples, clips_per_chunk=args.chunk_clips, model=args.model)
    base_pcm = synth.musiclike_device(args.baseline_clips, CLIP_SECONDS, sr, seed=30_000, device=dev,
                                      fmax=1500.0, noise=0.08)
    base_emb = torch.cat([helper.embed(base_pcm[s:s + args.chunk_clips]) for s in range(0, args.baseline_clips, args.chunk_clips)])
    shift = base_emb[:4096].float().mean(0).to(torch.float16)
    acc = eng.stats_accumulate(base_emb, shift, eng.stats_new(d))
    mu_b, cov_b = eng.stats_finalize(acc, shift, d)
    del base_pcm

    pcm = synth.musiclike_device(args.clips, CLIP_SECONDS, sr, seed=20_000 + rank, device=dev)
    job = EvalSetFAD(eng, mu_b, cov_b, clip_samples, clips_per_chunk=args.chunk_clips, model=args.model)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        res = job.run_device(pcm)
    sync_all()
    launches0 = eng.launches
    eng.profile(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = job.run_device(pcm)
    e1.record()
    sync_all()
    ms = dist.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    prof = eng.profile_collect()
    eng.profile(False)
    launches = eng.launches - launches0
    fad_value = float(res[0].item())
    audio_s = world * args.clips * CLIP_SECONDS * args.steps
    value = audio_s / (ms / 1000.0)

    # ---- roofline of the dominant kernel: the tcgen05 conv/FC (GEMM) kernel
    peak_tf, peak_hbm, peak_src = measured_peaks()
    rows = args.clips * (1 if args.model.startswith("whisper-") or args.model in ("encodec-emb", "w2v2-base") else ROWS_PER_CLIP) * args.steps   # examples (VGGish) / 10-s windows (CLAP) / clips (Whisper)
    gemm_keys = list(UMMA_LAYER_FLOP) if args.model == "vggish" else ["clap_gemm"]
    umma_ms = sum(prof[k][0] for k in gemm_keys if k in prof)
    umma_launch = sum(prof[k][1] for k in gemm_keys if k in prof)
    umma_flop = spec["rows_flop"] * rows
    achieved = umma_flop / (umma_ms / 1000.0) / 1e12 if umma_ms > 0 else 0.0
    per_layer = {k: {"ms_per_launch": prof[k][0] / prof[k][1], "tflops": UMMA_LAYER_FLOP[k] * rows / (prof[k][0] / 1000.0) / 1e12}
                 for k in UMMA_LAYER_FLOP if k in prof and prof[k][0] > 0} if args.model == "vggish" else None
    other = {k: {"ms_total": v[0], "launches": v[1]} for k, v in prof.items() if k not in gemm_keys}
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/), scaled
    # to this run's rows per launch; None when no capture exists for the model
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists() and umma_launch:
        t = json.loads(tf.read_text()).get(args.model)
        if t:
            traffic = t["dram_gb_per_row"] * rows / umma_launch
            traffic_src = t["source"]
    roofline = {"kernel": "fad::conv_gemm_kernel<128,4,SPLIT_W> (tcgen05 kind::f16, hi/lo split fp16 weights: 2 MMAs per K step)",
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "peak_source": peak_src, "traffic": traffic, "traffic_unit": "GB per launch",
                "traffic_source": traffic_src,
                "issued_tflops": 2.0 * achieved, "issued_frac": 2.0 * achieved / peak_tf,
                "note": "achieved counts ALGORITHMIC FLOPs (2*M*N*K once); the kernel issues twice that (W = Wh + Wl) to keep FAD within 1e-4",
                "launches": umma_launch, "avg_launch_ms": umma_ms / max(1, umma_launch),
                "algorithmic_gflop_per_row": spec["rows_flop"] / 1e9,
                "share_of_step": umma_ms / ms if ms > 0 else None,
                "per_layer": per_layer, "other_kernels": other}

    # ---- end to end from pinned host memory
    e2e = None
    if not args.no_e2e:
        host = torch.empty((args.clips, clip_samples), dtype=torch.int16, pin_memory=True)
        host.copy_(pcm)
        torch.cuda.synchronize()
        for _ in range(2):
            fad_h = job.run_host(host)
        sync_all()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            fad_h = job.run_host(host)
        g1.record()
        sync_all()
        ms_e = dist.max_over_ranks(max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1000.0))
        e2e = {"value": audio_s / (ms_e / 1000.0), "unit": "audio-s/s", "ms_per_step": ms_e / args.steps,
               "h2d_bytes_per_step": int(args.clips * clip_samples * 2), "d2h_bytes_per_step": 8, "fad": fad_h,
               "api": "fadtk_b200.pipeline.EvalSetFAD.run_host (pinned int16 PCM in, FAD float out)"}

    if rank != 0:
        dist.shutdown()
        return

    # ---- CPU baseline + parity sample (rank 0, N = 1 only)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_sample = 64 if args.model == "vggish" else 24
        sample = pcm[:n_sample].cpu().numpy()
        base_stats = (mu_b.cpu().numpy(), cov_b.cpu().numpy())
        cpu = cpu_reference_leg(args.model, sample, base_stats, state, budget_s=15.0)
        n = cpu["clips"]
        # same clips through the GPU path -> FAD vs the CPU oracle's FAD on identical audio
        sub = EvalSetFAD(eng, mu_b, cov_b, clip_samples, clips_per_chunk=args.chunk_clips, model=args.model)
        sub.shift = job.shift
        fad_gpu_sample = float(sub.run_device(pcm[:n].contiguous())[0].item())
        parity = {"clips": n, "fad_gpu": fad_gpu_sample, "fad_cpu_oracle": cpu["fad"],
                  "rel_err": abs(fad_gpu_sample - cpu["fad"]) / abs(cpu["fad"])}
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": "audio_seconds_embedded_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": config, "fad": fad_value, "fad_wallclock_s": ms / args.steps / 1000.0,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "parity_sample": parity}
    print(json.dumps(line))
    dist.shutdown()


if __name__ == "__main__":
    main()
