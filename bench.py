#!/usr/bin/env python
"""Benchmark of the hot path: VGGish FAD on 10 000 x 10 s synthetic 16 kHz clips per GPU
(BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over the eval set: PCM16 -> log-mel -> VGGish ->
fp16 embeddings -> (n, sum, outer-product) statistics -> [all-reduce] -> Frechet distance against
fixed baseline statistics.  `value` is audio-seconds embedded per second over all ranks with the
PCM already resident in HBM; `e2e` is the same step fed from pinned HOST memory (H2D inside the
timed region, FAD scalar read back).  `--impl reference` times the reference's CPU implementation
of the path (torch-CPU fp32 VGGish restatement + reference-pinned numpy statistics/Frechet,
oracle/) on a bounded sample of the same workload.
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

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR = 16000
CLIP_SECONDS = 10.0
CLIP_SAMPLES = int(SR * CLIP_SECONDS)
ROWS_PER_CLIP = 10
GFLOP_PER_EXAMPLE = 1.727791104            # BASELINE.md section 5 (conv 1.5925 + FC 0.1353)
# tensor-core layers only (conv1's 7.1 MFLOP run on the CUDA cores): 2*M*N*K per example
UMMA_LAYER_FLOP = {
    "conv2": 2 * 48 * 32 * 128 * 576, "conv3_1": 2 * 24 * 16 * 256 * 1152, "conv3_2": 2 * 24 * 16 * 256 * 2304,
    "conv4_1": 2 * 12 * 8 * 512 * 2304, "conv4_2": 2 * 12 * 8 * 512 * 4608,
    "fc1": 2 * 12288 * 4096, "fc2": 2 * 4096 * 4096, "fc3": 2 * 4096 * 128,
}


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


def cpu_reference_leg(pcm_clips: np.ndarray, base_stats, state, budget_s: float = 15.0):
    """Reference CPU path on a bounded sample: per-clip loop (fad_batch.py semantics), fp32 torch
    VGGish restatement, fp16 cache rounding, per-file statistics + Chan merge (utils.py:13-46),
    eig-route Frechet.  -> dict"""
    from oracle import fad_oracle as fo, vggish_oracle as vo
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    embs = []
    used = 0
    for i in range(pcm_clips.shape[0]):
        embs.append(vo.embed(vo.load_wav_semantics(pcm_clips[i]), state))
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
                      f"stats {t_stats:.3f}s, frechet {t_fr:.3f}s; oracle/ torch-CPU fp32 VGGish + numpy/scipy "
                      f"(reference third-party model is not installable offline)",
            "fad": float(fad), "clips": used, "seconds": total}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=10000, help="eval clips per GPU")
    ap.add_argument("--baseline-clips", type=int, default=1000)
    ap.add_argument("--chunk-clips", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"VGGish FAD, {args.clips} x 10 s synthetic 16 kHz clips per GPU vs {args.baseline_clips}-clip baseline "
                          f"(BASELINE.json configs[1])",
              "model": "vggish (seeded synthetic weights, real architecture)", "clips_per_gpu": args.clips,
              "clip_seconds": CLIP_SECONDS, "chunk_clips": args.chunk_clips,
              "l2": "inputs (3.2 GB PCM per GPU) exceed L2; no explicit flush", "parallelism": f"dp{world}"}

    from fadtk_b200 import synth, weights
    state = weights.synthetic_vggish_state(0)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import fad_oracle as fo, vggish_oracle as vo
        base = np.concatenate([vo.embed(vo.load_wav_semantics(synth.musiclike_clip(i, CLIP_SECONDS, SR, True)), state)
                               for i in range(16)])
        base_stats = (base.astype(np.float64).mean(0), np.cov(base.astype(np.float64), rowvar=False))
        sample = np.stack([synth.musiclike_clip(i, CLIP_SECONDS, SR) for i in range(64)])
        per_step = max(4.0, 40.0 / max(1, args.steps + args.warmup))
        for _ in range(args.warmup):
            cpu_reference_leg(sample, base_stats, state, budget_s=per_step)
        legs = [cpu_reference_leg(sample, base_stats, state, budget_s=per_step) for _ in range(args.steps)]
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
    eng.vggish_load(weights.pack_vggish(state))

    # baseline statistics (identical on every rank), outside the timed region
    base_pcm = synth.musiclike_device(args.baseline_clips, CLIP_SECONDS, SR, seed=30_000, device=dev)
    off = np.arange(args.baseline_clips + 1, dtype=np.int64) * CLIP_SAMPLES
    ex, _ = eng.vggish_plan(off)
    base_emb = eng.vggish_forward(base_pcm.reshape(-1), torch.from_numpy(ex).to(dev))
    shift = base_emb[:4096].float().mean(0).to(torch.float16)
    acc = eng.stats_accumulate(base_emb, shift, eng.stats_new(128))
    mu_b, cov_b = eng.stats_finalize(acc, shift, 128)
    del base_pcm

    pcm = synth.musiclike_device(args.clips, CLIP_SECONDS, SR, seed=20_000 + rank, device=dev)
    job = EvalSetFAD(eng, mu_b, cov_b, CLIP_SAMPLES, clips_per_chunk=args.chunk_clips)

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

    # ---- roofline of the dominant kernel (the tcgen05 conv/FC kernel, all 8 layers)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    umma_ms = sum(prof[k][0] for k in UMMA_LAYER_FLOP if k in prof)
    umma_launch = sum(prof[k][1] for k in UMMA_LAYER_FLOP if k in prof)
    examples = args.clips * ROWS_PER_CLIP * args.steps
    umma_flop = sum(UMMA_LAYER_FLOP.values()) * examples
    achieved = umma_flop / (umma_ms / 1000.0) / 1e12 if umma_ms > 0 else 0.0
    per_layer = {k: {"ms_per_launch": prof[k][0] / prof[k][1], "tflops": UMMA_LAYER_FLOP[k] * examples / (prof[k][0] / 1000.0) / 1e12}
                 for k in UMMA_LAYER_FLOP if k in prof and prof[k][0] > 0}
    other = {k: {"ms_total": v[0], "launches": v[1]} for k, v in prof.items() if k not in UMMA_LAYER_FLOP}
    roofline = {"kernel": "fad::conv_gemm_kernel (tcgen05 kind::f16 implicit-GEMM conv3x3 / FC, 8 layer launches per chunk)",
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "peak_source": peak_src, "traffic": None,
                "launches": umma_launch, "avg_launch_ms": umma_ms / max(1, umma_launch),
                "algorithmic_gflop_per_example": sum(UMMA_LAYER_FLOP.values()) / 1e9,
                "share_of_step": umma_ms / ms if ms > 0 else None,
                "per_layer": per_layer, "other_kernels": other}

    # ---- end to end from pinned host memory
    e2e = None
    if not args.no_e2e:
        host = torch.empty((args.clips, CLIP_SAMPLES), dtype=torch.int16, pin_memory=True)
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
               "h2d_bytes_per_step": int(args.clips * CLIP_SAMPLES * 2 + args.chunk_clips * ROWS_PER_CLIP * 8),
               "d2h_bytes_per_step": 8, "fad": fad_h,
               "api": "fadtk_b200.pipeline.EvalSetFAD.run_host (pinned int16 PCM in, FAD float out)"}

    if rank != 0:
        return

    # ---- CPU baseline + parity sample (rank 0, N = 1 only)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import fad_oracle as fo, vggish_oracle as vo
        sample = pcm[:64].cpu().numpy()
        base_stats = (mu_b.cpu().numpy(), cov_b.cpu().numpy())
        cpu = cpu_reference_leg(sample, base_stats, state, budget_s=15.0)
        n = cpu["clips"]
        # same clips through the GPU path -> FAD vs the CPU oracle's FAD on identical audio
        sub = EvalSetFAD(eng, mu_b, cov_b, CLIP_SAMPLES, clips_per_chunk=args.chunk_clips)
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


if __name__ == "__main__":
    main()
