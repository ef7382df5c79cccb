"""Scoring-stage benchmarks for SURVEY.md §8 rows a11 (FAD-inf) and a12 (per-song FAD).

Not the headline metric (that is bench.py): these time the statistics + Frechet stages on cached
embeddings at the sizes of BASELINE.json configs 4 and 5, through the public API the reference
exposes (FrechetAudioDistance.score_individual / score_inf on .npy caches), next to the CPU oracle
(the reference's own arithmetic) on a bounded sample.  One JSON line per mode.

    python benchmarks/scoring.py --mode indiv [--songs 5000 --rows 750 --dim 128]
    python benchmarks/scoring.py --mode inf   [--n 50000 --dim 768 --steps 25]
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

import os as _os
_os.environ.setdefault("FADTK_SYNTHETIC", "1")      # benchmarks run the real architectures on seeded random weights (no checkpoints offline)
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import fadtk_b200 as fk                      # noqa: E402
from fadtk_b200 import _native               # noqa: E402
from bench import cpu_scoring_indiv, cpu_scoring_inf   # noqa: E402  (the CPU-oracle baseline legs live in bench.py)


class CachedLoader(fk.ModelLoader):
    """Plugin without a forward pass: scoring of cached embeddings only."""

    def __init__(self, d):
        super().__init__("cached", d, 16000)

    def load_model(self):
        pass

    def _get_embedding(self, audio):
        raise NotImplementedError


def synth_rows(rng, n, d, mix, gain=1.0, shift=0.0):
    return ((rng.standard_normal((n, d), dtype=np.float32) @ mix) * gain + shift).astype(np.float16)


def baseline_statistics(rows):
    """fp64 mean / covariance of the synthetic baseline (what load_stats hands out, fad.py:286-288)."""
    x = rows.astype(np.float64)
    return x.mean(0), np.cov(x, rowvar=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["indiv", "inf"], required=True)
    ap.add_argument("--songs", type=int, default=5000)
    ap.add_argument("--rows", type=int, default=750)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--cpu-sample", type=int, default=0, help="songs / bootstrap steps timed on the CPU oracle")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    eng = _native.engine(0)
    dev = eng.torch_device
    rng = np.random.default_rng(0)

    if args.mode == "indiv":
        d = args.dim or 128
        mix = (rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
        base_rows = synth_rows(rng, 20000, d, mix)
        mu_b, cov_b = baseline_statistics(base_rows)
        songs = [synth_rows(rng, args.rows, d, mix, 0.7 + 0.6 * rng.random(), 0.2 * rng.random()) for _ in range(args.songs)]
        # ---- device-resident: one ragged batch
        base = _native.Baseline(eng, mu_b, cov_b)
        offs = torch.from_numpy(np.arange(args.songs + 1, dtype=np.int64) * args.rows).to(dev)
        flat = torch.from_numpy(np.concatenate(songs)).to(dev)
        base.frechet_batched(flat, offs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = base.frechet_batched(flat, offs)
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
        # ---- public API on .npy caches (file reads + H2D + D2H + sort + csv inside the timed region)
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            np.savez(td / "base.npz", **{"cached.mu": mu_b, "cached.cov": cov_b})
            (td / "ev" / "embeddings" / "cached").mkdir(parents=True)
            for i, s in enumerate(songs):
                (td / "ev" / f"s{i:05d}.wav").write_bytes(b"")
                np.save(td / "ev" / "embeddings" / "cached" / f"s{i:05d}.npy", s)
            fad = fk.FrechetAudioDistance(CachedLoader(d), audio_load_worker=1, load_model=False)
            t0 = time.perf_counter()
            csv = fad.score_individual(td / "base.npz", td / "ev", td / "out.csv")
            api_s = time.perf_counter() - t0
            rows = csv.read_text().splitlines()
        # ---- CPU oracle (reference arithmetic) on a bounded sample
        k = args.cpu_sample or min(args.songs, 24)
        want, cpu_s = cpu_scoring_indiv(mu_b, cov_b, songs[:k])
        got = out[:k, 0].cpu().numpy()
        rel = float(np.max(np.abs(got - np.array(want)) / np.abs(want)))
        print(json.dumps({
            "mode": "indiv", "workload": f"{args.songs} songs x [{args.rows}, {d}] fp16 vs one baseline (BASELINE.json configs[3] scoring stage)",
            "device_ms": dev_ms, "songs_per_s_device": args.songs / (dev_ms / 1e3),
            "api_s": api_s, "songs_per_s_api": args.songs / api_s, "csv_rows": len(rows),
            "api": "FrechetAudioDistance.score_individual on .npy caches -> csv",
            "cpu_baseline": {"songs_per_s": k / cpu_s, "sample": f"{k} songs, oracle (numpy/scipy reference arithmetic), {torch.get_num_threads()} threads", "kind": "port"},
            "parity_max_rel_err": rel, "gpu_launches": int(eng.launches)}))
    else:
        d = args.dim or 768
        mix = (rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
        base_rows = synth_rows(rng, args.n, d, mix)
        mu_b, cov_b = baseline_statistics(base_rows)
        rows = synth_rows(rng, args.n, d, mix, 1.1, 0.05)
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            np.savez(td / "base.npz", **{"cached.mu": mu_b, "cached.cov": cov_b})
            files = []
            for i, chunk in enumerate(np.array_split(rows, 250)):
                np.save(td / f"e{i:04d}.npy", chunk)
                files.append(td / f"e{i:04d}.npy")
            fad = fk.FrechetAudioDistance(CachedLoader(d), audio_load_worker=1, load_model=False)
            np.random.seed(0)
            fad.score_inf(td / "base.npz", files, steps=3)              # warm-up (allocations, baseline root)
            np.random.seed(0)
            t0 = time.perf_counter()
            res = fad.score_inf(td / "base.npz", files, steps=args.steps)
            api_s = time.perf_counter() - t0
        k = args.cpu_sample or 2
        cpu_pts, t_stats, t_fr, sizes = cpu_scoring_inf(mu_b, cov_b, rows, args.steps, k)
        cpu_s = t_stats + t_fr
        gpu_pts = [p[1] for p in res.points[:k]]
        rel = float(np.max(np.abs(np.array(gpu_pts) - np.array(cpu_pts)) / np.abs(cpu_pts)))
        # per step the gather + np.cov cost grows linearly in n, the eig/sqrtm cost is constant
        cpu_est = t_stats * float(sum(sizes)) / float(sum(sizes[:k])) + t_fr * args.steps / k
        print(json.dumps({
            "mode": "inf", "workload": f"FAD-inf, N = {args.n} x {d} fp16, {args.steps} bootstrap sizes (BASELINE.json configs[4] scoring stage)",
            "api_s": api_s, "steps_per_s": args.steps / api_s, "fad_inf": res.score, "r2": res.r2,
            "api": "FrechetAudioDistance.score_inf on .npy caches",
            "cpu_baseline": {"seconds_measured": cpu_s, "steps_measured": k, "seconds_extrapolated_all_steps": cpu_est,
                             "sample": f"first {k} of {args.steps} bootstrap sizes, oracle (numpy/scipy reference arithmetic), same RNG stream", "kind": "port"},
            "parity_max_rel_err_first_steps": rel, "gpu_launches": int(eng.launches)}))


if __name__ == "__main__":
    main()
