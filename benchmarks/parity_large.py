"""FAD parity at a larger scale than the unit tests ("nothing checks parity at configs[1] scale or for CLAP
beyond 21 clips"): the same synthetic clips through the GPU path (plugin forward -> fp16 embeddings -> exact statistics ->
Newton-Schulz Frechet) and through the CPU oracle (fp32 torch restatement of the model, fp16 cache rounding, numpy
statistics, eig-route Frechet), relative error of the two FAD values.  One JSON line.

    python benchmarks/parity_large.py --model vggish --clips 1000          (~2-3 min of CPU oracle on 64 cores)
    python benchmarks/parity_large.py --model clap-laion-audio --clips 200
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np      # noqa: E402
import torch            # noqa: E402

import bench            # noqa: E402  (model table, oracle embedders, host core count)
import fadtk_b200 as fk  # noqa: E402
from fadtk_b200 import synth  # noqa: E402
from oracle import fad_oracle as fo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vggish", choices=["vggish", "clap-laion-audio", "encodec-emb", "whisper-small"])
    ap.add_argument("--clips", type=int, default=1000)
    ap.add_argument("--seconds", type=float, default=10.0)
    args = ap.parse_args()
    spec = bench.MODELS[args.model]
    sr = spec["sr"]
    torch.set_num_threads(bench.HOST_CORES)
    state = bench.synthetic_state(args.model)
    embed_cpu = bench.oracle_embed_fn(args.model, state)
    ml = bench.make_loader(args.model, min(args.clips, spec["chunk_clips"]))
    ml.load_model()
    # two clearly different populations (as bench.py builds them): a baseline with low-passed partials and more noise, so
    # the FAD is not a difference of nearly equal numbers (two draws of the SAME population give FAD ~ 1e-3 with CLAP's
    # L2-normalised embeddings: 1e-4 relative of that is below the eig-route noise of the reference itself)
    dev = torch.device("cuda", 0)
    base = synth.musiclike_device(args.clips, args.seconds, sr, seed=30_000, device=dev, fmax=1500.0, noise=0.08).cpu().numpy()
    evl = synth.musiclike_device(args.clips, args.seconds, sr, seed=20_000, device=dev).cpu().numpy()
    sets = {"base": [base[i] for i in range(args.clips)], "eval": [evl[i] for i in range(args.clips)]}
    t0 = time.perf_counter()
    step = spec["chunk_clips"]
    gpu = {k: np.concatenate([e for s in range(0, args.clips, step) for e in ml.embed_pcm_batch(v[s:s + step])]) for k, v in sets.items()}
    t_gpu = time.perf_counter() - t0
    fad_gpu = float(fk.calc_frechet_distance(*fk.calc_embd_statistics(gpu["base"]), *fk.calc_embd_statistics(gpu["eval"])))
    # the reference arithmetic on the GPU embeddings isolates the statistics + Frechet stages
    fad_gpu_emb_cpu_stats = float(fo.frechet_distance(*fo.embd_statistics(gpu["base"]), *fo.embd_statistics(gpu["eval"])))
    t1 = time.perf_counter()
    cpu = {k: np.concatenate([embed_cpu(c) for c in v]) for k, v in sets.items()}
    t_cpu = time.perf_counter() - t1
    fad_cpu = float(fo.frechet_distance(*fo.embd_statistics(cpu["base"]), *fo.embd_statistics(cpu["eval"])))
    emb_rel = float(np.sqrt(((gpu["eval"].astype(np.float64) - cpu["eval"].astype(np.float64)) ** 2).mean()
                            / (cpu["eval"].astype(np.float64) ** 2).mean()))
    print(json.dumps({"model": args.model, "clips_per_set": args.clips, "clip_seconds": args.seconds, "rows_per_set": int(gpu["eval"].shape[0]),
                      "fad_gpu": fad_gpu, "fad_cpu_oracle": fad_cpu, "rel_err": abs(fad_gpu - fad_cpu) / abs(fad_cpu),
                      "stats_frechet_rel_err_on_gpu_embeddings": abs(fad_gpu - fad_gpu_emb_cpu_stats) / abs(fad_gpu_emb_cpu_stats),
                      "embedding_rms_rel_err": emb_rel, "gpu_seconds_incl_host": t_gpu, "cpu_oracle_seconds": t_cpu,
                      "cpu_cores": bench.HOST_CORES, "tolerance": 1e-4,
                      "what": "FAD(base, eval) of identical synthetic clips: GPU path vs CPU oracle (fp32 model, fp16 cache rounding, numpy/scipy)"}))


if __name__ == "__main__":
    main()
