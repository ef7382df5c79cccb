"""Which term of the Frechet distance carries the wav2vec family's GPU-vs-reference offset?  Same audio through the
GPU forward and the CPU reference path (transformers fp32), then |dmu|^2, tr C1, tr C2, tr sqrt(C1 C2) for both, the
score with the means / covariances swapped between the two, and the regression slope of the centred GPU rows on the
centred reference rows (a uniform gain of the forward shows up there)."""
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import scipy.linalg  # noqa: E402
import fadtk_b200 as fk  # noqa: E402
from fadtk_b200 import synth, weights_w2v as ww  # noqa: E402
from oracle import w2v_oracle as wo  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sets = {"base": [synth.noise_clip(i, 4.0, 16000) for i in range(n)],
        "eval": [synth.musiclike_clip(i, 4.0, 16000) for i in range(n)]}
ml = fk.W2V2Model('base', 12, max_clips=8)
ml.load_model()
sd = ww.synthetic_w2v_state(0)
model, fe = wo.build(sd, "w2v2")
gpu = {k: np.concatenate([e for s in range(0, n, 8) for e in ml.embed_pcm_batch(v[s:s + 8])]).astype(np.float64) for k, v in sets.items()}
cpu = {k: np.concatenate([wo.embed(c / 32768.0, model, fe, 12) for c in v]).astype(np.float64) for k, v in sets.items()}


def stats(x):
    return x.mean(0), np.cov(x, rowvar=False)


def terms(m1, c1, m2, c2):
    s = scipy.linalg.sqrtm(c1 @ c2).real
    t = {"dmu2": float(((m1 - m2) ** 2).sum()), "tr1": float(np.trace(c1)), "tr2": float(np.trace(c2)), "trsqrt": float(np.trace(s))}
    t["fad"] = t["dmu2"] + t["tr1"] + t["tr2"] - 2 * t["trsqrt"]
    return t


sg = {k: stats(v) for k, v in gpu.items()}
sc = {k: stats(v) for k, v in cpu.items()}
out = {"clips": n, "gpu": terms(*sg["base"], *sg["eval"]), "cpu": terms(*sc["base"], *sc["eval"]),
       "gpu_means_cpu_covs": terms(sg["base"][0], sc["base"][1], sg["eval"][0], sc["eval"][1]),
       "cpu_means_gpu_covs": terms(sc["base"][0], sg["base"][1], sc["eval"][0], sg["eval"][1])}
out["rel"] = {k: out["gpu"][k] / out["cpu"][k] - 1 for k in out["gpu"]}
for k in sets:
    g, c = gpu[k] - gpu[k].mean(0), cpu[k] - cpu[k].mean(0)
    out[f"slope_centred_{k}"] = float((g * c).sum() / (c * c).sum() - 1)
    out[f"mean_over_rms_{k}"] = float(np.sqrt((cpu[k].mean(0) ** 2).sum() / (cpu[k] ** 2).sum() * cpu[k].shape[0]))
    e = gpu[k].mean(0) - cpu[k].mean(0)
    out[f"mean_err_along_mean_{k}"] = float((e * cpu[k].mean(0)).sum() / (cpu[k].mean(0) ** 2).sum())
    out[f"mean_err_norm_rel_{k}"] = float(np.linalg.norm(e) / np.linalg.norm(cpu[k].mean(0)))
    # per-dimension gain of the fluctuations
    gain = (g * c).sum(0) / (c * c).sum(0) - 1
    out[f"gain_dims_{k}"] = {"mean": float(gain.mean()), "std": float(gain.std())}
dm = sc["base"][0] - sc["eval"][0]
eg = (sg["base"][0] - sg["eval"][0]) - dm
out["dmu_err_along_dmu"] = float((eg * dm).sum() / (dm * dm).sum())
print(json.dumps(out))
