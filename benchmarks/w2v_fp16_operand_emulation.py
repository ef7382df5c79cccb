"""CPU experiment behind DESIGN.md section 3, finding 6: is the wav2vec family's FAD offset a property of fp16 GEMM
OPERANDS on this (seeded random, mean-dominated) model, or of the kernels?  The reference path (transformers fp32) is run
twice on the same audio: as it is, and with every Linear / Conv1d input and the q, k, v projections rounded to fp16 the way
the GPU forward stores them (weights stay fp32 - the GPU's hi/lo pair carries 22 bits; accumulation fp32).  No GPU.
Usage: python benchmarks/w2v_fp16_operand_emulation.py [clips per set, default 32]"""
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fadtk_b200 import synth, weights_w2v as ww  # noqa: E402
from oracle import fad_oracle as fo, w2v_oracle as wo  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sets = {"base": [synth.noise_clip(i, 4.0, 16000) for i in range(n)],
        "eval": [synth.musiclike_clip(i, 4.0, 16000) for i in range(n)]}
sd = ww.synthetic_w2v_state(0)
model, fe = wo.build(sd, "w2v2")


def r16(t):
    return t.to(torch.float16).to(torch.float32)


def run(round_operands: bool):
    hooks = []
    if round_operands:
        for name, m in model.named_modules():
            if isinstance(m, (torch.nn.Linear, torch.nn.Conv1d)):
                hooks.append(m.register_forward_pre_hook(lambda mod, args: (r16(args[0]),) + tuple(args[1:])))
                if name.endswith(("q_proj", "k_proj", "v_proj")):
                    hooks.append(m.register_forward_hook(lambda mod, args, out: r16(out)))
    # fp16 arrays, as the reference caches them: its np.mean then rounds the mean vector to fp16 too (fad.py:42-48)
    emb = {k: np.concatenate([wo.embed(c / 32768.0, model, fe, 12) for c in v]) for k, v in sets.items()}
    for h in hooks:
        h.remove()
    return emb


ref = run(False)
emu = run(True)
fad_ref = fo.frechet_distance(*fo.embd_statistics(ref["base"]), *fo.embd_statistics(ref["eval"]))
fad_emu = fo.frechet_distance(*fo.embd_statistics(emu["base"]), *fo.embd_statistics(emu["eval"]))
err = np.concatenate([emu[k].astype(np.float64) - ref[k].astype(np.float64) for k in ref])
rms = np.sqrt((np.concatenate([ref[k] for k in ref]).astype(np.float64) ** 2).mean())
print(json.dumps({"clips_per_set": n, "fad_reference_fp32": fad_ref, "fad_fp16_operands": fad_emu,
                  "rel": fad_emu / fad_ref - 1.0, "embedding_rms_rel_err": float(np.sqrt((err ** 2).mean()) / rms),
                  "what": "transformers fp32 on the CPU vs the same with Linear/Conv1d inputs and q/k/v rounded to fp16"}))
