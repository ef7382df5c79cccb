"""CLAP-LAION (HTSAT-tiny) parameters: seeded synthetic set, checkpoint loading, device packing.

State-dict keys follow transformers' ``ClapAudioModelWithProjection`` without the
``audio_model.audio_encoder.`` prefix (``audio_projection.*`` kept), i.e. the HF conversion of
LAION's ``630k-audioset-best.pt`` (model_loader.py:301) loads directly.  No checkpoint exists
offline, so tests and benches run on seeded synthetic parameters with the real architecture.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch

from .weights import split_hi_lo_tiles

HEADS, WINDOW, N_MEL = (4, 8, 16, 32), 8, 64
# HTSAT-tiny = clap-laion-audio, HTSAT-base = clap-laion-music (model_loader.py:385): (embed dim, depths)
VARIANTS = {"tiny": (96, (2, 2, 6, 2)), "base": (128, (2, 2, 12, 2))}
EMBED, DEPTHS = VARIANTS["tiny"]
N_TENSORS = 6 + 12 * 13 + 3 * 4 + 6


def n_tensors(variant: str = "tiny") -> int:
    return 6 + sum(VARIANTS[variant][1]) * 13 + 3 * 4 + 6


def config_of(sd: dict):
    """(embed dim, depths) read off a state dict."""
    embed = sd["patch_embed.proj.weight"].shape[0]
    depths = tuple(len({k.split(".")[3] for k in sd if k.startswith(f"layers.{i}.blocks.")}) for i in range(4))
    return embed, depths


def _pad_to(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def synthetic_clap_state(seed: int = 0, variant: str = "tiny") -> dict:
    """Seeded random parameters (float32, CPU); same recipe as oracle.clap_oracle.synthetic_state."""
    EMBED, DEPTHS = VARIANTS[variant]
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(key, out_f, in_f, bias=True):
        sd[key + ".weight"] = torch.randn((out_f, in_f), generator=g) * (1.0 / math.sqrt(in_f))
        if bias:
            sd[key + ".bias"] = torch.randn((out_f,), generator=g) * 0.02

    def ln(key, n):
        sd[key + ".weight"] = 1.0 + 0.1 * torch.randn((n,), generator=g)
        sd[key + ".bias"] = 0.05 * torch.randn((n,), generator=g)

    sd["batch_norm.weight"] = 1.0 + 0.1 * torch.randn((N_MEL,), generator=g)
    sd["batch_norm.bias"] = 0.1 * torch.randn((N_MEL,), generator=g)
    sd["batch_norm.running_mean"] = -30.0 + 5.0 * torch.randn((N_MEL,), generator=g)
    sd["batch_norm.running_var"] = 200.0 + 50.0 * torch.rand((N_MEL,), generator=g)
    sd["patch_embed.proj.weight"] = torch.randn((EMBED, 1, 4, 4), generator=g) * 0.25
    sd["patch_embed.proj.bias"] = torch.randn((EMBED,), generator=g) * 0.02
    ln("patch_embed.norm", EMBED)
    c = EMBED
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        for j in range(depth):
            p = f"layers.{i}.blocks.{j}."
            ln(p + "layernorm_before", c)
            for n in ("query", "key", "value"):
                lin(p + "attention.self." + n, c, c)
            sd[p + "attention.self.relative_position_bias_table"] = 0.2 * torch.randn(((2 * WINDOW - 1) ** 2, heads), generator=g)
            lin(p + "attention.output.dense", c, c)
            ln(p + "layernorm_after", c)
            lin(p + "intermediate.dense", 4 * c, c)
            lin(p + "output.dense", c, 4 * c)
        if i < len(DEPTHS) - 1:
            ln(f"layers.{i}.downsample.norm", 4 * c)
            lin(f"layers.{i}.downsample.reduction", 2 * c, 4 * c, bias=False)
            c *= 2
    ln("norm", c)
    lin("audio_projection.linear1", 512, c)
    lin("audio_projection.linear2", 512, 512)
    return sd


def load_clap_state(path=None, seed: int = 0, variant: str = "tiny") -> dict:
    """HF-format checkpoint if ``path`` (or $FADTK_CLAP_CKPT / $FADTK_CLAP_MUSIC_CKPT) exists, else synthetic."""
    from .weights import resolve_checkpoint
    path = resolve_checkpoint(path, "FADTK_CLAP_CKPT" if variant == "tiny" else "FADTK_CLAP_MUSIC_CKPT",
                              "clap-laion-" + ("audio" if variant == "tiny" else "music"))
    if path is not None:
        from .weights import load_checkpoint_file
        raw = load_checkpoint_file(path)
        out = {}
        for k, v in raw.items():
            k = k.replace("audio_model.audio_encoder.", "")
            if k.startswith(("layers.", "patch_embed.", "batch_norm.", "norm.", "audio_projection.")) \
                    and "relative_position_index" not in k and "num_batches_tracked" not in k:
                out[k] = v.float().contiguous()
        return out
    return synthetic_clap_state(seed, variant)


def _rel_pos_index(ws: int = WINDOW) -> torch.Tensor:
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _padded(w: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    out = torch.zeros((rows, cols), dtype=torch.float32)
    out[: w.shape[0], : w.shape[1]] = w
    return out


def _padded_vec(b: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,), dtype=torch.float32)
    out[: b.shape[0]] = b
    return out


def pack_clap(sd: dict) -> list:
    """-> the contiguous CPU tensors fad_clap_load expects (180 for HTSAT-tiny, 258 for HTSAT-base;
    order: csrc/clap_host.inc).

    GEMM weights are zero padded to K % 64 == 0 / N % 128 == 0 and stored as fp16 hi/lo tiles
    (22-bit weights, see weights.split_hi_lo_tiles); everything else stays float32."""
    f = lambda t: t.float().contiguous()
    EMBED, DEPTHS = config_of(sd)
    out = []
    scale = sd["batch_norm.weight"] / torch.sqrt(sd["batch_norm.running_var"] + 1e-5)
    out += [f(scale), f(sd["batch_norm.bias"] - sd["batch_norm.running_mean"] * scale)]
    out += [f(sd["patch_embed.proj.weight"].reshape(EMBED, 16)), f(sd["patch_embed.proj.bias"]),
            f(sd["patch_embed.norm.weight"]), f(sd["patch_embed.norm.bias"])]
    idx = _rel_pos_index().view(-1)
    c = EMBED
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        kp, nq, cp = _pad_to(c, 64), _pad_to(3 * c, 128), _pad_to(c, 128)
        for j in range(depth):
            p = f"layers.{i}.blocks.{j}."
            a = p + "attention.self."
            wqkv = torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]], 0)
            bqkv = torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], 0)
            rel = sd[a + "relative_position_bias_table"][idx].view(64, 64, heads).permute(2, 0, 1)
            out += [f(sd[p + "layernorm_before.weight"]), f(sd[p + "layernorm_before.bias"]),
                    split_hi_lo_tiles(_padded(wqkv, nq, kp)), _padded_vec(bqkv, nq), f(rel),
                    split_hi_lo_tiles(_padded(sd[p + "attention.output.dense.weight"], cp, kp)),
                    _padded_vec(sd[p + "attention.output.dense.bias"], cp),
                    f(sd[p + "layernorm_after.weight"]), f(sd[p + "layernorm_after.bias"]),
                    split_hi_lo_tiles(_padded(sd[p + "intermediate.dense.weight"], 4 * c, kp)),
                    f(sd[p + "intermediate.dense.bias"]),
                    split_hi_lo_tiles(_padded(sd[p + "output.dense.weight"], cp, 4 * c)),
                    _padded_vec(sd[p + "output.dense.bias"], cp)]
        c *= 2
    c = EMBED
    for i in range(3):
        npad = _pad_to(2 * c, 128)
        p = f"layers.{i}.downsample."
        out += [f(sd[p + "norm.weight"]), f(sd[p + "norm.bias"]),
                split_hi_lo_tiles(_padded(sd[p + "reduction.weight"], npad, 4 * c)), torch.zeros(npad)]
        c *= 2
    out += [f(sd["norm.weight"]), f(sd["norm.bias"]),
            f(sd["audio_projection.linear1.weight"]), f(sd["audio_projection.linear1.bias"]),
            f(sd["audio_projection.linear2.weight"]), f(sd["audio_projection.linear2.bias"])]
    assert len(out) == 6 + sum(DEPTHS) * 13 + 3 * 4 + 6
    return out
