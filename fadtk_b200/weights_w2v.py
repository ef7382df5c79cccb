"""wav2vec 2.0 / HuBERT / MERT parameters (group-norm feature encoder, post-LN transformer): seeded synthetic
set, checkpoint loading, device packing.  State-dict keys are those of ``transformers.Wav2Vec2Model`` /
``HubertModel`` (identical for these checkpoints), which is what the reference loads
(model_loader.py:262-265, 540-544, 578-582).  No checkpoint exists offline."""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch

from .weights import split_hi_lo_tiles

CONV_KERNEL, CONV_STRIDE, CONV_DIM = (10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2), 512
POS_K, POS_GROUPS = 128, 16
# registry family/size -> architecture (transformers configs of the checkpoints the reference names)
ARCH = {
    ("w2v2", "base"): dict(d=768, layers=12, ffn=3072, variant="group"),      # facebook/wav2vec2-base-960h
    ("w2v2", "large"): dict(d=1024, layers=24, ffn=4096, variant="group"),    # facebook/wav2vec2-large-960h
    ("hubert", "base"): dict(d=768, layers=12, ffn=3072, variant="group"),    # facebook/hubert-base-ls960
    ("hubert", "large"): dict(d=1024, layers=24, ffn=4096, variant="layer"),  # facebook/hubert-large-ls960: layer-norm convs, stable LN
    ("mert", "v1-95M"): dict(d=768, layers=12, ffn=3072, variant="group"),    # m-a-p/MERT-v1-95M (24 kHz)
    # patrickvonplaten/wavlm-libri-clean-100h-{base,base-plus,large}: WavLM adds a gated relative position bias
    ("wavlm", "base"): dict(d=768, layers=12, ffn=3072, variant="group", wavlm=True),
    ("wavlm", "base-plus"): dict(d=768, layers=12, ffn=3072, variant="group", wavlm=True),
    ("wavlm", "large"): dict(d=1024, layers=24, ffn=4096, variant="layer", wavlm=True),
}
WAVLM_BUCKETS = 320


def synthetic_w2v_state(seed: int = 0, d: int = 768, layers: int = 12, ffn: int = 3072, variant: str = "group",
                        wavlm: bool = False) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(key, out_f, in_f):
        sd[key + ".weight"] = torch.randn((out_f, in_f), generator=g) * (1.0 / math.sqrt(in_f))
        sd[key + ".bias"] = torch.randn((out_f,), generator=g) * 0.02

    def ln(key, n):
        sd[key + ".weight"] = 1.0 + 0.1 * torch.randn((n,), generator=g)
        sd[key + ".bias"] = 0.05 * torch.randn((n,), generator=g)

    cin = 1
    for i, k in enumerate(CONV_KERNEL):
        sd[f"feature_extractor.conv_layers.{i}.conv.weight"] = torch.randn((CONV_DIM, cin, k), generator=g) * math.sqrt(2.0 / (cin * k))
        cin = CONV_DIM
        if variant == "layer":                                 # conv bias + LayerNorm after every conv
            sd[f"feature_extractor.conv_layers.{i}.conv.bias"] = 0.02 * torch.randn((CONV_DIM,), generator=g)
            ln(f"feature_extractor.conv_layers.{i}.layer_norm", CONV_DIM)
    if variant == "group":
        ln("feature_extractor.conv_layers.0.layer_norm", CONV_DIM)
    ln("feature_projection.layer_norm", CONV_DIM)
    lin("feature_projection.projection", d, CONV_DIM)
    cg = d // POS_GROUPS
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = torch.randn((d, cg, POS_K), generator=g)
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = 2.0 * (0.8 + 0.4 * torch.rand((1, 1, POS_K), generator=g))
    sd["encoder.pos_conv_embed.conv.bias"] = 0.02 * torch.randn((d,), generator=g)
    ln("encoder.layer_norm", d)
    for i in range(layers):
        p = f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(p + "attention." + n, d, d)
        if wavlm:
            heads = d // 64
            sd[p + "attention.gru_rel_pos_linear.weight"] = torch.randn((8, 64), generator=g) * 0.125
            sd[p + "attention.gru_rel_pos_linear.bias"] = 0.1 * torch.randn((8,), generator=g)
            sd[p + "attention.gru_rel_pos_const"] = 1.0 + 0.2 * torch.randn((1, heads, 1, 1), generator=g)
            if i == 0:
                sd[p + "attention.rel_attn_embed.weight"] = 0.5 * torch.randn((WAVLM_BUCKETS, heads), generator=g)
        ln(p + "layer_norm", d)
        lin(p + "feed_forward.intermediate_dense", ffn, d)
        lin(p + "feed_forward.output_dense", d, ffn)
        ln(p + "final_layer_norm", d)
    return sd


def load_w2v_state(path=None, seed: int = 0, env: str = "FADTK_W2V_CKPT", **cfg) -> dict:
    from .weights import resolve_checkpoint
    path = resolve_checkpoint(path, env, env.removeprefix("FADTK_").removesuffix("_CKPT").lower())
    if path is not None:
        from .weights import load_checkpoint_file
        raw = load_checkpoint_file(path)                      # .safetensors or torch pickle; weight_g / weight_v renamed
        drop = ("masked_spec_embed", "lm_head", "quantizer", "project_", "label_embs")
        return {k.removeprefix("wav2vec2.").removeprefix("hubert.").removeprefix("wavlm."): v.float().contiguous() for k, v in raw.items()
                if not any(x in k for x in drop)}
    return synthetic_w2v_state(seed, **cfg)


def variant_of(sd: dict) -> str:
    """"layer": LayerNorm after every feature-encoder conv (those checkpoints also use the stable-LN transformer)."""
    return "layer" if "feature_extractor.conv_layers.1.layer_norm.weight" in sd else "group"


def config_of(sd: dict) -> tuple:
    """(d_model, heads, layers, ffn, layer-norm convs, stable layer norm, WavLM)"""
    d = sd["feature_projection.projection.weight"].shape[0]
    layers = len({k.split(".")[2] for k in sd if k.startswith("encoder.layers.")})
    lay = int(variant_of(sd) == "layer")
    wavlm = int("encoder.layers.0.attention.rel_attn_embed.weight" in sd)
    return d, d // 64, layers, sd["encoder.layers.0.feed_forward.intermediate_dense.weight"].shape[0], lay, lay, wavlm


def pos_conv_weight(sd: dict) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / ||v|| with the norm over (out, in) per tap -> [d, d/16, 128]"""
    v = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    g_ = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    return g_ * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()


def _pad_to(v, m):
    return (v + m - 1) // m * m


def pack_w2v(sd: dict) -> list:
    d, heads, layers, ffn = config_of(sd)[:4]
    fl = lambda t: t.float().contiguous()
    out = []
    for i, k in enumerate(CONV_KERNEL):
        w = sd[f"feature_extractor.conv_layers.{i}.conv.weight"]                      # [512, Cin, k]
        cout, cin, _ = w.shape
        m = torch.zeros((cout, _pad_to(k * cin, 64)))
        m[:, :k * cin] = w.permute(0, 2, 1).reshape(cout, k * cin)                     # column = tap*Cin + c
        b = sd.get(f"feature_extractor.conv_layers.{i}.conv.bias", torch.zeros(cout))
        ng = sd.get(f"feature_extractor.conv_layers.{i}.layer_norm.weight", torch.ones(cout))
        nb = sd.get(f"feature_extractor.conv_layers.{i}.layer_norm.bias", torch.zeros(cout))
        out += [split_hi_lo_tiles(m), fl(b), fl(ng), fl(nb)]
    out += [fl(sd["feature_projection.layer_norm.weight"]), fl(sd["feature_projection.layer_norm.bias"]),
            split_hi_lo_tiles(fl(sd["feature_projection.projection.weight"])), fl(sd["feature_projection.projection.bias"])]
    wp = pos_conv_weight(sd)
    cg = d // POS_GROUPS
    for g in range(POS_GROUPS):
        wg = torch.zeros((128, POS_K * cg))
        wg[:cg] = wp[g * cg:(g + 1) * cg].permute(0, 2, 1).reshape(cg, POS_K * cg)    # column = tap*cg + ci
        out.append(split_hi_lo_tiles(wg))
    out += [fl(sd["encoder.pos_conv_embed.conv.bias"]), fl(sd["encoder.layer_norm.weight"]), fl(sd["encoder.layer_norm.bias"])]
    wavlm = config_of(sd)[6]
    if wavlm:
        out.append(fl(sd["encoder.layers.0.attention.rel_attn_embed.weight"]))               # [320, heads]
    for i in range(layers):
        p = f"encoder.layers.{i}."
        a = p + "attention."
        qkv_w = torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)
        qkv_b = torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)
        out += [split_hi_lo_tiles(qkv_w), fl(qkv_b), split_hi_lo_tiles(fl(sd[a + "out_proj.weight"])), fl(sd[a + "out_proj.bias"]),
                fl(sd[p + "layer_norm.weight"]), fl(sd[p + "layer_norm.bias"]),
                split_hi_lo_tiles(fl(sd[p + "feed_forward.intermediate_dense.weight"])), fl(sd[p + "feed_forward.intermediate_dense.bias"]),
                split_hi_lo_tiles(fl(sd[p + "feed_forward.output_dense.weight"])), fl(sd[p + "feed_forward.output_dense.bias"]),
                fl(sd[p + "final_layer_norm.weight"]), fl(sd[p + "final_layer_norm.bias"])]
        if wavlm:
            out += [fl(sd[a + "gru_rel_pos_linear.weight"]), fl(sd[a + "gru_rel_pos_linear.bias"]), fl(sd[a + "gru_rel_pos_const"].flatten())]
    assert len(out) == 28 + 4 + 17 + 2 + (1 + 15 * layers if wavlm else 12 * layers)
    return out
