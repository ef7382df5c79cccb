"""Encodec-24 kHz SEANet encoder parameters: seeded synthetic set, checkpoint loading, device packing.

State-dict keys are those of ``transformers.EncodecModel(...).encoder`` (weight-normalised convs stored as
``parametrizations.weight.original0`` = g and ``original1`` = v; LSTM ``weight_ih_l*`` ...), the HF port of
facebookresearch/encodec's ``EncodecModel.encodec_model_24khz().encoder`` that the reference calls
(model_loader.py:123-130, 160).  No checkpoint exists offline: tests and benches use seeded synthetic
parameters with the real architecture.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch

from .weights import split_hi_lo_tiles

RATIOS = (2, 4, 5, 8)          # encoder order (upsampling_ratios reversed)
N_FILTERS, DIM, LSTM_LAYERS = 32, 128, 2
# (HF layer index, kind, Cin, Cout, kernel, stride) in execution order
def conv_table():
    t = [(0, "in", 1, N_FILTERS, 7, 1)]
    ch, idx = N_FILTERS, 1
    for r in RATIOS:
        t += [(idx, "res", ch, ch, 3, 1), (idx + 2, "down", ch, 2 * ch, 2 * r, r)]
        ch, idx = 2 * ch, idx + 3
    t.append((idx + 2, "out", ch, DIM, 7, 1))       # idx = 13 (LSTM), 14 ELU, 15 conv
    return t


def _conv_keys(prefix):
    return prefix + ".conv.parametrizations.weight.original0", prefix + ".conv.parametrizations.weight.original1", prefix + ".conv.bias"


def synthetic_encodec_state(seed: int = 0, variant: str = "24k") -> dict:
    """24k: causal, weight-normalised convs, mono.  48k: non-causal, plain convs followed by GroupNorm(1, C)
    ("time_group_norm"), stereo input (encodec_model_48khz, model_loader.py:124-126)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, cout, cin, k):
        if variant == "48k":
            sd[prefix + ".conv.weight"] = torch.randn((cout, cin, k), generator=g) * math.sqrt(2.0 / (cin * k))
            sd[prefix + ".conv.bias"] = 0.02 * torch.randn((cout,), generator=g)
            sd[prefix + ".norm.weight"] = 1.0 + 0.1 * torch.randn((cout,), generator=g)
            sd[prefix + ".norm.bias"] = 0.05 * torch.randn((cout,), generator=g)
            return
        kg, kv, kb = _conv_keys(prefix)
        v = torch.randn((cout, cin, k), generator=g)
        sd[kv] = v
        sd[kg] = (math.sqrt(2.0) * (0.8 + 0.4 * torch.rand((cout, 1, 1), generator=g)))      # |w| ~ sqrt(2) per row: unit-gain-ish
        sd[kb] = 0.02 * torch.randn((cout,), generator=g)

    for idx, kind, cin, cout, k, s in conv_table():
        if kind == "in" and variant == "48k":
            cin = 2
        if kind == "res":
            conv(f"layers.{idx}.block.1", cin // 2, cin, 3)
            conv(f"layers.{idx}.block.3", cin, cin // 2, 1)
            conv(f"layers.{idx}.shortcut", cin, cin, 1)
        else:
            conv(f"layers.{idx}", cout, cin, k)
    h = 16 * N_FILTERS
    for l in range(LSTM_LAYERS):
        for name, shape in (("weight_ih", (4 * h, h)), ("weight_hh", (4 * h, h)), ("bias_ih", (4 * h,)), ("bias_hh", (4 * h,))):
            sd[f"layers.13.lstm.{name}_l{l}"] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(h)
    return sd


def load_encodec_state(path=None, seed: int = 0, variant: str = "24k") -> dict:
    from .weights import resolve_checkpoint
    path = resolve_checkpoint(path, "FADTK_ENCODEC_CKPT" if variant == "24k" else "FADTK_ENCODEC48_CKPT",
                              "encodec-emb" + ("" if variant == "24k" else "-48k"))
    if path is not None:
        from .weights import load_checkpoint_file
        raw = load_checkpoint_file(path)
        return {k.removeprefix("encoder."): v.float().contiguous() for k, v in raw.items() if "layers." in k and not k.startswith(("decoder.", "quantizer."))}
    return synthetic_encodec_state(seed, variant)


def variant_of(sd: dict) -> str:
    return "48k" if "layers.0.conv.weight" in sd else "24k"


def effective_weight(sd: dict, prefix: str) -> torch.Tensor:
    """weight_norm: w = g * v / ||v|| with the norm over (in, k) per output channel -> [Cout, Cin, k]
    (the 48 kHz model stores plain weights)"""
    if prefix + ".conv.weight" in sd:
        return sd[prefix + ".conv.weight"]
    kg, kv, _ = _conv_keys(prefix)
    v = sd[kv]
    return sd[kg] * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def _pad_to(v, m):
    return (v + m - 1) // m * m


def _gemm_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k] -> fp16 hi/lo tiles [2*Npad, Kpad], column = tap*Cin + c (the im2col order)"""
    cout, cin, k = w.shape
    m = torch.zeros((_pad_to(cout, 128), _pad_to(k * cin, 64)))
    m[:cout, :k * cin] = w.permute(0, 2, 1).reshape(cout, k * cin)
    return split_hi_lo_tiles(m)


def _bias(b, n):
    out = torch.zeros((_pad_to(n, 128),))
    out[:b.shape[0]] = b
    return out


def time_pack(cout: int) -> int:
    """outputs computed per GEMM row: thin layers (Cout < 128) pack P consecutive time steps into one row so the
    128-wide tensor-core tile is full (P * Cout = 128)"""
    return max(1, 128 // cout) if cout < 128 else 1


def _packed_weight(w: torch.Tensor, stride: int, P: int) -> torch.Tensor:
    """[Cout, Cin, k] -> banded [P*Cout, (k + (P-1) stride) * Cin]: row p*Cout + co, column j*Cin + ci holds
    w[co, ci, j - p*stride] (P outputs of the strided conv from one window of k + (P-1) stride inputs)"""
    cout, cin, k = w.shape
    kp = k + (P - 1) * stride
    m = torch.zeros((P * cout, kp, cin))
    wt = w.permute(0, 2, 1)                                            # [Cout, k, Cin]
    for p_ in range(P):
        m[p_ * cout:(p_ + 1) * cout, p_ * stride:p_ * stride + k, :] = wt
    full = torch.zeros((_pad_to(P * cout, 128), _pad_to(kp * cin, 64)))
    full[:P * cout, :kp * cin] = m.reshape(P * cout, kp * cin)
    return split_hi_lo_tiles(full)


def pack_encodec(sd: dict) -> list:
    """-> contiguous CPU tensors in the order fad_encodec_load expects (csrc/encodec_host.inc):
    per conv (execution order; a residual block contributes conv3, conv1, shortcut): weight tiles, bias, GroupNorm
    weight, GroupNorm bias (ones / zeros for the 24 kHz model, which has no norm layers), time-packed weight tiles
    and bias (see time_pack; identical to the plain ones when P = 1);
    then per LSTM layer: W_ih tiles [2048, 512], W_hh tiles over [h_hi | h_lo] = [2048, 1024], bias_ih + bias_hh."""
    out = []
    for idx, kind, cin, cout, k, s in conv_table():
        names = [f"layers.{idx}.block.1", f"layers.{idx}.block.3", f"layers.{idx}.shortcut"] if kind == "res" else [f"layers.{idx}"]
        for p in names:
            w = effective_weight(sd, p)
            cout = w.shape[0]
            P = time_pack(cout)
            out += [_gemm_weight(w), _bias(sd[p + ".conv.bias"], cout),
                    sd.get(p + ".norm.weight", torch.ones(cout)).float().contiguous(),
                    sd.get(p + ".norm.bias", torch.zeros(cout)).float().contiguous(),
                    _packed_weight(w, s, P), _bias(sd[p + ".conv.bias"].repeat(P), P * cout)]
    for l in range(LSTM_LAYERS):
        wih, whh = sd[f"layers.13.lstm.weight_ih_l{l}"], sd[f"layers.13.lstm.weight_hh_l{l}"]
        out += [split_hi_lo_tiles(wih.contiguous()), split_hi_lo_tiles(torch.cat([whh, whh], 1).contiguous()),
                (sd[f"layers.13.lstm.bias_ih_l{l}"] + sd[f"layers.13.lstm.bias_hh_l{l}"]).float().contiguous()]
    return out
