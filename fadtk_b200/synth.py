"""Deterministic synthetic audio.

PCM16 mono at the model's sample rate, ``round(32767 * clip(x, -1, 1))``.  Used by
the tests and bench.py; there is no network for real datasets.
"""
from __future__ import annotations

import wave as _wave
from pathlib import Path

import numpy as np


def to_pcm16(x: np.ndarray) -> np.ndarray:
    return np.round(32767.0 * np.clip(x, -1.0, 1.0)).astype(np.int16)


def sine_clip(i: int, seconds: float, sr: int) -> np.ndarray:
    t = np.arange(int(round(seconds * sr))) / sr
    f = 110.0 * 2.0 ** ((i % 48) / 12.0)
    return to_pcm16(0.5 * np.sin(2 * np.pi * f * t))


def noise_clip(i: int, seconds: float, sr: int) -> np.ndarray:
    rng = np.random.default_rng(10_000 + i)
    return to_pcm16(rng.normal(0.0, 0.1, int(round(seconds * sr))))


def musiclike_clip(i: int, seconds: float, sr: int, baseline: bool = False) -> np.ndarray:
    rng = np.random.default_rng((30_000 if baseline else 20_000) + i)
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    freqs = rng.uniform(80.0, 4000.0, 4)
    amps = rng.uniform(0.05, 0.25, 4)
    x = (amps[:, None] * np.sin(2 * np.pi * freqs[:, None] * t[None, :])).sum(0)
    return to_pcm16(x + rng.normal(0.0, 0.02, n))


def clip_set(kind: str, count: int, seconds: float, sr: int, start: int = 0) -> np.ndarray:
    """[count, samples] int16."""
    gen = {"sine": sine_clip, "noise": noise_clip,
           "music": lambda i, s, r: musiclike_clip(i, s, r, False),
           "music-baseline": lambda i, s, r: musiclike_clip(i, s, r, True)}[kind]
    return np.stack([gen(start + i, seconds, sr) for i in range(count)])


def musiclike_device(count: int, seconds: float, sr: int, seed: int, device, chunk: int = 512,
                     fmax: float = 4000.0, noise: float = 0.02):
    """Large music-like set generated on the GPU (bench only): [count, samples] int16.

    Same recipe as ``musiclike_clip`` (4 sines 80-4000 Hz, amplitude 0.05-0.25, plus
    N(0, 0.02^2)) but drawn from a torch generator so 10 000 x 10 s clips take a
    fraction of a second.  The CPU baseline leg reads the same tensor back.  ``fmax`` / ``noise``
    change the timbre (the bench draws its baseline set darker and noisier than the eval set, so
    the FAD it reports is a distance between different distributions, as in real use).
    """
    import torch
    n = int(round(seconds * sr))
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((count, n), dtype=torch.int16, device=device)
    t = torch.arange(n, device=device, dtype=torch.float32) / sr
    for s in range(0, count, chunk):
        c = min(chunk, count - s)
        freqs = 80.0 + (fmax - 80.0) * torch.rand((c, 4), generator=g, device=device)
        amps = 0.05 + 0.20 * torch.rand((c, 4), generator=g, device=device)
        x = noise * torch.randn((c, n), generator=g, device=device)
        for k in range(4):
            x += amps[:, k:k + 1] * torch.sin(2 * np.pi * freqs[:, k:k + 1] * t[None, :])
        out[s:s + c] = torch.round(32767.0 * x.clamp_(-1.0, 1.0)).to(torch.int16)
    return out


def write_wav(path, pcm16: np.ndarray, sr: int) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with _wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm16, dtype="<i2").tobytes())


def read_wav(path):
    """-> (int16 [T] mono or [T, ch], sample_rate).  PCM16 RIFF only."""
    with _wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2")
    if ch > 1:
        data = data.reshape(-1, ch)
    return data, sr
