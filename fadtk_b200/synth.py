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


def read_wav_float(path):
    """Any uncompressed RIFF/WAVE sample format -> (float32 [channels, T] in [-1, 1), sample_rate), normalised the
    way ``torchaudio.load`` hands it to the reference's load_audio (fadtk/fad.py:147): 8-bit unsigned, 16 / 24 / 32-bit
    signed PCM (integer / 2^(bits-1)), 32 / 64-bit IEEE float (as is), plain or WAVE_FORMAT_EXTENSIBLE headers.
    Python's ``wave`` module reads only integer PCM with a plain header; this covers what music datasets ship."""
    data = Path(path).read_bytes()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    at, fmt, payload = 12, None, None
    while at + 8 <= len(data):
        tag, size = data[at:at + 4], int.from_bytes(data[at + 4:at + 8], "little")
        body = at + 8
        if tag == b"fmt ":
            if size < 16:
                raise ValueError(f"{path}: short fmt chunk")
            code = int.from_bytes(data[body:body + 2], "little")
            ch = int.from_bytes(data[body + 2:body + 4], "little")
            sr = int.from_bytes(data[body + 4:body + 8], "little")
            block = int.from_bytes(data[body + 12:body + 14], "little")
            bits = int.from_bytes(data[body + 14:body + 16], "little")
            if code == 0xFFFE and size >= 26:                  # extensible: the sub-format GUID starts with the real code
                code = int.from_bytes(data[body + 24:body + 26], "little")
            fmt = (code, ch, sr, block, bits)
        elif tag == b"data":
            if size in (0, 0xFFFFFFFF) or body + size > len(data):
                size = len(data) - body                        # streamed writers leave the size open
            payload = data[body:body + size]
            break
        at = body + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: fmt or data chunk missing")
    code, ch, sr, block, bits = fmt
    if ch < 1 or sr < 1 or block != ch * bits // 8:
        raise ValueError(f"{path}: inconsistent fmt chunk")
    n = len(payload) // block
    raw = np.frombuffer(payload, dtype=np.uint8, count=n * block)
    if code == 1 and bits == 8:
        x = (raw.astype(np.float32) - 128.0) / 128.0
    elif code == 1 and bits == 16:
        x = raw.view("<i2").astype(np.float32) / 32768.0
    elif code == 1 and bits == 24:
        b = raw.reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / float(1 << 23)
    elif code == 1 and bits == 32:
        x = (raw.view("<i4").astype(np.float64) / float(1 << 31)).astype(np.float32)
    elif code == 3 and bits == 32:
        x = raw.view("<f4").astype(np.float32)
    elif code == 3 and bits == 64:
        x = raw.view("<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV sample format (code {code}, {bits} bits)")
    return np.ascontiguousarray(x.reshape(n, ch).T), sr

