"""ctypes binding of libfadtk_io.so (include/fadtk_b200_io.h): batched, multi-threaded WAV / .npy I/O.

Host-side only - it loads and runs without a GPU.  The directory flow (fad_batch.cache_embedding_files,
FrechetAudioDistance.score_individual / score_inf, utils.calculate_embd_statistics_online) uses it to read a
whole batch of clips into ONE pinned buffer and to write the convert cache and the fp16 ``.npy`` embedding
cache on native threads; files it cannot take (non-PCM16 WAVs, other containers, other dtypes) are reported by
status code and go through the per-file Python path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).parent / "csrc" / "libfadtk_io.so"
_lib = None

OK, EOPEN, EFORMAT, EUNSUPPORTED, ESHORT = 0, 1, 2, 3, 4

_pp = C.POINTER(C.c_char_p)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
SIGNATURES = {
    "fad_io_version": (C.c_int, []),
    "fad_io_wav_probe": (C.c_int, [_pp, C.c_int, C.c_int, _ip, _ip, _lp, _ip]),
    "fad_io_wav_read": (C.c_int, [_pp, C.c_int, C.c_int, C.c_void_p, _lp, _lp, _ip, _ip]),
    "fad_io_wav_write": (C.c_int, [_pp, C.c_int, C.c_int, C.c_void_p, _lp, _lp, C.c_int, _ip]),
    "fad_io_npy_write_f16": (C.c_int, [_pp, C.c_int, C.c_int, C.c_void_p, _lp, _lp, C.c_int, _ip]),
    "fad_io_npy_probe": (C.c_int, [_pp, C.c_int, C.c_int, _lp, _ip, _ip, _ip, _ip]),
    "fad_io_npy_read_f16": (C.c_int, [_pp, C.c_int, C.c_int, C.c_void_p, _lp, _lp, C.c_int, _ip]),
}


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def _paths(paths):
    enc = [os.fsencode(str(p)) for p in paths]
    return (C.c_char_p * len(enc))(*enc), len(enc)


def _i32(n):
    return np.zeros(n, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a, t):
    return a.ctypes.data_as(t)


def wav_probe(paths, threads: int = 0):
    """-> (sample_rate int32[n], channels int32[n], frames int64[n], status int32[n])"""
    arr, n = _paths(paths)
    sr, ch, st = _i32(n), _i32(n), _i32(n)
    fr = np.zeros(n, dtype=np.int64)
    if n and lib().fad_io_wav_probe(arr, n, threads, _p(sr, _ip), _p(ch, _ip), _p(fr, _lp), _p(st, _ip)) < 0:
        raise ValueError("fad_io_wav_probe: bad arguments")
    return sr, ch, fr, st


def wav_read(paths, frames, channels, out: np.ndarray, offsets=None, threads: int = 0):
    """Samples of PCM16 files into ``out`` (int16, 1-D, C-contiguous - e.g. the numpy view of a pinned torch
    tensor).  ``offsets`` defaults to back-to-back placement.  -> (offsets int64[n+1 or n], status)."""
    arr, n = _paths(paths)
    frames, channels = _i64(frames), np.ascontiguousarray(channels, dtype=np.int32)
    if offsets is None:
        offsets = np.zeros(n + 1, dtype=np.int64)
        offsets[1:] = np.cumsum(frames * channels)
    offsets = _i64(offsets)
    assert out.dtype == np.int16 and out.flags.c_contiguous
    assert n == 0 or int((offsets[:n] + frames * channels).max()) <= out.size
    st = _i32(n)
    if n and lib().fad_io_wav_read(arr, n, threads, out.ctypes.data, _p(offsets, _lp), _p(frames, _lp), _p(channels, _ip), _p(st, _ip)) < 0:
        raise ValueError("fad_io_wav_read: bad arguments")
    return offsets, st


def wav_write(paths, src: np.ndarray, offsets, frames, sample_rate: int, threads: int = 0):
    """Mono PCM16 WAVs: file i = src[offsets[i] : offsets[i] + frames[i]].  -> status"""
    arr, n = _paths(paths)
    offsets, frames = _i64(offsets), _i64(frames)
    assert src.dtype == np.int16 and src.flags.c_contiguous
    assert n == 0 or int((offsets[:n] + frames).max()) <= src.size
    st = _i32(n)
    if n and lib().fad_io_wav_write(arr, n, threads, src.ctypes.data, _p(offsets, _lp), _p(frames, _lp), int(sample_rate), _p(st, _ip)) < 0:
        raise ValueError("fad_io_wav_write: bad arguments")
    return st


def npy_write_f16(paths, src: np.ndarray, row_offsets, rows, threads: int = 0):
    """fp16 [rows[i], d] ``.npy`` files (byte-identical to np.save) from a row-major fp16 [N, d] array.  -> status"""
    arr, n = _paths(paths)
    row_offsets, rows = _i64(row_offsets), _i64(rows)
    assert src.dtype == np.float16 and src.ndim == 2 and src.flags.c_contiguous
    assert n == 0 or int((row_offsets[:n] + rows).max()) <= src.shape[0]
    st = _i32(n)
    if n and lib().fad_io_npy_write_f16(arr, n, threads, src.ctypes.data, _p(row_offsets, _lp), _p(rows, _lp), int(src.shape[1]), _p(st, _ip)) < 0:
        raise ValueError("fad_io_npy_write_f16: bad arguments")
    return st


def npy_probe(paths, threads: int = 0):
    """-> (rows int64[n], cols int32[n], ndim int32[n], itemsize-coded float dtype int32[n], status)"""
    arr, n = _paths(paths)
    rows = np.zeros(n, dtype=np.int64)
    cols, nd, dt, st = _i32(n), _i32(n), _i32(n), _i32(n)
    if n and lib().fad_io_npy_probe(arr, n, threads, _p(rows, _lp), _p(cols, _ip), _p(nd, _ip), _p(dt, _ip), _p(st, _ip)) < 0:
        raise ValueError("fad_io_npy_probe: bad arguments")
    return rows, cols, nd, dt, st


def npy_read_f16(paths, rows, d: int, out: np.ndarray, row_offsets=None, threads: int = 0):
    """fp16 [rows[i], d] files into ``out`` (fp16 [N, d]).  -> (row_offsets, status)"""
    arr, n = _paths(paths)
    rows = _i64(rows)
    if row_offsets is None:
        row_offsets = np.zeros(n + 1, dtype=np.int64)
        row_offsets[1:] = np.cumsum(rows)
    row_offsets = _i64(row_offsets)
    assert out.dtype == np.float16 and out.ndim == 2 and out.shape[1] == d and out.flags.c_contiguous
    assert n == 0 or int((row_offsets[:n] + rows).max()) <= out.shape[0]
    st = _i32(n)
    if n and lib().fad_io_npy_read_f16(arr, n, threads, out.ctypes.data, _p(row_offsets, _lp), _p(rows, _lp), int(d), _p(st, _ip)) < 0:
        raise ValueError("fad_io_npy_read_f16: bad arguments")
    return row_offsets, st


def plan_embedding_chunks(files, max_bytes: int = 1 << 30, max_files: int = 4096, threads: int = 0):
    """Consecutive runs of ``files`` whose payloads (as far as the .npy headers tell) stay below ``max_bytes`` each -
    a single larger file gets a run of its own."""
    files = list(files)
    rows, cols, _, dt, st = npy_probe(files, threads)
    size = np.where(st == OK, rows * cols * np.maximum(dt, 1), 0)
    chunks, cur, tot = [], [], 0
    for f, b in zip(files, size):
        if cur and (len(cur) >= max_files or tot + int(b) > max_bytes):
            chunks.append(cur)
            cur, tot = [], 0
        cur.append(f)
        tot += int(b)
    if cur:
        chunks.append(cur)
    return chunks


def load_embedding_files(files, threads: int = 0):
    """Ragged concatenation of fp16 2-D ``.npy`` embedding caches: -> (fp16 [N, d], row offsets int64[n+1]).
    Files that are not C-ordered fp16 2-D arrays of one common width are read with np.load and converted the
    way the reference's ``np.load`` + ``np.concatenate`` would see them (must still agree on the width)."""
    files = [Path(f) for f in files]
    n = len(files)
    rows, cols, nd, dt, st = npy_probe(files, threads)
    fast = (st == OK) & (dt == 2) & (nd == 2)
    slow = {}
    for i in np.nonzero(~fast)[0]:
        a = np.load(files[i])
        if a.ndim != 2:
            raise ValueError(f"{files[i]}: expected a 2-D embedding array, got shape {a.shape}")
        slow[int(i)] = a
        rows[i], cols[i] = a.shape
    if n == 0:
        return np.zeros((0, 0), dtype=np.float16), np.zeros(1, dtype=np.int64)
    d = int(cols[0])
    if not np.all(cols == d):
        raise ValueError("embedding files disagree on the feature dimension")
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(rows)
    dtype = np.float16 if all(a.dtype == np.float16 for a in slow.values()) else np.result_type(np.float16, *[a.dtype for a in slow.values()])
    out16 = np.empty((int(off[-1]), d), dtype=np.float16)
    idx = np.nonzero(fast)[0]
    if len(idx):
        _, st2 = npy_read_f16([files[i] for i in idx], rows[idx], d, out16, off[idx], threads)
        for j in np.nonzero(st2 != OK)[0]:
            raise OSError(f"{files[idx[j]]}: read failed (status {int(st2[j])})")
    out = out16 if dtype == np.float16 else out16.astype(dtype)
    for i, a in slow.items():
        out[off[i]:off[i + 1]] = a
    return out, off
