"""Batch embedding driver (mirror of fadtk/fad_batch.py:15-48).

The reference shards the file list over ``workers`` spawn processes, each loading its own copy
of the model on cuda:0 and looping file by file at batch size one.  Here one process owns one
GPU: ``workers`` host threads decode / convert audio and write ``.npy`` files, while clips are
packed into large batches for the sm_100a forward.  Under torchrun (one rank per GPU) the file
list is sharded across ranks the way the reference shards it across processes.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Union

import numpy as np

from . import dist
from .fad import FrechetAudioDistance, log
from .model_loader import ModelLoader
from .utils import get_cache_embedding_path

# clips per GPU launch sequence: bounded by audio seconds so ragged sets keep batches even
_BATCH_AUDIO_SECONDS = 4096.0


def _batches(files, lengths_s, limit_s):
    cur, tot = [], 0.0
    for f, s in zip(files, lengths_s):
        if cur and tot + s > limit_s:
            yield cur
            cur, tot = [], 0.0
        cur.append(f)
        tot += s
    if cur:
        yield cur


def cache_embedding_files(files: Union[list[Path], str, Path], ml: ModelLoader, workers: int = 8, **kwargs):
    """Get embeddings for all audio files in a directory (or list) and cache them as
    ``<dir>/embeddings/<model>/<stem>.npy`` (fp16 [n_frames, d]), skipping files already done.

    Same signature as the reference; ``kwargs`` are forwarded to FrechetAudioDistance.
    """
    if isinstance(files, (str, Path)):
        files = list(Path(files).glob('*.*'))

    files = [Path(f) for f in files if not get_cache_embedding_path(ml.name, f).exists()]
    if len(files) == 0:
        log.info("All files already have embeddings, skipping.")
        return

    files = list(dist.shard(sorted(files)))
    log.info(f"[Frechet Audio Distance] Loading {len(files)} audio files...")

    kwargs.setdefault("audio_load_worker", workers)
    fad = FrechetAudioDistance(ml, **kwargs)
    workers = max(1, int(workers))

    def save(item):
        f, embd = item
        cache = get_cache_embedding_path(ml.name, f)
        cache.parent.mkdir(parents=True, exist_ok=True)
        np.save(cache, embd)

    with ThreadPoolExecutor(workers) as pool:
        # decode ahead in chunks so host I/O overlaps the GPU
        chunk = 2048
        pending = None
        for s in range(0, len(files), chunk):
            part = files[s:s + chunk]
            clips = list(pool.map(fad.convert_audio, part))
            secs = [len(c) / ml.sr for c in clips]
            by_file = dict(zip(part, clips))
            for group in _batches(part, secs, _BATCH_AUDIO_SECONDS):
                embs = ml.embed_pcm_batch([by_file[f] for f in group])
                if pending is not None:
                    list(pending)
                pending = pool.map(save, list(zip(group, embs)))
        if pending is not None:
            list(pending)
    dist.barrier()
