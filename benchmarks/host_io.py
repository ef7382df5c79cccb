"""Host side of the directory flow, measured without a GPU: how fast can clips get from .wav files to the
embedder and embeddings back into the .npy cache?  (three filesystem round trips per
clip in the reference.)

Legs, on N synthetic 10-s PCM16 clips at 16 kHz with a stub embedder ([10, 128] fp16 per clip):
  python : the per-file flow on `workers` threads - wave.open read, convert-cache write, np.save
  native : fad_batch.cache_embedding_files (libfadtk_io.so: batched reads into one pinned buffer, hard-linked
           convert cache, batched .npy writes), first pass and with the convert cache already present
  npy    : reading the embedding caches back: np.load loop + concatenate vs _io_native.load_embedding_files
Prints one JSON object.  Usage: python benchmarks/host_io.py [--clips 4000] [--workers 8] [--dir /tmp/fadtk_host_io]
"""
import argparse
import json
import os
import shutil
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fadtk_b200 import _io_native, fad_batch, synth  # noqa: E402
from fadtk_b200.model_loader import ModelLoader  # noqa: E402


class Stub(ModelLoader):
    def __init__(self):
        super().__init__("stub", 128, 16000)
        self.emb = np.zeros((10, 128), dtype=np.float16)

    def load_model(self):
        pass

    def _get_embedding(self, audio):
        raise NotImplementedError

    def embed_pcm_batch(self, clips):
        return [self.emb for _ in clips]


def python_flow(files, workers):
    ml = Stub()

    def one(f):
        pcm, sr = synth.read_wav(f)
        conv = f.parent / "convert" / str(sr) / f.name
        conv.parent.mkdir(parents=True, exist_ok=True)
        synth.write_wav(conv, pcm, sr)
        emb = ml.embed_pcm_batch([pcm])[0]
        out = f.parent / "embeddings" / ml.name / (f.stem + ".npy")
        out.parent.mkdir(parents=True, exist_ok=True)
        np.save(out, emb)

    with ThreadPoolExecutor(workers) as pool:
        list(pool.map(one, files))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4000)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--dir", default="/tmp/fadtk_host_io")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    root = Path(args.dir) / "set"
    root.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(0)
    files = [root / f"c{i:06d}.wav" for i in range(args.clips)]
    for f in files:
        if not f.exists():
            synth.write_wav(f, (rng.standard_normal(160000) * 3000).astype(np.int16), 16000)
    audio_s = args.clips * 10.0

    def clean(convert=True):
        shutil.rmtree(root / "embeddings", ignore_errors=True)
        if convert:
            shutil.rmtree(root / "convert", ignore_errors=True)

    def timed(fn):
        os.sync()
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    res = {"clips": args.clips, "workers": args.workers, "cores": os.cpu_count(), "clip_seconds": 10.0, "sample_rate": 16000,
           "repeats": args.repeats, "statistic": "median of the repeats (shared disks are noisy); best in *_best"}

    def leg(name, convert, fn):
        ts = []
        for _ in range(args.repeats):
            clean(convert=convert)
            ts.append(timed(fn))
        t = float(np.median(ts))
        res[name] = {"files_per_s": args.clips / t, "audio_s_per_s": audio_s / t, "files_per_s_best": args.clips / min(ts)}

    native = lambda: fad_batch.cache_embedding_files(root, Stub(), workers=args.workers, load_model=False)  # noqa: E731
    leg("python_per_file", True, lambda: python_flow(files, args.workers))
    leg("native_first_pass", True, native)
    clean()
    native()                                                   # leaves the convert cache behind for the next leg
    leg("native_convert_cached", False, native)

    # embedding caches of per-song size ([750, 128], encodec-emb) read back for --indiv / statistics
    songs = Path(args.dir) / "songs"
    songs.mkdir(parents=True, exist_ok=True)
    n_songs = min(args.clips, 5000)
    paths = [songs / f"s{i:05d}.npy" for i in range(n_songs)]
    if not paths[-1].exists():
        emb = np.tile(rng.standard_normal((750, 128)).astype(np.float16), (n_songs, 1))
        _io_native.npy_write_f16(paths, emb, np.arange(n_songs) * 750, np.full(n_songs, 750), args.workers)
    t_np = timed(lambda: np.concatenate([np.load(p) for p in paths]))
    t_nat = timed(lambda: _io_native.load_embedding_files(paths, args.workers))
    res["npy_read"] = {"songs": n_songs, "rows_per_song": 750, "numpy_loop_s": t_np, "native_s": t_nat,
                       "native_gb_per_s": n_songs * 750 * 128 * 2 / 1e9 / t_nat}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
