"""Where the directory flow (`fadtk vggish <dir> <dir>`) spends its host time: cProfile of cache_embedding_files +
FrechetAudioDistance.score on synthetic .wav directories (run on the GPU box).  Prints the top cumulative entries."""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np          # noqa: E402
import torch                # noqa: E402
from fadtk_b200 import _io_native, synth                      # noqa: E402
from fadtk_b200.fad import FrechetAudioDistance               # noqa: E402
from fadtk_b200.fad_batch import cache_embedding_files        # noqa: E402
from fadtk_b200.model_loader import VGGishModel               # noqa: E402


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    root = Path(tempfile.mkdtemp(prefix="fadtk_prof_"))
    dev = torch.device("cuda", 0)

    def write_set(sub, count, seed):
        (root / sub).mkdir(parents=True)
        pcm = synth.musiclike_device(count, 10.0, 16000, seed, dev).cpu().numpy()
        paths = [root / sub / f"clip{i:06d}.wav" for i in range(count)]
        assert not _io_native.wav_write(paths, pcm.reshape(-1), np.arange(count) * pcm.shape[1], np.full(count, pcm.shape[1]), 16000, 16).any()
    write_set("warm", 64, 3)
    write_set("base", 500, 2)
    write_set("eval", clips, 1)
    ml = VGGishModel()
    ml.load_model()
    cache_embedding_files(root / "warm", ml, workers=16)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    cache_embedding_files(root / "base", ml, workers=16)
    cache_embedding_files(root / "eval", ml, workers=16)
    t1 = time.perf_counter()
    score = FrechetAudioDistance(ml, audio_load_worker=16, load_model=False).score(root / "base", root / "eval")
    pr.disable()
    t2 = time.perf_counter()
    print(f"files {clips + 500}: embed {t1 - t0:.3f}s score {t2 - t1:.3f}s -> {(clips + 500) * 10 / (t2 - t0):.0f} audio-s/s, fad {score:.4f}")
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45)
    print(out.getvalue()[:9000])


if __name__ == "__main__":
    main()
