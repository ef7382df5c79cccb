"""Directory flow (cache_embedding_files) against the chunk size of the batch driver: how much of the file reading
overlaps the forward.  One process, model loaded once, a fresh directory of 2000 x 10 s .wav files per setting."""
import json
import shutil
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from benchmarks.file_flow import write_set  # noqa: E402
from fadtk_b200 import fad_batch  # noqa: E402
from fadtk_b200.model_loader import VGGishModel  # noqa: E402

root = Path("/tmp/fadtk_chunk_sweep")
shutil.rmtree(root, ignore_errors=True)
ml = VGGishModel()
ml.load_model()
write_set(root / "warm", 64, 3)
fad_batch.cache_embedding_files(root / "warm", ml, workers=16)
out = []
for rep in range(2):
    for mega in (256, 64, 32, 16):
        d = root / f"eval_{mega}_{rep}"
        write_set(d, 2000, 1 + rep)
        fad_batch._CHUNK_SAMPLES = mega * 1024 * 1024
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fad_batch.cache_embedding_files(d, ml, workers=16)
        dt = time.perf_counter() - t0
        out.append({"chunk_msamples": mega, "rep": rep, "seconds": round(dt, 4), "audio_s_per_s": round(20000 / dt)})
        shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
shutil.rmtree(root, ignore_errors=True)
