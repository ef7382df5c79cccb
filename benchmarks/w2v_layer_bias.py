"""Where does the wav2vec forward pick up its systematic FAD offset?  Per tapped layer: rms error and the error of the
per-dimension MEAN over all rows (what moves |mu1 - mu2|^2 and does not average out), both relative to the rms of the
reference hidden state.  16 four-second clips; reference = transformers fp32 on the CPU (the oracle)."""
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("FADTK_SYNTHETIC", "1")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fadtk_b200 import _native, synth, weights_w2v as ww  # noqa: E402
from oracle import w2v_oracle as wo  # noqa: E402

eng = _native.engine(0)
sd = ww.synthetic_w2v_state(0)
eng.w2v_load(ww.config_of(sd), ww.pack_w2v(sd), 8, max_len=16000 * 5)
model, fe = wo.build(sd, "w2v2")
clips = [synth.musiclike_clip(i, 4.0, 16000) for i in range(8)] + [synth.noise_clip(i, 4.0, 16000) for i in range(8)]
pcm = torch.from_numpy(np.stack(clips)).to(eng.torch_device)
out = []
for layer in (0, 1, 2, 4, 8, 12):
    got = torch.cat([eng.w2v_forward(pcm[s:s + 8], layer) for s in (0, 8)]).float().cpu().numpy().reshape(-1, 768).astype(np.float64)
    want = np.concatenate([wo.embed(c / 32768.0, model, fe, layer) for c in clips]).astype(np.float64).reshape(-1, 768)
    want32 = np.concatenate([wo.embed(c / 32768.0, model, fe, layer).astype(np.float16) for c in clips]).astype(np.float64).reshape(-1, 768)
    rms = np.sqrt((want ** 2).mean())
    err = got - want
    out.append({"layer": layer, "rms_rel": float(np.sqrt((err ** 2).mean()) / rms),
                "mean_err_rel": float(err.mean() / rms),
                "dim_mean_err_rms_rel": float(np.sqrt((err.mean(0) ** 2).mean()) / rms),
                "expected_if_random": float(np.sqrt((err ** 2).mean()) / rms / np.sqrt(err.shape[0])),
                "fp16_rounding_of_reference_rms_rel": float(np.sqrt(((want32 - want) ** 2).mean()) / rms)})
print(json.dumps(out))
