"""CPU oracle for the Whisper embedder (test infrastructure only - never imported by the product path).

The reference's WhisperModel loader (fadtk/model_loader.py:636-672) is three calls into the
``transformers`` package: ``AutoFeatureExtractor`` (WhisperFeatureExtractor), ``WhisperModel`` and its
``last_hidden_state`` for ``decoder_input_ids = [[sot, sot]]``.  transformers is installed in this image
(5.5.0; the reference pins 4.52.3 - same Whisper architecture), so the oracle IS the reference's
dependency, driven exactly as the reference drives it, with the synthetic weights of
fadtk_b200/weights_whisper.py loaded into it.  No pretrained checkpoint exists offline: parity against
real openai/whisper weights is unpinned.
"""
from __future__ import annotations

import numpy as np
import torch


def build(sd: dict, decoder_start_token_id: int):
    import transformers as tr
    d = sd["encoder.conv1.weight"].shape[0]
    n_enc = len({k.split(".")[2] for k in sd if k.startswith("encoder.layers.")})
    n_dec = len({k.split(".")[2] for k in sd if k.startswith("decoder.layers.")})
    cfg = tr.WhisperConfig(d_model=d, encoder_layers=n_enc, decoder_layers=n_dec, encoder_attention_heads=d // 64,
                           decoder_attention_heads=d // 64, encoder_ffn_dim=sd["encoder.layers.0.fc1.weight"].shape[0],
                           decoder_ffn_dim=sd["decoder.layers.0.fc1.weight"].shape[0], num_mel_bins=80,
                           vocab_size=sd["decoder.embed_tokens.weight"].shape[0], pad_token_id=0, bos_token_id=1, eos_token_id=2,
                           decoder_start_token_id=decoder_start_token_id, suppress_tokens=None, begin_suppress_tokens=None,
                           dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    model = tr.WhisperModel(cfg).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "proj_out" not in m], (missing, unexpected)
    return model, tr.WhisperFeatureExtractor()


@torch.no_grad()
def embed(wave: np.ndarray, model, fe, decoder_start_token_id: int) -> np.ndarray:
    """model_loader.py:663-669 -> fp16 [2, d_model] (ModelLoader.get_embedding's fp32 -> fp16, :47-48)."""
    feats = fe(np.asarray(wave, dtype=np.float64), sampling_rate=16000, return_tensors="pt").input_features
    ids = torch.tensor([[1, 1]]) * decoder_start_token_id
    out = model(feats, decoder_input_ids=ids).last_hidden_state.squeeze()
    return out.numpy().astype(np.float16)


def features(wave: np.ndarray, fe) -> np.ndarray:
    """input_features [80, 3000] float32 of one clip."""
    return fe(np.asarray(wave, dtype=np.float64), sampling_rate=16000, return_tensors="np").input_features[0]
