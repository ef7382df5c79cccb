"""CPU oracle for the wav2vec 2.0 / HuBERT / MERT embedders (test infrastructure only).

The reference's loaders (fadtk/model_loader.py:262-288, 540-560, 578-596) are transformers calls: the
processor (Wav2Vec2FeatureExtractor: zero-mean unit-variance per clip) and Wav2Vec2Model / HubertModel with
``output_hidden_states=True``, then ``hidden_states[layer]``.  transformers is installed here, so the oracle is
the reference's own dependency driven the same way, with the synthetic weights of fadtk_b200/weights_w2v.py.
MERT-v1-95M ships remote code that is not available offline; it is a HuBERT-base architecture at 24 kHz
(``feature_extractor_cqt`` off, ``conv_pos_batch_norm`` forced off by the reference, :263) and is pinned through
HubertModel - parity against the real remote code and against any pretrained weights is unpinned.
"""
from __future__ import annotations

import numpy as np
import torch


def build(sd: dict, family: str = "w2v2", sr: int = 16000):
    import transformers as tr
    d = sd["feature_projection.projection.weight"].shape[0]
    layers = len({k.split(".")[2] for k in sd if k.startswith("encoder.layers.")})
    kw = dict(hidden_size=d, num_hidden_layers=layers, num_attention_heads=d // 64,
              intermediate_size=sd["encoder.layers.0.feed_forward.intermediate_dense.weight"].shape[0],
              hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0, layerdrop=0.0,
              mask_time_prob=0.0, mask_feature_prob=0.0, apply_spec_augment=False)
    if "feature_extractor.conv_layers.1.layer_norm.weight" in sd:      # hubert-large-ls960 style
        kw.update(feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True)
    if family == "wavlm":
        model = tr.WavLMModel(tr.WavLMConfig(**kw)).eval()
    else:
        model = (tr.Wav2Vec2Model(tr.Wav2Vec2Config(**kw)) if family == "w2v2" else tr.HubertModel(tr.HubertConfig(**kw))).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {"masked_spec_embed"}, (missing, unexpected)
    return model, tr.Wav2Vec2FeatureExtractor(sampling_rate=sr, do_normalize=True, return_attention_mask=False)


@torch.no_grad()
def embed(wave: np.ndarray, model, fe, layer: int, sr: int = 16000) -> np.ndarray:
    """-> fp16 [frames, d_model] (ModelLoader.get_embedding's fp32 -> fp16)."""
    inputs = fe(np.asarray(wave, dtype=np.float64), sampling_rate=sr, return_tensors="pt")
    out = model(**inputs, output_hidden_states=True)
    return torch.stack(out.hidden_states).squeeze(1)[layer].numpy().astype(np.float16)
