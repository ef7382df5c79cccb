"""Whisper embedder (fad_whisper_forward) against the reference's own dependency: transformers'
WhisperFeatureExtractor + WhisperModel driven as fadtk/model_loader.py:657-669 does (oracle/whisper_oracle.py),
with the seeded synthetic weights of fadtk_b200/weights_whisper.py in both."""
import numpy as np
import pytest
import torch

import fadtk_b200 as fk
from fadtk_b200 import synth, weights_whisper as ww
from oracle import whisper_oracle as wo


def test_packing_and_registry():
    sd, start = ww.load_whisper_state(size="tiny")
    cfg = ww.config_of(sd)
    assert cfg == (384, 6, 4, 4, 1536)
    pk = ww.pack_whisper(sd, start)
    assert len(pk) == 5 + 12 * 4 + 3 + 20 * 4 + 2
    assert pk[0].shape == (2 * 384, 384) and pk[0].dtype == torch.float16           # conv1 as a 3-tap GEMM, mel padded to 128
    w1 = sd["encoder.conv1.weight"]
    assert torch.equal(pk[0][5, 128 + 7].float(), w1[5, 7, 1].to(torch.float16).float())   # hi tile, tap 1, mel bin 7
    assert not pk[0][:128, 80:128].any()
    qkv_b = pk[5 + 3]
    assert qkv_b.shape == (3 * 384,) and not qkv_b[384:768].any()                   # k_proj has no bias
    assert pk[5 + 12 * 4 + 2].shape == (2, 384)                                      # decoder start embedding for both tokens
    names = {m.name: m for m in fk.get_all_models()}
    assert isinstance(names["whisper-small"], fk.WhisperModel) and names["whisper-small"].num_features == 768
    assert names["whisper-large"].num_features == 1280 and names["whisper-tiny"].sr == 16000


def test_oracle_is_the_reference_dependency():
    """The oracle drives transformers exactly like the reference: 2 rows of d_model per clip, fp16."""
    sd, start = ww.load_whisper_state(size="tiny")
    model, fe = wo.build(sd, start)
    e = wo.embed(synth.musiclike_clip(1, 2.0, 16000) / 32768.0, model, fe, start)
    assert e.shape == (2, 384) and e.dtype == np.float16 and np.isfinite(e.astype(np.float32)).all()


def _clips():
    return [synth.musiclike_clip(4, 10.0, 16000), synth.noise_clip(2, 1.7, 16000), synth.musiclike_clip(9, 31.0, 16000)]


def _upload(engine, clips):
    dev = engine.torch_device
    lens = np.array([len(c) for c in clips], dtype=np.int32)
    starts = np.zeros(len(clips), dtype=np.int64)
    starts[1:] = np.cumsum(lens[:-1])
    return (torch.from_numpy(np.concatenate(clips)).to(dev), torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev))


@pytest.mark.gpu
def test_feature_extractor_stage(engine):
    sd, start = ww.load_whisper_state(size="tiny")
    engine.whisper_load(ww.config_of(sd), ww.pack_whisper(sd, start), max_clips=4)
    clips = _clips()                                     # 10 s, 1.7 s and a 31-s clip (truncated to 30 s)
    got = engine.whisper_features(*_upload(engine, clips)).cpu().numpy()
    import transformers as tr
    fe = tr.WhisperFeatureExtractor()
    for i, c in enumerate(clips):
        want = wo.features(c / 32768.0, fe).T           # [3000, 80]
        err = np.abs(got[i] - want)
        assert err.max() < 2e-3 and err.mean() < 2e-5, (i, err.max(), err.mean())


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["tiny", "small"])
def test_embeddings_match_transformers(engine, size):
    clips = _clips()[:2] if size == "small" else _clips()
    ml = fk.WhisperModel(size, max_clips=2)
    ml.load_model()
    got = np.stack(ml.embed_pcm_batch(clips)).astype(np.float32)
    sd, start = ww.load_whisper_state(size=size)
    model, fe = wo.build(sd, start)
    want = np.stack([wo.embed(c / 32768.0, model, fe, start) for c in clips]).astype(np.float32)
    assert got.shape == want.shape == (len(clips), 2, ml.num_features)
    rel = np.sqrt(((got - want) ** 2).mean() / (want ** 2).mean())
    cos = (got * want).sum(-1) / (np.linalg.norm(got, axis=-1) * np.linalg.norm(want, axis=-1))
    print(f"whisper-{size}: rms rel err {rel:.2e}, min cosine {cos.min():.6f}")
    assert rel < 5e-3 and cos.min() > 0.9999, (rel, cos)
    one = ml.get_embedding(clips[1] / 32768.0)           # plugin contract, single clip, batch-invariant
    assert one.dtype == np.float16 and one.shape == (2, ml.num_features)
    assert np.array_equal(one, got[1].astype(np.float16))


@pytest.mark.gpu
def test_whisper_fad_parity_on_identical_audio(engine):
    """FAD (whisper-tiny embeddings, 2 rows per clip) on the same audio: CUDA path vs the reference's CPU path
    (transformers forward + reference statistics / Frechet arithmetic)."""
    from oracle import fad_oracle as fo
    n = 40
    sets = {"base": [synth.noise_clip(i, 4.0 + 0.1 * i, 16000) for i in range(n)],
            "eval": [synth.musiclike_clip(i, 5.0 + 0.1 * i, 16000) for i in range(n)]}
    ml = fk.WhisperModel("tiny", max_clips=8)
    ml.load_model()
    sd, start = ww.load_whisper_state(size="tiny")
    model, fe = wo.build(sd, start)
    gpu = {k: np.concatenate(ml.embed_pcm_batch(v)) for k, v in sets.items()}
    cpu = {k: np.concatenate([wo.embed(c / 32768.0, model, fe, start) for c in v]) for k, v in sets.items()}
    assert gpu["eval"].shape == cpu["eval"].shape == (2 * n, 384)
    fad_gpu = fk.calc_frechet_distance(*fk.calc_embd_statistics(gpu["base"]), *fk.calc_embd_statistics(gpu["eval"]))
    fad_cpu = fo.frechet_distance(*fo.embd_statistics(cpu["base"]), *fo.embd_statistics(cpu["eval"]))
    rel = abs(fad_gpu - fad_cpu) / abs(fad_cpu)
    print(f"whisper-tiny FAD gpu {fad_gpu:.6f} cpu reference path {fad_cpu:.6f} rel {rel:.2e}")
    assert rel < 1e-4, (fad_gpu, fad_cpu, rel)
