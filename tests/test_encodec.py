"""Encodec-24 kHz encoder (fad_encodec_forward) against oracle/encodec_oracle.py, which
tests/test_encodec_oracle.py pins to transformers' independent port."""
import numpy as np
import pytest

import fadtk_b200 as fk
from fadtk_b200 import synth, weights_encodec as we
from oracle import encodec_oracle as eo


def test_registry_entry():
    names = {m.name: m for m in fk.get_all_models()}
    ml = names["encodec-emb"]
    assert isinstance(ml, fk.EncodecEmbModel) and ml.num_features == 128 and ml.sr == 24000
    assert names["encodec-emb-48k"].sr == 48000 and isinstance(names["encodec-emb-48k"], fk.EncodecEmbModel)


@pytest.mark.gpu
def test_embeddings_match_oracle(engine):
    clips = [synth.musiclike_clip(3, 2.0, 24000), synth.musiclike_clip(5, 2.0, 24000), synth.noise_clip(1, 0.73, 24000)]
    ml = fk.EncodecEmbModel('24k', max_chunk_samples=4 * 48000)
    ml.load_model()
    got = ml.embed_pcm_batch(clips)
    sd = we.synthetic_encodec_state(0)
    for g, c in zip(got, clips):
        want = eo.embed(c / 32768.0, sd).astype(np.float32)
        g = g.astype(np.float32)
        assert g.shape == want.shape == (-(-len(c) // 320), 128)
        rel = np.sqrt(((g - want) ** 2).mean() / (want ** 2).mean())
        print(f"encodec {len(c)} samples: rms rel err {rel:.2e}, max abs {np.abs(g - want).max():.3e} (scale {np.abs(want).max():.2f})")
        assert rel < 3e-3, rel
    one = ml.get_embedding(clips[2] / 32768.0)
    assert one.dtype == np.float16 and np.array_equal(one, got[2])


@pytest.mark.gpu
def test_encodec_fad_parity_on_identical_audio(engine):
    from oracle import fad_oracle as fo
    n = 6
    sets = {"base": [synth.noise_clip(i, 3.0, 24000) for i in range(n)],
            "eval": [synth.musiclike_clip(i, 3.0, 24000) for i in range(n)]}
    ml = fk.EncodecEmbModel('24k', max_chunk_samples=6 * 72000)
    ml.load_model()
    sd = we.synthetic_encodec_state(0)
    gpu = {k: np.concatenate(ml.embed_pcm_batch(v)) for k, v in sets.items()}
    cpu = {k: np.concatenate([eo.embed(c / 32768.0, sd) for c in v]) for k, v in sets.items()}
    assert gpu["eval"].shape == cpu["eval"].shape == (n * 225, 128)
    fad_gpu = fk.calc_frechet_distance(*fk.calc_embd_statistics(gpu["base"]), *fk.calc_embd_statistics(gpu["eval"]))
    fad_cpu = fo.frechet_distance(*fo.embd_statistics(cpu["base"]), *fo.embd_statistics(cpu["eval"]))
    rel = abs(fad_gpu - fad_cpu) / abs(fad_cpu)
    print(f"encodec FAD gpu {fad_gpu:.6f} cpu reference path {fad_cpu:.6f} rel {rel:.2e}")
    assert rel < 1e-4, (fad_gpu, fad_cpu, rel)


@pytest.mark.gpu
def test_48k_variant_matches_oracle(engine):
    """encodec-emb-48k: non-causal GroupNorm encoder on 1-s segments of the duplicated-mono signal."""
    clips = [synth.musiclike_clip(3, 2.4, 48000), synth.noise_clip(1, 2.4, 48000)]
    ml = fk.EncodecEmbModel('48k', max_chunk_samples=8 * 48000)
    ml.load_model()
    got = ml.embed_pcm_batch(clips)
    sd = we.synthetic_encodec_state(0, "48k")
    for g, c in zip(got, clips):
        want = eo.embed(c / 32768.0, sd).astype(np.float32)
        g = g.astype(np.float32)
        assert g.shape == want.shape == (150 + 150 + 60, 128)
        rel = np.sqrt(((g - want) ** 2).mean() / (want ** 2).mean())
        print(f"encodec-48k: rms rel err {rel:.2e}")
        assert rel < 3e-3, rel
