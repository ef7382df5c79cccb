"""Pin oracle/clap_oracle.py (network part) to an independent implementation of the same published
architecture: transformers' ClapAudioModelWithProjection (HTSAT-tiny), with shared random weights.
CPU only.  The front-end is checked against torch.stft-free closed forms and librosa-style
properties (the reference's torchlibrosa is not installable offline)."""
import numpy as np
import pytest
import torch

from oracle import clap_oracle as co


def _hf_model(sd):
    tr = pytest.importorskip("transformers")
    embed, depths = co.config_of(sd)
    if embed == 96:
        cfg = tr.ClapAudioConfig()           # defaults == HTSAT-tiny (depths 2,2,6,2; heads 4,8,16,32)
    else:                                    # HTSAT-base (clap-laion-music): embed 128, depths 2,2,12,2, final width 1024
        cfg = tr.ClapAudioConfig(patch_embeds_hidden_size=embed, depths=list(depths), hidden_size=8 * embed)
    assert cfg.patch_embeds_hidden_size == embed and list(cfg.depths) == list(depths)
    cfg.hidden_dropout_prob = 0.0
    model = tr.ClapAudioModelWithProjection(cfg).eval()
    hf = model.state_dict()
    mapped = {}
    for k, v in sd.items():
        name = k if k.startswith("audio_projection.") else "audio_model.audio_encoder." + k
        assert name in hf, name
        assert hf[name].shape == v.shape, (name, hf[name].shape, v.shape)
        mapped[name] = v
    missing = [k for k in hf if k not in mapped and "relative_position_index" not in k and "num_batches_tracked" not in k]
    assert not missing, missing
    model.load_state_dict(mapped, strict=False)
    return model


@pytest.mark.parametrize("variant", ["tiny", "base"])
def test_network_matches_independent_hf_port(variant):
    sd = co.synthetic_state(1, variant)
    model = _hf_model(sd)
    g = torch.Generator().manual_seed(0)
    lm = -30.0 + 12.0 * torch.randn((2, co.FRAMES, co.N_MEL), generator=g)
    with torch.no_grad():
        want = model(input_features=lm[:, None], is_longer=torch.zeros(2, 1, dtype=torch.bool)).audio_embeds
        want = torch.nn.functional.normalize(want, dim=-1)       # laion_clap normalises, the HF head does not
    got = co.network(lm, sd)
    assert got.shape == (2, 512)
    assert torch.allclose(got, want, atol=2e-5), (got - want).abs().max()


def test_quantisation_and_chunking_follow_reference_call_site():
    # model_loader.py:413-418: truncation towards zero of x*32767, then /32767
    x = np.array([0.0, 1.0 / 32768, 100 / 32768, -100 / 32768, 0.999, -1.0, 1.5])
    q = co.quantize_like_reference(x)
    assert q.dtype == np.float32
    assert np.array_equal(np.round(q * 32767).astype(int), [0, 0, 99, -99, 32734, -32767, 32767])
    # model_loader.py:396-404: a 10-s clip gives 10 windows, the last one 1 s of audio + 9 s of zeros
    wave = np.arange(10 * co.SR, dtype=np.float32) / (10 * co.SR)
    ch = co.chunks_of(wave)
    assert ch.shape == (10, co.CHUNK)
    assert np.array_equal(ch[9, :co.SR], wave[9 * co.SR:]) and not ch[9, co.SR:].any()
    assert co.chunks_of(wave[: co.SR // 2]).shape == (1, co.CHUNK)


def test_front_end_shapes_and_mel_filterbank_properties():
    fb = co.mel_filterbank()
    assert fb.shape == (64, 513) and (fb >= 0).all()
    freqs = np.linspace(0, co.SR / 2, 513)
    centres = (fb * freqs).sum(1) / fb.sum(1)
    assert (np.diff(centres) > 0).all() and 50 < centres[0] < 150 and 12500 < centres[-1] < 14000
    # Slaney normalisation: each filter has unit area in Hz -> a white spectrum maps to a flat mel spectrum
    assert np.allclose(fb.sum(1) * (co.SR / 2 / 512), 1.0, rtol=0.08)
    tone = torch.sin(2 * np.pi * 1000.0 * torch.arange(co.CHUNK) / co.SR)[None] * 0.5
    lm = co.log_mel(tone)
    assert lm.shape == (1, co.FRAMES, 64)
    peak_bin = lm[0, 500].argmax().item()
    assert abs(centres[peak_bin] - 1000.0) < 80
    # power of a 0.5-amplitude sine through a periodic Hann window of 1024: (0.5 * 512 / 2)^2
    assert lm[0, 500].max().item() == pytest.approx(10 * np.log10((0.5 * 256) ** 2 * fb[peak_bin].max()), abs=1.5)


def test_embed_returns_reference_shape_and_dtype():
    sd = co.synthetic_state(0)
    rng = np.random.default_rng(0)
    wave = (rng.normal(0, 0.1, co.SR * 2)).astype(np.float32)     # 2 s -> 2 windows
    e = co.embed(wave, sd)
    assert e.shape == (2, 512) and e.dtype == np.float16
    assert np.allclose(np.linalg.norm(e.astype(np.float32), axis=1), 1.0, atol=2e-3)
