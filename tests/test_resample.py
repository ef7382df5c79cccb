"""GPU resampler (fad_resample) against torchaudio's Resample with the reference's parameters
(fadtk/fad.py:151-158).  torchaudio is the checker here (library oracle), never the product path."""
import math

import numpy as np
import pytest
import torch

from fadtk_b200 import _native, synth

REF_KW = dict(lowpass_filter_width=64, rolloff=0.9475937167399596, resampling_method="sinc_interp_kaiser",
              beta=14.769656459379492)
RATES = [(44100, 16000), (48000, 16000), (16000, 48000), (22050, 48000), (32000, 24000), (44100, 48000)]


@pytest.mark.parametrize("sr_in,sr_out", RATES)
def test_filter_bank_is_bit_identical_to_torchaudio(sr_in, sr_out):
    from torchaudio.functional import functional as F
    g = math.gcd(sr_in, sr_out)
    want, width = F._get_sinc_resample_kernel(sr_in, sr_out, g, REF_KW["lowpass_filter_width"], REF_KW["rolloff"],
                                              REF_KW["resampling_method"], REF_KW["beta"])
    orig, new, w, taps = _native.Engine.resample_geometry(sr_in, sr_out)
    assert (orig, new, w, taps) == (sr_in // g, sr_out // g, width, 2 * width + sr_in // g)
    got = _native.Engine.resample_bank(sr_in, sr_out)
    assert got.shape == (new, taps) and got.dtype == np.float32
    assert np.array_equal(got, want[:, 0, :].numpy())            # float32, bit for bit
    for length in (0, 1, 441, 44100, 123457):
        assert _native.lib().fad_resample_length(sr_in, sr_out, length) == math.ceil(new * length / orig)


def _reference_convert(x_float: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
    import torchaudio
    x = torch.mean(x_float, 0).unsqueeze(0)                       # fad.py:150
    if sr_in != sr_out:
        x = torchaudio.transforms.Resample(sr_in, sr_out, **REF_KW)(x)
    return x[0]


@pytest.mark.gpu
@pytest.mark.parametrize("sr_in,sr_out,channels", [(44100, 16000, 2), (48000, 16000, 1), (16000, 48000, 2),
                                                   (22050, 48000, 1), (16000, 16000, 2)])
def test_resample_matches_torchaudio(engine, sr_in, sr_out, channels):
    rng = np.random.default_rng(sr_in + channels)
    n = int(2.3 * sr_in) + 17
    t = np.arange(n) / sr_in
    pcm = np.stack([np.round(32767 * np.clip(0.4 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t) + 0.1 * rng.standard_normal(n), -1, 1))
                    for c in range(channels)], 1).astype(np.int16)
    want = _reference_convert(torch.from_numpy(pcm.T.astype(np.float32) / 32768.0), sr_in, sr_out)
    dev = engine.torch_device
    got_pcm, got_f = engine.resample(torch.from_numpy(pcm).to(dev), sr_in, sr_out, return_float=True)
    assert got_f.shape == want.shape
    err = (got_f.cpu() - want).abs().max().item()
    assert err < 2e-6, err                                         # fp32 summation-order noise only
    want_pcm = torch.clamp(torch.round(want * 32768.0), -32768, 32767).to(torch.int16)
    diff = (got_pcm.cpu().int() - want_pcm.int()).abs()
    assert diff.max().item() <= 1 and (diff != 0).float().mean().item() < 0.02
    # planar float input (decoded non-WAV formats) takes the same path
    got2 = engine.resample(torch.from_numpy(pcm.T.astype(np.float32) / 32768.0).to(dev), sr_in, sr_out)
    assert torch.equal(got2, got_pcm)


@pytest.mark.gpu
def test_convert_audio_resamples_on_the_gpu(engine, tmp_path):
    """fad.py:139-160 flow: a 44.1 kHz stereo WAV lands as mono PCM16 at the model rate under convert/<sr>/."""
    import fadtk_b200 as fk

    class Loader(fk.ModelLoader):
        def __init__(self):
            super().__init__("gold", 128, 16000)

        def load_model(self):
            pass

        def _get_embedding(self, audio):
            raise NotImplementedError

    n = 44100
    t = np.arange(n) / 44100
    pcm = np.stack([np.round(12000 * np.sin(2 * np.pi * 330 * t)), np.round(9000 * np.sin(2 * np.pi * 550 * t))], 1).astype(np.int16)
    import wave
    with wave.open(str(tmp_path / "a.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    fad = fk.FrechetAudioDistance(Loader(), audio_load_worker=1, load_model=False)
    wav = fad.load_audio(tmp_path / "a.wav")
    out, sr = synth.read_wav(tmp_path / "convert" / "16000" / "a.wav")
    assert sr == 16000 and out.ndim == 1 and len(out) == 16000 and len(wav) == 16000
    want = _reference_convert(torch.from_numpy(pcm.T.astype(np.float32) / 32768.0), 44100, 16000)
    assert np.abs(out.astype(np.float64) / 32768.0 - want.numpy()).max() < 1.0 / 32768.0 + 1e-6
