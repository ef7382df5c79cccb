"""Property tests of the HOST-ONLY planning entry points of the C ABI (no GPU): window / example / frame plans against the
reference's own slicing rules restated in oracle/, on random ragged clip lengths (hypothesis)."""
import math

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from fadtk_b200 import _native
from oracle import clap_oracle as co, vggish_oracle as vo

lengths = st.lists(st.integers(min_value=0, max_value=1_300_000), min_size=1, max_size=12)


def _offsets(lens):
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


@settings(max_examples=80, deadline=None)
@given(lengths)
def test_vggish_plan_is_the_reference_framing(lens):
    """torchvggish frames 0.96-s examples with a 0.96-s hop after a 25 ms / 10 ms STFT (oracle/vggish_oracle.num_examples):
    a clip's examples start every 96 * 160 samples from its first sample; clips do not leak into each other."""
    off = _offsets(lens)
    ex, rows = _native.Engine.vggish_plan(off)
    assert list(rows) == [vo.num_examples(n) for n in lens]
    want = [off[i] + 96 * 160 * np.arange(r) for i, r in enumerate(rows)]
    assert np.array_equal(ex, np.concatenate(want) if len(want) else np.zeros(0, np.int64))
    for i, r in enumerate(rows):                                # every example lies inside its own clip
        if r:
            assert off[i] + 96 * 160 * (r - 1) + 15600 <= off[i + 1]


@settings(max_examples=60, deadline=None)
@given(lengths)
def test_clap_plan_is_one_window_per_started_second(lens):
    """fadtk/model_loader.py:396-404: a 10-s window every second of audio, the tail zero padded (oracle/clap_oracle.chunks_of)."""
    off = _offsets(lens)
    start, valid, rows = _native.Engine.clap_plan(off)
    assert list(rows) == [math.ceil(n / co.SR) for n in lens]
    k = 0
    for i, n in enumerate(lens):
        for w in range(int(rows[i])):
            assert start[k] == off[i] + w * co.SR and valid[k] == min(co.CHUNK, n - w * co.SR)
            k += 1
    assert k == len(start)
    small = [n for n in lens if 0 < n <= 3 * co.SR][:2]         # cross-check the count against the oracle's own slicing
    for n in small:
        assert co.chunks_of(np.zeros(n, np.float32)).shape == (math.ceil(n / co.SR), co.CHUNK)


@settings(max_examples=60, deadline=None)
@given(lengths)
def test_clap_frame_pool_indexes_every_window_frame(lens):
    """The frame pool computes each distinct STFT frame once; every window's 1001 frames must index a pool entry whose
    position in the clip is window start + frame * hop (frames of the zero padding share one entry per distinct position)."""
    lens = [n for n in lens if n > 0] or [1]
    off = _offsets(lens)
    plan = _native.Engine.clap_plan_frames(off)
    start, valid, rows = _native.Engine.clap_plan(off)
    fi = plan["frame_index"]
    assert fi.shape == (len(start), 1001) and fi.min() >= 0 and fi.max() < len(plan["pool_start"])
    assert list(plan["rows_per_clip"]) == list(rows)
    assert len(plan["pool_start"]) <= fi.size                   # sharing never creates more work than the naive plan


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=400, max_value=2_000_000))
def test_w2v_frame_count_is_the_conv_stack_formula(n):
    """Seven valid convolutions (k, s) = (10,5) (3,2)x4 (2,2)x2: transformers' _get_feat_extract_output_lengths."""
    t = n
    for k, s in ((10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2)):
        t = torch.div(torch.tensor(t - k), s, rounding_mode="floor").item() + 1
    assert _native.Engine.w2v_frames(n) == t


@settings(max_examples=100, deadline=None)
@given(st.sampled_from([8000, 16000, 22050, 24000, 32000, 44100, 48000]), st.sampled_from([16000, 24000, 48000]),
       st.integers(min_value=1, max_value=5_000_000))
def test_resample_length_is_torchaudio_rule(sr_in, sr_out, n):
    """torchaudio.functional.resample: target_length = ceil(new * length / orig) with the rates reduced by their gcd."""
    g = math.gcd(sr_in, sr_out)
    orig, new = sr_in // g, sr_out // g
    assert int(_native.lib().fad_resample_length(sr_in, sr_out, n)) == -(-new * n // orig)
    if sr_in != sr_out:
        o, nw, width, taps = _native.Engine.resample_geometry(sr_in, sr_out)
        assert (o, nw) == (orig, new) and taps == 2 * width + orig
