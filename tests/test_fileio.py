"""libfadtk_io.so (include/fadtk_b200_io.h): batched WAV / .npy I/O against Python's wave module and numpy -
the formats of the reference's convert cache (fad.py:160), load_wav (model_loader.py:63-70) and embedding
cache (fad.py:200-209).  Host only: runs without a GPU."""
import io
import re
import struct
import wave

import torch
from pathlib import Path

import numpy as np
import pytest

from fadtk_b200 import _io_native as ion

ROOT = Path(__file__).resolve().parents[1]


def _write_wave(path, pcm, sr, channels=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "fadtk_b200_io.h").read_text()
    declared = set(re.findall(r"\b(fad_io_\w+)\s*\(", header))
    assert declared == set(ion.SIGNATURES)
    lib = ion.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fad_io_version() == 1


def test_wav_probe_and_read_match_the_wave_module(tmp_path):
    rng = np.random.default_rng(0)
    specs = [(16000, 1, 16000), (48000, 2, 12345), (24000, 1, 1), (44100, 3, 777), (16000, 1, 0)]
    paths, want = [], []
    for i, (sr, ch, n) in enumerate(specs):
        pcm = rng.integers(-32768, 32768, size=n * ch, dtype=np.int16)
        p = tmp_path / f"clip {i} ü.wav"                      # spaces / non-ASCII in the path
        _write_wave(p, pcm, sr, ch)
        paths.append(p)
        want.append(pcm)
    sr, ch, fr, st = ion.wav_probe(paths, threads=3)
    assert st.tolist() == [0] * len(specs)
    assert list(zip(sr.tolist(), ch.tolist(), fr.tolist())) == specs
    out = np.full(int((fr * ch).sum()) + 5, 99, dtype=np.int16)
    off, st = ion.wav_read(paths, fr, ch, out, threads=2)
    assert st.tolist() == [0] * len(specs)
    for i, pcm in enumerate(want):
        np.testing.assert_array_equal(out[off[i]:off[i + 1]], pcm)
    assert (out[off[-1]:] == 99).all()                        # nothing written past the last clip


def test_wav_chunks_extensible_header_and_errors(tmp_path):
    pcm = np.arange(-50, 50, dtype=np.int16)
    # LIST chunk before fmt, WAVE_FORMAT_EXTENSIBLE fmt, odd-sized chunk (padded), data size 0xFFFFFFFF (streamed writer)
    fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, 1, 22050, 44100, 2, 16, 22, 16, 4, 1, b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    body = b"WAVE" + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + b"fmt " + struct.pack("<I", len(fmt)) + fmt \
        + b"data" + struct.pack("<I", 0xFFFFFFFF) + pcm.tobytes()
    ext = tmp_path / "ext.wav"
    ext.write_bytes(b"RIFF" + struct.pack("<I", 4 + len(body)) + body)
    f32 = tmp_path / "float.wav"                              # IEEE float WAV: valid but not PCM16
    fmt3 = struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32)
    body3 = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt3 + b"data" + struct.pack("<I", 8) + b"\x00" * 8
    f32.write_bytes(b"RIFF" + struct.pack("<I", len(body3)) + body3)
    junk = tmp_path / "junk.wav"
    junk.write_bytes(b"not a riff file at all")
    missing = tmp_path / "missing.wav"
    sr, ch, fr, st = ion.wav_probe([ext, f32, junk, missing])
    assert st.tolist() == [ion.OK, ion.EUNSUPPORTED, ion.EFORMAT, ion.EOPEN]
    assert (sr[0], ch[0], fr[0]) == (22050, 1, 100)
    out = np.zeros(100, dtype=np.int16)
    _, st = ion.wav_read([ext], fr[:1], ch[:1], out)
    assert st.tolist() == [0]
    np.testing.assert_array_equal(out, pcm)
    # a file that shrank between probe and read is reported, not read past
    _write_wave(ext, pcm[:10], 22050)
    _, st = ion.wav_read([ext], fr[:1], ch[:1], out)
    assert st.tolist() == [ion.ESHORT]


def test_wav_write_is_readable_by_the_wave_module(tmp_path):
    rng = np.random.default_rng(1)
    src = rng.integers(-32768, 32768, size=5000, dtype=np.int16)
    offsets, frames = np.array([0, 1000, 1000, 4000]), np.array([1000, 0, 3000, 1000])
    paths = [tmp_path / f"o{i}.wav" for i in range(4)]
    st = ion.wav_write(paths, src, offsets, frames, 24000, threads=4)
    assert st.tolist() == [0, 0, 0, 0]
    for p, o, n in zip(paths, offsets, frames):
        with wave.open(str(p), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, n)
            np.testing.assert_array_equal(np.frombuffer(w.readframes(n), dtype="<i2"), src[o:o + n])
    st = ion.wav_write([tmp_path / "no_such_dir" / "x.wav"], src, [0], [10], 24000)
    assert st.tolist() == [ion.EOPEN]


def test_npy_files_are_byte_identical_to_numpy_save(tmp_path):
    rng = np.random.default_rng(2)
    emb = rng.standard_normal((1500, 128)).astype(np.float16)
    rows = np.array([10, 0, 750, 1, 739])
    off = np.concatenate([[0], np.cumsum(rows)])
    paths = [tmp_path / f"e{i}.npy" for i in range(len(rows))]
    st = ion.npy_write_f16(paths, emb, off[:-1], rows, threads=2)
    assert st.tolist() == [0] * len(rows)
    for p, o, n in zip(paths, off, rows):
        buf = io.BytesIO()
        np.save(buf, emb[o:o + n])
        assert p.read_bytes() == buf.getvalue()
    # header widths around the 64-byte alignment boundary
    wide = rng.standard_normal((3, 1280)).astype(np.float16)
    for r, d in ((3, 1280), (1, 1), (99999, 2)):
        a = np.resize(wide, (r, d)).astype(np.float16)
        p = tmp_path / f"w{r}_{d}.npy"
        assert ion.npy_write_f16([p], a, [0], [r]).tolist() == [0]
        buf = io.BytesIO()
        np.save(buf, a)
        assert p.read_bytes() == buf.getvalue()


def test_npy_probe_and_ragged_read(tmp_path):
    rng = np.random.default_rng(3)
    arrays = [rng.standard_normal((n, 64)).astype(np.float16) for n in (5, 1, 300, 0, 17)]
    paths = []
    for i, a in enumerate(arrays):
        p = tmp_path / f"a{i}.npy"
        np.save(p, a)
        paths.append(p)
    odd = {"f32": rng.standard_normal((4, 64)).astype(np.float32), "vec": rng.standard_normal(64).astype(np.float16),
           "fortran": np.asfortranarray(rng.standard_normal((4, 64)).astype(np.float16)), "int": np.arange(6).reshape(2, 3)}
    for k, a in odd.items():
        np.save(tmp_path / f"{k}.npy", a)
    rows, cols, nd, dt, st = ion.npy_probe(paths + [tmp_path / f"{k}.npy" for k in odd] + [tmp_path / "nope.npy"])
    assert st.tolist() == [0] * 5 + [ion.OK, ion.OK, ion.EUNSUPPORTED, ion.EUNSUPPORTED, ion.EOPEN]
    assert rows[:7].tolist() == [5, 1, 300, 0, 17, 4, 64] and dt[:7].tolist() == [2] * 5 + [4, 2] and nd[:7].tolist() == [2] * 6 + [1]
    out, off = ion.load_embedding_files(paths, threads=3)
    np.testing.assert_array_equal(out, np.concatenate(arrays))
    assert off.tolist() == [0, 5, 6, 306, 306, 323]
    # a float32 cache among fp16 ones: same result as np.load + np.concatenate
    mixed, off = ion.load_embedding_files(paths + [tmp_path / "f32.npy"])
    ref = np.concatenate(arrays + [odd["f32"]])
    assert mixed.dtype == ref.dtype
    np.testing.assert_array_equal(mixed, ref)
    with pytest.raises(ValueError):
        ion.load_embedding_files([paths[0], tmp_path / "vec.npy"])


def test_batch_driver_paths_match_the_reference_layout(tmp_path):
    """fad_batch derives cache paths with string operations; they must be the names utils.get_cache_embedding_path
    (fadtk/utils.py:60-68) and FrechetAudioDistance._converted_path (fad.py:143) produce."""
    from fadtk_b200 import fad_batch
    from fadtk_b200.fad import FrechetAudioDistance
    from fadtk_b200.utils import get_cache_embedding_path

    class _ML:
        name, sr = "clap-laion-audio", 48000

    f = FrechetAudioDistance.__new__(FrechetAudioDistance)
    f.ml = _ML()
    files = [tmp_path / "a.wav", tmp_path / "x y" / "b.c.flac", tmp_path / "noext", Path("rel") / "d.WAV", Path("e.mp3")]
    emb, conv = fad_batch._derived_paths(files, _ML.name, _ML.sr)
    assert [Path(p) for p in emb] == [get_cache_embedding_path(_ML.name, x) for x in files]
    assert [Path(p) for p in conv] == [f._converted_path(x) for x in files]


def test_cache_embedding_files_with_a_stub_model(tmp_path):
    """The whole host flow without a GPU: native batch read -> stub embedder -> convert cache + .npy cache."""
    from fadtk_b200 import fad_batch, synth
    from fadtk_b200.model_loader import ModelLoader

    class Stub(ModelLoader):
        def __init__(self):
            super().__init__("stub", 4, 16000)

        def load_model(self):
            pass

        def _get_embedding(self, audio):
            raise NotImplementedError

        def embed_pcm_batch(self, clips):                     # [n, 4]: n = seconds, features = simple statistics
            return [np.stack([[c[:16000 * (k + 1)].astype(np.float64).mean(), c.min(), c.max(), len(c)] for k in range(max(1, len(c) // 16000))]).astype(np.float16)
                    for c in clips]

    rng = np.random.default_rng(5)
    clips = {f"c{i}.wav": rng.integers(-3000, 3000, size=16000 * (1 + i % 3), dtype=np.int16) for i in range(7)}
    for name, pcm in clips.items():
        synth.write_wav(tmp_path / name, pcm, 16000)
    stereo = rng.integers(-3000, 3000, size=(16000, 2), dtype=np.int16)   # not mono: must NOT take the native fast path
    _write_wave(tmp_path / "stereo.wav", stereo, 16000, 2)
    ml = Stub()
    calls = []

    class _FAD(fad_batch.FrechetAudioDistance):
        def convert_audio(self, f):
            calls.append(Path(f).name)
            return np.zeros(16000, dtype=np.int16)

    orig = fad_batch.FrechetAudioDistance
    fad_batch.FrechetAudioDistance = _FAD
    try:
        monkey_chunk = fad_batch._CHUNK_FILES
        fad_batch._CHUNK_FILES = 3                             # several chunks: exercises the read-ahead and both staging slots
        assert [len(c) for c in fad_batch._plan_chunks(sorted(tmp_path.glob("*.wav")), ml, 2)] == [3, 3, 2]
        fad_batch.cache_embedding_files(tmp_path, ml, workers=3, load_model=False)
        fad_batch._CHUNK_FILES = monkey_chunk
        budget = fad_batch._CHUNK_SAMPLES
        fad_batch._CHUNK_SAMPLES = 40000                       # chunks are also bounded by samples (long files)
        sizes = [sum(len(clips[f.name]) if f.name in clips else 16000 for f in c) for c in fad_batch._plan_chunks(sorted(tmp_path.glob("*.wav")), ml, 2)]
        fad_batch._CHUNK_SAMPLES = budget
        assert max(sizes) <= 48000 and sum(sizes) == sum(len(v) for v in clips.values()) + 16000 and len(sizes) >= 4
        assert calls == ["stereo.wav"]
        for name, pcm in clips.items():
            e = np.load(tmp_path / "embeddings" / "stub" / (Path(name).stem + ".npy"))
            np.testing.assert_array_equal(e, ml.embed_pcm_batch([pcm])[0])
            with wave.open(str(tmp_path / "convert" / "16000" / name), "rb") as w:
                np.testing.assert_array_equal(np.frombuffer(w.readframes(w.getnframes()), dtype="<i2"), pcm)
        calls.clear()
        fad_batch.cache_embedding_files(tmp_path, ml, workers=3, load_model=False)       # everything cached: no work
        assert calls == []
        # new file next to a populated convert cache; one embedding removed -> read back from the convert cache
        synth.write_wav(tmp_path / "late.wav", clips["c1.wav"], 16000)
        (tmp_path / "embeddings" / "stub" / "c2.npy").unlink()
        (tmp_path / "c2.wav").unlink()
        fad_batch.cache_embedding_files([tmp_path / "late.wav", tmp_path / "c2.wav"], ml, workers=2, load_model=False)
        np.testing.assert_array_equal(np.load(tmp_path / "embeddings" / "stub" / "c2.npy"), ml.embed_pcm_batch([clips["c2.wav"]])[0])
        assert (tmp_path / "embeddings" / "stub" / "late.npy").exists()
    finally:
        fad_batch.FrechetAudioDistance = orig


def test_per_file_means_equal_numpy_file_by_file():
    """utils.per_file_means reduces equal-length files together; _process_file (fadtk/utils.py:14) calls np.mean
    per file on an fp16 array.  The fp16 results must be bit-identical (they feed the mirrored rounding quirk)."""
    from fadtk_b200.utils import per_file_means
    rng = np.random.default_rng(7)
    rows = [750, 750, 3, 750, 0, 1, 3, 10, 10, 10, 1234]
    arrays = [(rng.standard_normal((r, 96)) * rng.uniform(0.1, 30)).astype(np.float16) for r in rows]
    flat = np.concatenate(arrays)
    off = np.concatenate([[0], np.cumsum(rows)])
    m_in, m64, counts = per_file_means(flat, off)
    kept = [a for a in arrays if len(a)]
    assert counts.tolist() == [len(a) for a in kept] and m_in.dtype == np.float16
    for k, a in enumerate(kept):
        assert np.array_equal(m_in[k], np.mean(a, axis=0))
        np.testing.assert_allclose(m64[k], a.mean(axis=0, dtype=np.float64), rtol=1e-14, atol=1e-300)
    wide = flat.astype(np.float32)
    m_in32, _, _ = per_file_means(wide, off)
    assert m_in32.dtype == np.float32
    for k, a in enumerate(kept):
        assert np.array_equal(m_in32[k], np.mean(a.astype(np.float32), axis=0))


def test_read_wav_float_decodes_every_pcm_and_float_format(tmp_path):
    """synth.read_wav_float against scipy.io.wavfile (which writes / reads these formats) with torchaudio.load's
    normalisation: the convert step (fad.py:147) sees float32 [channels, T] whatever the file stores."""
    from scipy.io import wavfile
    from fadtk_b200 import synth
    rng = np.random.default_rng(11)
    T, sr = 1000, 44100
    cases = {
        "u8": (rng.integers(0, 256, size=(T, 2), dtype=np.uint8), lambda a: (a.astype(np.float32) - 128) / 128),
        "i16": (rng.integers(-32768, 32768, size=(T, 1), dtype=np.int16), lambda a: a.astype(np.float32) / 32768),
        "i32": (rng.integers(-2**31, 2**31, size=(T, 2), dtype=np.int32), lambda a: (a.astype(np.float64) / 2**31).astype(np.float32)),
        "f32": (rng.uniform(-1, 1, size=(T, 3)).astype(np.float32), lambda a: a),
        "f64": (rng.uniform(-1, 1, size=(T, 1)), lambda a: a.astype(np.float32)),
    }
    for name, (a, norm) in cases.items():
        p = tmp_path / f"{name}.wav"
        wavfile.write(p, sr, a)
        x, got_sr = synth.read_wav_float(p)
        assert got_sr == sr and x.dtype == np.float32 and x.shape == (a.shape[1], T)
        np.testing.assert_array_equal(x, norm(a).T)
    # 24-bit PCM, written by hand (scipy cannot write it), WAVE_FORMAT_EXTENSIBLE header
    v = rng.integers(-2**23, 2**23, size=(T, 2), dtype=np.int64)
    b = (v & 0xFFFFFF).astype(np.uint32)
    payload = np.stack([(b >> s) & 255 for s in (0, 8, 16)], axis=-1).astype(np.uint8).tobytes()
    fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, 2, sr, sr * 6, 6, 24, 22, 24, 3, 1, b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
    p24 = tmp_path / "i24.wav"
    p24.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    x, _ = synth.read_wav_float(p24)
    np.testing.assert_array_equal(x, (v.astype(np.float32) / 2**23).T)
    sr_sp, a_sp = wavfile.read(p24)                            # scipy reads 24-bit into the high bytes of int32
    np.testing.assert_array_equal(x, (a_sp.astype(np.float64) / 2**31).astype(np.float32).T)
    with pytest.raises(ValueError):
        synth.read_wav_float(__file__)                         # not a RIFF file


_SHARD_WORKER = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from pathlib import Path
from fadtk_b200 import dist, fad_batch
from fadtk_b200.model_loader import ModelLoader
dist.init_from_env("gloo")

class Stub(ModelLoader):
    def __init__(self):
        super().__init__("stub", 2, 16000)
    def load_model(self):
        pass
    def _get_embedding(self, audio):
        raise NotImplementedError
    def embed_pcm_batch(self, clips):                          # row 0: this rank, clip length
        return [np.array([[dist.rank(), len(c) // 16000]] * 2, dtype=np.float16) for c in clips]

fad_batch.cache_embedding_files(Path(sys.argv[2]), Stub(), workers=2, load_model=False)
sys.stdout.write(f"[rank{dist.rank()}:ok]\n"); sys.stdout.flush()
"""


def test_two_ranks_shard_the_directory_gloo(tmp_path):
    """Under torchrun every rank embeds its contiguous share of the sorted file list (fad_batch.py:43's array_split)
    and all of them meet at the barrier: every file embedded exactly once, by the rank array_split assigns."""
    import os
    import subprocess
    import sys
    from fadtk_b200 import synth
    data = tmp_path / "set"
    data.mkdir()
    for i in range(7):
        synth.write_wav(data / f"c{i}.wav", np.zeros(16000 * (1 + i % 2), dtype=np.int16), 16000)
    script = tmp_path / "worker.py"
    script.write_text(_SHARD_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), str(ROOT), str(data)],
                         capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "rank0:ok" in out.stdout and "rank1:ok" in out.stdout
    owners = [int(np.load(data / "embeddings" / "stub" / f"c{i}.npy")[0, 0]) for i in range(7)]
    assert owners == [0, 0, 0, 0, 1, 1, 1]                     # np.array_split(7 files, 2)
    assert [int(np.load(data / "embeddings" / "stub" / f"c{i}.npy")[0, 1]) for i in range(7)] == [1 + i % 2 for i in range(7)]


def test_flat_pcm_is_zero_copy_for_consecutive_views():
    from fadtk_b200.model_loader import flat_pcm
    buf = np.arange(1000, dtype=np.int16)
    stage = buf[:900]                                          # like fad_batch._host_buffer: a slice of the staging array
    clips = [stage[100:300], stage[300:301], stage[301:800]]
    flat = flat_pcm(clips)
    assert np.shares_memory(flat, buf) and flat.shape == (700,) and flat[0] == 100 and flat[-1] == 799
    gap = flat_pcm([stage[0:10], stage[20:30]])                # not adjacent: a copy with the same contents
    assert not np.shares_memory(gap, buf) and gap.tolist() == list(range(10)) + list(range(20, 30))
    swapped = flat_pcm([stage[10:20], stage[0:10]])
    assert swapped.tolist() == list(range(10, 20)) + list(range(0, 10))
    padded = flat_pcm([np.pad(stage[0:5], (0, 3)), stage[5:10]])
    assert padded.tolist() == [0, 1, 2, 3, 4, 0, 0, 0, 5, 6, 7, 8, 9]
    one = flat_pcm([stage[7:9]])
    assert one.tolist() == [7, 8]
    other = flat_pcm([np.arange(4, dtype=np.int16), np.arange(4, 8, dtype=np.int16)])   # separate allocations
    assert other.tolist() == list(range(8))
    assert flat_pcm([stage[0:4], stage[4:8]]).reshape(2, -1).tolist() == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_c_abi_is_usable_from_plain_c(tmp_path):
    """gcc -std=c99 compiles a C program against include/*.h, links libfadtk_io.so and round-trips WAV / .npy files:
    the boundary really is `extern "C"` with plain pointers and sizes."""
    import os
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    ion.lib()                                                  # make sure the library has been built
    exe = tmp_path / "io_abi_check"
    lib_dir = ROOT / "fadtk_b200" / "csrc"
    build = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", str(ROOT / "tests" / "c" / "io_abi_check.c"),
                            "-o", str(exe), f"-L{lib_dir}", "-lfadtk_io", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    work = tmp_path / "files"
    work.mkdir()
    run = subprocess.run([str(exe), str(work)], capture_output=True, text=True, env=dict(os.environ))
    assert run.returncode == 0 and run.stdout.strip() == "ok v1", (run.returncode, run.stdout, run.stderr)
    a = np.load(work / "b.npy")                                # and numpy reads what the C program wrote
    assert a.dtype == np.float16 and a.shape == (4, 4)
    assert a.view(np.uint16).ravel().tolist() == [0x3c00 + i for i in range(8, 24)]


def test_embedding_chunks_are_bounded_by_bytes(tmp_path):
    paths = []
    for i, n in enumerate((10, 10, 300, 10, 10, 10)):
        p = tmp_path / f"e{i}.npy"
        np.save(p, np.zeros((n, 64), dtype=np.float16))
        paths.append(p)
    chunks = ion.plan_embedding_chunks(paths, max_bytes=3000)            # 10 x 64 x 2 = 1280 B, 300 rows = 38 400 B
    assert [[p.name for p in c] for c in chunks] == [["e0.npy", "e1.npy"], ["e2.npy"], ["e3.npy", "e4.npy"], ["e5.npy"]]
    assert [len(c) for c in ion.plan_embedding_chunks(paths, max_files=4)] == [4, 2]
    assert ion.plan_embedding_chunks([]) == []


def test_statistics_from_files_match_the_reference_merge(tmp_path, golden_dir, monkeypatch):
    """utils.calculate_embd_statistics_online end to end on the CPU: files -> native chunked reads -> per-file means
    -> mirrored fp16-mean rounding, with the GPU Gram accumulator replaced by exact numpy; the result must equal
    what the REAL reference's calculate_embd_statistics_online returned for the same files (tests/golden)."""
    import torch
    from fadtk_b200 import utils

    class ExactStatistics:                                     # stands in for DeviceStatistics (the GPU kernel is tested with -m gpu)
        def __init__(self, d):
            self.rows = []

        def add(self, rows):
            self.rows.append(np.asarray(rows, dtype=np.float64))

        def finalize(self):
            x = np.concatenate(self.rows)
            return torch.from_numpy(x.mean(0)), torch.from_numpy(np.cov(x, rowvar=False))

    monkeypatch.setattr(utils, "DeviceStatistics", ExactStatistics)
    monkeypatch.setattr(utils, "_BYTES_PER_READ", 2000)        # several chunks
    g = np.load(golden_dir / "stats_cases.npz")
    cat, sizes = g["cat"], g["sizes"]
    files = []
    for i, a in enumerate(np.split(cat, np.cumsum(sizes)[:-1])):
        p = tmp_path / f"f{i:03d}.npy"
        np.save(p, a)
        files.append(p)
    assert len(ion.plan_embedding_chunks(files, 2000)) > 1
    mu, cov = utils.calculate_embd_statistics_online(files)
    assert np.allclose(mu, g["mu_online"], rtol=1e-13, atol=1e-15)
    assert np.allclose(cov, g["cov_online"], rtol=1e-10, atol=1e-13)


def test_decode_container_falls_back_to_ffmpeg(tmp_path, monkeypatch):
    """Non-WAV inputs: torchaudio.load first; without a decoder backend the stream is unpacked by ffmpeg (float32 WAV at
    the native rate) and read by synth.read_wav_float.  ffmpeg is stood in for by a script that honours the same
    command line."""
    import stat
    import sys
    from scipy.io import wavfile
    from fadtk_b200 import fad as fad_mod
    rng = np.random.default_rng(4)
    audio = rng.uniform(-1, 1, size=(500, 2)).astype(np.float32)
    decoded = tmp_path / "what_ffmpeg_would_produce.wav"
    wavfile.write(decoded, 44100, audio)
    fake = tmp_path / "ffmpeg"
    fake.write_text("\n".join([
        f"#!{sys.executable}",
        "import shutil, sys",
        "args = sys.argv[1:]",
        "assert args[args.index('-acodec') + 1] == 'pcm_f32le' and args[args.index('-i') + 1].endswith('.mp3')",
        f"shutil.copy({str(decoded)!r}, args[-1])", ""]))
    fake.chmod(fake.stat().st_mode | stat.S_IXUSR)
    src = tmp_path / "song.mp3"
    src.write_bytes(b"ID3 not really an mp3")                   # torchaudio cannot decode it (and has no backend here)
    monkeypatch.setattr(fad_mod, "ffmpeg_path", str(fake))
    x, sr = fad_mod.decode_container(src)
    assert sr == 44100 and tuple(x.shape) == (2, 500) and x.dtype == torch.float32
    np.testing.assert_array_equal(x.numpy(), audio.T)
    monkeypatch.setattr(fad_mod, "ffmpeg_path", str(tmp_path / "no-such-binary"))
    with pytest.raises(RuntimeError, match="ffmpeg"):
        fad_mod.decode_container(src)
