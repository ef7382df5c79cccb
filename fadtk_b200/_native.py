"""ctypes binding of the C ABI in include/fadtk_b200.h (csrc/libfadtk_b200.so).

PyTorch tensors are only containers here: every call passes ``tensor.data_ptr()`` and the
current CUDA stream.  There is no CPU fallback - if the shared library or a B200 is missing
the import of a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np
import torch

# FADTK_B200_LIB: another build of the same library (A/B measurements of a compile-time variant on one box)
_LIB_PATH = Path(os.environ.get("FADTK_B200_LIB") or Path(__file__).parent / "csrc" / "libfadtk_b200.so")
_lib = None

c_ll = C.c_longlong
c_vp = C.c_void_p


class NativeError(RuntimeError):
    pass


class VggishWeights(C.Structure):
    _fields_ = [("conv1_w_host", c_vp), ("conv1_b_host", c_vp),
                ("conv_w_host", c_vp * 5), ("conv_b_host", c_vp * 5),
                ("fc_w_host", c_vp * 3), ("fc_b_host", c_vp * 3), ("split_mask", C.c_uint32)]


# name -> (restype, argtypes); mirrors include/fadtk_b200.h one to one
SIGNATURES = {
    "fad_version": (C.c_int, []),
    "fad_last_error": (C.c_char_p, []),
    "fad_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(c_vp)]),
    "fad_destroy": (C.c_int, [c_vp]),
    "fad_vggish_load": (C.c_int, [c_vp, C.POINTER(VggishWeights)]),
    "fad_vggish_num_examples": (c_ll, [c_ll]),
    "fad_vggish_plan": (c_ll, [c_vp, c_ll, c_vp, c_ll, c_vp]),
    "fad_vggish_forward": (C.c_int, [c_vp, c_vp, c_vp, c_ll, c_vp, c_vp]),
    "fad_vggish_logmel": (C.c_int, [c_vp, c_vp, c_vp, c_ll, c_vp, C.c_int, c_vp]),
    "fad_vggish_conv1": (C.c_int, [c_vp, c_vp, c_ll, c_vp, c_vp]),
    "fad_umma_layer": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "fad_clap_load": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int]),
    "fad_clap_plan": (c_ll, [c_vp, c_ll, c_vp, c_vp, c_ll, c_vp]),
    "fad_clap_plan_frames": (c_ll, [c_vp, c_ll, c_vp, c_vp, c_vp, c_ll, c_vp]),
    "fad_clap_forward": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp]),
    "fad_clap_logmel": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp]),
    "fad_stats_acc_len": (C.c_size_t, [C.c_int]),
    "fad_stats_accumulate": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, c_vp, c_vp, C.c_int, c_vp]),
    "fad_stats_accumulate_gather": (C.c_int, [c_vp, c_vp, c_ll, c_vp, c_ll, C.c_int, c_vp, c_vp, c_vp]),
    "fad_stats_finalize": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "fad_file_means": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "fad_stats_accumulate_f64": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, c_vp, c_vp]),
    "fad_stats_finalize_mirrored": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "fad_frechet": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "fad_sqrt_psd": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "fad_frechet_presqrt": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "fad_whisper_load": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "fad_whisper_forward": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp]),
    "fad_whisper_logmel": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp]),
    "fad_w2v_load": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int]),
    "fad_w2v_forward": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, C.c_int, c_vp, c_vp]),
    "fad_encodec_load": (C.c_int, [c_vp, c_vp, C.c_int, c_ll, C.c_int]),
    "fad_encodec_forward": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, c_vp, c_vp]),
    "fad_resample_geometry": (C.c_int, [C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "fad_resample_length": (c_ll, [C.c_int, C.c_int, c_ll]),
    "fad_resample_bank": (C.c_int, [C.c_int, C.c_int, c_vp]),
    "fad_resample": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_ll, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "fad_frechet_batched": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, C.c_int, C.c_int, c_vp, c_vp]),
    "fad_attention": (C.c_int, [c_vp, c_vp, c_ll, C.c_int, C.c_int, c_vp, C.c_int, c_vp]),
    "fad_bench_dmma_peak": (C.c_int, [c_vp, C.c_int, c_vp]),
    "fad_bench_umma_mode": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "fad_comm_unique_id": (C.c_int, [c_vp]),
    "fad_comm_init": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int]),
    "fad_comm_destroy": (C.c_int, [c_vp]),
    "fad_stats_allreduce": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_vp]),
    "fad_allreduce_sum_f64": (C.c_int, [c_vp, c_vp, c_vp, c_ll, c_vp]),
    "fad_launch_count": (c_ll, [c_vp]),
    "fad_profile_enable": (C.c_int, [c_vp, C.c_int]),
    "fad_profile_collect": (C.c_int, [c_vp, c_vp, c_vp, C.c_int]),
}

PROF_CATEGORIES = 20
PROF_NAMES = {0: "logmel", 1: "conv1", 2: "conv2", 3: "conv3_1", 4: "conv3_2", 5: "conv4_1", 6: "conv4_2",
              7: "fc1", 8: "fc2", 9: "fc3", 10: "stats", 11: "stats_reduce", 12: "frechet",
              13: "clap_front", 14: "clap_gemm", 15: "clap_attn", 16: "clap_other"}


def library_path() -> Path:
    return _LIB_PATH


def lib():
    """Load the shared library (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeError(
                f"{_LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "fadtk_b200 has no CPU fallback.")
        _lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def _check(rc: int):
    if rc != 0:
        raise NativeError(lib().fad_last_error().decode())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class Engine:
    """One native handle bound to one CUDA device."""

    def __init__(self, device: int | None = None, max_examples: int = 2048):
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device visible: fadtk_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_examples = int(max_examples)
        h = c_vp()
        _check(lib().fad_create(self.device, self.max_examples, C.byref(h)))
        self._h = h
        self._keep = []
        self.has_comm = False     # fad_comm_init done: the statistics all-reduce goes through the C ABI
        self.owners = {}          # weight slot -> token of the loader whose weights it holds (model_loader._DeviceBatch)

    def close(self):
        if getattr(self, "_h", None):
            lib().fad_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    @property
    def launches(self) -> int:
        return int(lib().fad_launch_count(self._h))

    # ------------------------------------------------------------ cross-GPU merge
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: the 128-byte NCCL rendezvous id (fad_comm_unique_id)"""
        buf = C.create_string_buffer(128)
        _check(lib().fad_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """join the communicator of the C ABI (fad_comm_init); afterwards allreduce_sum_ runs through it"""
        assert len(unique_id) == 128
        _check(lib().fad_comm_init(self._h, C.create_string_buffer(unique_id, 128), int(rank), int(world)))
        self.has_comm = True

    def allreduce_sum_(self, buf: torch.Tensor) -> torch.Tensor:
        """in-place sum of an fp64 cuda tensor over the ranks of the handle's communicator (fad_allreduce_sum_f64)"""
        assert buf.dtype == torch.float64 and buf.is_cuda and buf.is_contiguous()
        _check(lib().fad_allreduce_sum_f64(self._h, None, buf.data_ptr(), buf.numel(), _stream()))
        return buf

    def attention(self, qkv: torch.Tensor, n_clips: int, legacy: bool = False) -> torch.Tensor:
        """qkv fp16 [n_clips * S, 3 d] (cuda) -> fp16 [n_clips * S, d]: per-head softmax(q k^T / 8) v, heads of 64 dims"""
        assert qkv.dtype == torch.float16 and qkv.is_cuda and qkv.is_contiguous() and qkv.shape[0] % n_clips == 0
        S, d = qkv.shape[0] // n_clips, qkv.shape[1] // 3
        out = torch.empty((qkv.shape[0], d), dtype=torch.float16, device=qkv.device)
        _check(lib().fad_attention(self._h, qkv.data_ptr(), n_clips, S, d, out.data_ptr(), int(legacy), _stream()))
        return out

    def umma_mode_ms(self, mode: int, ksteps: int = 40000) -> float:
        """tensor-pipe microbenchmark (csrc/umma_bench.cuh): ms for `ksteps` split-weight K steps per SM under issue pattern `mode`"""
        out = C.c_double(0.0)
        _check(lib().fad_bench_umma_mode(self._h, int(mode), int(ksteps), C.byref(out)))
        return float(out.value)

    def dmma_peak_tflops(self, iters: int = 0) -> float:
        """measured fp64 tensor-pipe (DMMA) rate, TFLOP/s: roofline denominator of the fp64 kernels"""
        out = C.c_double(0.0)
        _check(lib().fad_bench_dmma_peak(self._h, int(iters), C.byref(out)))
        return float(out.value)

    def profile(self, on: bool):
        _check(lib().fad_profile_enable(self._h, int(on)))

    def profile_collect(self, reset: bool = True) -> dict:
        """-> {category: (milliseconds, launches)} measured with CUDA events on the launch stream."""
        ms = (C.c_double * PROF_CATEGORIES)()
        cnt = (c_ll * PROF_CATEGORIES)()
        _check(lib().fad_profile_collect(self._h, ms, cnt, int(reset)))
        return {PROF_NAMES[i]: (ms[i], int(cnt[i])) for i in PROF_NAMES if cnt[i]}

    # ------------------------------------------------------------------ VGGish
    def vggish_load(self, packed: dict):
        """``packed`` comes from fadtk_b200.weights.pack_vggish (CPU tensors)."""
        self.owners.pop("vggish", None)          # whoever loads claims the slot afterwards (model_loader._DeviceBatch)
        w = VggishWeights()
        keep = {k: v.contiguous() for k, v in packed.items() if hasattr(v, "contiguous")}
        w.conv1_w_host = keep["conv1.w"].data_ptr()
        w.conv1_b_host = keep["conv1.b"].data_ptr()
        for i in range(5):
            w.conv_w_host[i] = keep[f"conv{i + 2}.w"].data_ptr()
            w.conv_b_host[i] = keep[f"conv{i + 2}.b"].data_ptr()
        for i in range(3):
            w.fc_w_host[i] = keep[f"fc{i + 1}.w"].data_ptr()
            w.fc_b_host[i] = keep[f"fc{i + 1}.b"].data_ptr()
        w.split_mask = int(packed.get("split_mask", 0))
        _check(lib().fad_vggish_load(self._h, C.byref(w)))

    @staticmethod
    def vggish_plan(clip_offsets: np.ndarray):
        """-> (ex_start int64 [n_examples], rows_per_clip int64 [n_clips])"""
        off = np.ascontiguousarray(clip_offsets, dtype=np.int64)
        n_clips = off.shape[0] - 1
        rows = np.empty(n_clips, dtype=np.int64)
        n = lib().fad_vggish_plan(off.ctypes.data, n_clips, None, 0, rows.ctypes.data)
        ex = np.empty(n, dtype=np.int64)
        lib().fad_vggish_plan(off.ctypes.data, n_clips, ex.ctypes.data, n, None)
        return ex, rows

    def vggish_forward(self, pcm: torch.Tensor, ex_start: torch.Tensor, out: torch.Tensor | None = None):
        """pcm int16 [samples] (cuda), ex_start int64 [n] (cuda) -> fp16 [n, 128] (cuda)."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda and pcm.is_contiguous()
        assert ex_start.dtype == torch.int64 and ex_start.is_cuda
        n = ex_start.shape[0]
        if out is None:
            out = torch.empty((n, 128), dtype=torch.float16, device=pcm.device)
        assert out.dtype == torch.float16 and out.is_contiguous() and out.shape[0] >= n
        _check(lib().fad_vggish_forward(self._h, pcm.data_ptr(), ex_start.data_ptr(), n,
                                        out.data_ptr(), _stream()))
        return out[:n]

    def vggish_logmel(self, pcm, ex_start, use_double=True):
        n = ex_start.shape[0]
        out = torch.empty((n, 96, 64), dtype=torch.float32, device=pcm.device)
        _check(lib().fad_vggish_logmel(self._h, pcm.data_ptr(), ex_start.data_ptr(), n,
                                       out.data_ptr(), int(use_double), _stream()))
        return out

    def vggish_conv1(self, logmel):
        """fp32 [n, 96, 64] log-mel examples -> fp16 NHWC [n, 48, 32, 64] (conv1 + ReLU + max-pool, loaded weights)"""
        assert logmel.dtype == torch.float32 and logmel.is_contiguous() and logmel.shape[1:] == (96, 64)
        n = logmel.shape[0]
        out = torch.empty((n, 48, 32, 64), dtype=torch.float16, device=logmel.device)
        _check(lib().fad_vggish_conv1(self._h, logmel.data_ptr(), n, out.data_ptr(), _stream()))
        return out

    def umma_layer(self, x, w, bias, taps, relu, pool, want_f32=False, split_w=False):
        """x fp16 NHWC [NB,H,W,Cin]; w fp16 [Cout, taps*Cin] (or [2*Cout, taps*Cin] hi/lo tiles when
        split_w, see weights.split_hi_lo_tiles); bias fp32 [Cout]."""
        nb, hh, ww, cin = x.shape
        cout = w.shape[0] // (2 if split_w else 1)
        oh, ow = (hh // 2, ww // 2) if pool else (hh, ww)
        out = torch.empty((nb, oh, ow, cout), dtype=torch.float16, device=x.device)
        out32 = torch.empty((nb, oh, ow, cout), dtype=torch.float32, device=x.device) if want_f32 else None
        _check(lib().fad_umma_layer(self._h, x.data_ptr(), nb, hh, ww, cin, w.data_ptr(), bias.data_ptr(),
                                    cout, taps, int(relu), int(pool), int(split_w), out.data_ptr(), _ptr(out32),
                                    _stream()))
        return (out, out32) if want_f32 else out

    # -------------------------------------------------------------------- CLAP
    def clap_load(self, tensors: list, max_chunks: int = 32):
        """``tensors`` comes from fadtk_b200.weights_clap.pack_clap (CPU tensors, fixed order)."""
        self.owners.pop("clap", None)          # whoever loads claims the slot afterwards (model_loader._DeviceBatch)
        keep = [t.contiguous() for t in tensors]
        arr = (c_vp * len(keep))(*[t.data_ptr() for t in keep])
        _check(lib().fad_clap_load(self._h, arr, len(keep), int(max_chunks)))

    @staticmethod
    def clap_plan(clip_offsets: np.ndarray):
        """-> (chunk_start int64 [n], chunk_valid int32 [n], rows_per_clip int64 [n_clips])"""
        off = np.ascontiguousarray(clip_offsets, dtype=np.int64)
        n_clips = off.shape[0] - 1
        rows = np.empty(n_clips, dtype=np.int64)
        n = lib().fad_clap_plan(off.ctypes.data, n_clips, None, None, 0, rows.ctypes.data)
        start = np.empty(n, dtype=np.int64)
        valid = np.empty(n, dtype=np.int32)
        lib().fad_clap_plan(off.ctypes.data, n_clips, start.ctypes.data, valid.ctypes.data, n, None)
        return start, valid, rows

    @staticmethod
    def clap_plan_frames(clip_offsets: np.ndarray):
        """-> dict(pool_start int64, pool_valid int32, pool_frame int32, frame_index int32 [n_chunks,1001],
        rows_per_clip int64): every distinct STFT frame once + the per-window index table."""
        off = np.ascontiguousarray(clip_offsets, dtype=np.int64)
        n_clips = off.shape[0] - 1
        rows = np.empty(n_clips, dtype=np.int64)
        n_chunks = lib().fad_clap_plan(off.ctypes.data, n_clips, None, None, 0, rows.ctypes.data)
        n_pool = lib().fad_clap_plan_frames(off.ctypes.data, n_clips, None, None, None, 0, None)
        ps = np.empty(n_pool, dtype=np.int64)
        pv = np.empty(n_pool, dtype=np.int32)
        pf = np.empty(n_pool, dtype=np.int32)
        fi = np.empty((n_chunks, 1001), dtype=np.int32)
        lib().fad_clap_plan_frames(off.ctypes.data, n_clips, ps.ctypes.data, pv.ctypes.data, pf.ctypes.data, n_pool,
                                   fi.ctypes.data)
        return {"pool_start": ps, "pool_valid": pv, "pool_frame": pf, "frame_index": fi, "rows_per_clip": rows}

    def clap_plan_to_device(self, plan: dict) -> dict:
        dev = self.torch_device
        return {k: (torch.from_numpy(v).to(dev) if k != "rows_per_clip" else v) for k, v in plan.items()}

    def clap_forward(self, pcm: torch.Tensor, plan_dev: dict):
        """pcm int16 (cuda, 48 kHz); plan_dev from clap_plan_to_device -> fp16 [n_chunks, 512]."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda
        fi = plan_dev["frame_index"]
        n = fi.shape[0]
        out = torch.empty((n, 512), dtype=torch.float16, device=pcm.device)
        _check(lib().fad_clap_forward(self._h, pcm.data_ptr(), plan_dev["pool_start"].data_ptr(),
                                      plan_dev["pool_valid"].data_ptr(), plan_dev["pool_frame"].data_ptr(),
                                      plan_dev["pool_start"].shape[0], fi.data_ptr(), n, out.data_ptr(), _stream()))
        return out

    def clap_logmel(self, pcm, plan_dev: dict):
        """-> fp32 [n_chunks, 1001, 64] BatchNorm-ed log-mel, gathered from the frame pool."""
        n_pool = plan_dev["pool_start"].shape[0]
        pool = torch.empty((n_pool, 64), dtype=torch.float32, device=pcm.device)
        _check(lib().fad_clap_logmel(self._h, pcm.data_ptr(), plan_dev["pool_start"].data_ptr(),
                                     plan_dev["pool_valid"].data_ptr(), plan_dev["pool_frame"].data_ptr(), n_pool,
                                     pool.data_ptr(), _stream()))
        return pool[plan_dev["frame_index"].long()]

    # ------------------------------------------------------------------ Whisper
    def whisper_load(self, cfg: tuple, tensors: list, max_clips: int = 16):
        """cfg = (d_model, heads, enc_layers, dec_layers, ffn); tensors from weights_whisper.pack_whisper."""
        self.owners.pop("whisper", None)          # whoever loads claims the slot afterwards (model_loader._DeviceBatch)
        keep = [t.contiguous() for t in tensors]
        arr = (c_vp * len(keep))(*[t.data_ptr() for t in keep])
        c = (C.c_int * 5)(*[int(v) for v in cfg])
        _check(lib().fad_whisper_load(self._h, c, arr, len(keep), int(max_clips)))
        self._whisper_d = int(cfg[0])

    def whisper_forward(self, pcm: torch.Tensor, clip_start: torch.Tensor, clip_len: torch.Tensor) -> torch.Tensor:
        """pcm int16 (cuda, 16 kHz); clip_start int64 / clip_len int32 [n] (cuda) -> fp16 [n, 2, d_model]."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda and clip_start.dtype == torch.int64 and clip_len.dtype == torch.int32
        n = clip_start.shape[0]
        out = torch.empty((n, 2, self._whisper_d), dtype=torch.float16, device=pcm.device)
        _check(lib().fad_whisper_forward(self._h, pcm.data_ptr(), clip_start.data_ptr(), clip_len.data_ptr(), n,
                                         out.data_ptr(), _stream()))
        return out

    def whisper_features(self, pcm: torch.Tensor, clip_start: torch.Tensor, clip_len: torch.Tensor) -> torch.Tensor:
        """-> fp32 [n, 3000, 80]: the feature extractor's input_features (time-major)."""
        n = clip_start.shape[0]
        buf = torch.empty(n * 3000 * 80 + n, dtype=torch.float32, device=pcm.device)
        _check(lib().fad_whisper_logmel(self._h, pcm.data_ptr(), clip_start.data_ptr(), clip_len.data_ptr(), n,
                                        buf.data_ptr(), _stream()))
        raw = buf[: n * 3000 * 80].view(n, 3000, 80)
        mx = buf[n * 3000 * 80:].view(n, 1, 1)
        return (torch.maximum(raw, mx - 8.0) + 4.0) / 4.0

    # ------------------------------------------------------- wav2vec 2.0 / HuBERT / MERT
    def w2v_load(self, cfg: tuple, tensors: list, max_clips: int = 8, max_len: int = 16000 * 30):
        """cfg = weights_w2v.config_of(state); tensors from weights_w2v.pack_w2v."""
        self.owners.pop("w2v", None)          # whoever loads claims the slot afterwards (model_loader._DeviceBatch)
        keep = [t.contiguous() for t in tensors]
        arr = (c_vp * len(keep))(*[t.data_ptr() for t in keep])
        c = (C.c_int * 7)(*[int(v) for v in cfg])
        _check(lib().fad_w2v_load(self._h, c, arr, len(keep), int(max_clips), int(max_len)))
        self._w2v_d = int(cfg[0])

    @staticmethod
    def w2v_frames(n_samples: int) -> int:
        t = n_samples
        for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
            t = (t - k) // s + 1
        return t

    def w2v_forward(self, pcm: torch.Tensor, layer: int) -> torch.Tensor:
        """pcm int16 [n_clips, L] (cuda, equal lengths) -> fp16 [n_clips, frames, d_model] = hidden_states[layer]."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda and pcm.ndim == 2 and pcm.is_contiguous()
        n, L = pcm.shape
        out = torch.empty((n, self.w2v_frames(L), self._w2v_d), dtype=torch.float16, device=pcm.device)
        _check(lib().fad_w2v_forward(self._h, pcm.data_ptr(), n, L, int(layer), out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ Encodec
    def encodec_load(self, tensors: list, max_chunk_samples: int = 16 * 240000, variant: str = "24k"):
        self.owners.pop("encodec", None)          # whoever loads claims the slot afterwards (model_loader._DeviceBatch)
        keep = [t.contiguous() for t in tensors]
        arr = (c_vp * len(keep))(*[t.data_ptr() for t in keep])
        _check(lib().fad_encodec_load(self._h, arr, len(keep), int(max_chunk_samples), 0 if variant == "24k" else 1))

    def encodec_forward(self, pcm: torch.Tensor) -> torch.Tensor:
        """pcm int16 [n_clips, T] (cuda, 24 kHz, equal lengths) -> fp16 [n_clips, ceil(T/320), 128]."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda and pcm.ndim == 2 and pcm.is_contiguous()
        n, T = pcm.shape
        frames = T
        for r in (2, 4, 5, 8):
            frames = -(-frames // r)
        out = torch.empty((n, frames, 128), dtype=torch.float16, device=pcm.device)
        _check(lib().fad_encodec_forward(self._h, pcm.data_ptr(), n, T, out.data_ptr(), _stream()))
        return out

    # -------------------------------------------------------------- audio conversion
    @staticmethod
    def resample_geometry(sr_in: int, sr_out: int):
        """-> (orig, new, width, taps) of torchaudio's polyphase resampler for this rate pair."""
        v = [C.c_int() for _ in range(4)]
        _check(lib().fad_resample_geometry(int(sr_in), int(sr_out), *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    @staticmethod
    def resample_bank(sr_in: int, sr_out: int) -> np.ndarray:
        """float32 [new, taps] filter bank (host only; what fad_resample uploads)."""
        _, new, _, taps = Engine.resample_geometry(sr_in, sr_out)
        bank = np.empty((new, taps), dtype=np.float32)
        _check(lib().fad_resample_bank(int(sr_in), int(sr_out), bank.ctypes.data))
        return bank

    def resample(self, x: torch.Tensor, sr_in: int, sr_out: int, return_float: bool = False):
        """x: int16 [length, channels] / [length] (PCM16, cuda) or float32 [channels, length] (cuda)
        -> int16 [ceil(new*length/orig)] mono at sr_out (and the un-quantised float32 if asked)."""
        assert x.is_cuda
        if x.dtype == torch.int16:
            x = x.contiguous()
            length, channels = x.shape[0], (1 if x.ndim == 1 else x.shape[1])
            pi, pf = x.data_ptr(), None
        else:
            assert x.dtype == torch.float32 and x.ndim == 2
            x = x.contiguous()
            channels, length = x.shape
            pi, pf = None, x.data_ptr()
        n_out = int(lib().fad_resample_length(int(sr_in), int(sr_out), length))
        out = torch.empty(n_out, dtype=torch.int16, device=x.device)
        outf = torch.empty(n_out, dtype=torch.float32, device=x.device) if return_float else None
        _check(lib().fad_resample(self._h, pi, pf, channels, length, int(sr_in), int(sr_out), out.data_ptr(),
                                  outf.data_ptr() if outf is not None else None, _stream()))
        return (out, outf) if return_float else out

    # -------------------------------------------------------------- statistics
    @staticmethod
    def stats_acc_len(d: int) -> int:
        return int(lib().fad_stats_acc_len(d))

    def stats_new(self, d: int) -> torch.Tensor:
        return torch.zeros(self.stats_acc_len(d), dtype=torch.float64, device=self.torch_device)

    def stats_accumulate(self, emb, shift, acc, tensor_core=False):
        assert emb.dtype == torch.float16 and emb.is_contiguous() and shift.dtype == torch.float16
        n, d = emb.shape
        _check(lib().fad_stats_accumulate(self._h, emb.data_ptr(), n, d, shift.data_ptr(),
                                          acc.data_ptr(), int(tensor_core), _stream()))
        return acc

    def stats_accumulate_gather(self, emb, idx, shift, acc):
        assert emb.dtype == torch.float16 and emb.is_contiguous() and idx.dtype == torch.int64
        n, d = emb.shape
        _check(lib().fad_stats_accumulate_gather(self._h, emb.data_ptr(), n, idx.data_ptr(), idx.shape[0], d,
                                                 shift.data_ptr(), acc.data_ptr(), _stream()))
        return acc

    def stats_finalize(self, acc, shift, d):
        mu = torch.empty(d, dtype=torch.float64, device=acc.device)
        cov = torch.empty((d, d), dtype=torch.float64, device=acc.device)
        _check(lib().fad_stats_finalize(self._h, acc.data_ptr(), shift.data_ptr(), d,
                                        mu.data_ptr(), cov.data_ptr(), _stream()))
        return mu, cov

    def file_means(self, emb: torch.Tensor, rows_per_file: int):
        """fp16 [n_files * r, d] -> (m64, m16) fp64 [n_files, d]: exact per-file means and the reference's fp16-rounded ones"""
        n_files, d = emb.shape[0] // rows_per_file, emb.shape[1]
        m64 = torch.empty((n_files, d), dtype=torch.float64, device=emb.device)
        m16 = torch.empty_like(m64)
        _check(lib().fad_file_means(self._h, emb.data_ptr(), n_files, rows_per_file, d, m64.data_ptr(), m16.data_ptr(), _stream()))
        return m64, m16

    def stats_accumulate_f64(self, rows: torch.Tensor, acc: torch.Tensor) -> torch.Tensor:
        assert rows.dtype == torch.float64 and rows.is_contiguous()
        _check(lib().fad_stats_accumulate_f64(self._h, rows.data_ptr(), rows.shape[0], rows.shape[1], acc.data_ptr(), _stream()))
        return acc

    def stats_finalize_mirrored(self, acc, acc64, acc16, shift, rows_per_file: int, d: int):
        """(mu, cov) as the reference's calculate_embd_statistics_online gives them for equal-length files (utils.py:13-46)"""
        mu = torch.empty(d, dtype=torch.float64, device=acc.device)
        cov = torch.empty((d, d), dtype=torch.float64, device=acc.device)
        _check(lib().fad_stats_finalize_mirrored(self._h, acc.data_ptr(), acc64.data_ptr(), acc16.data_ptr(), shift.data_ptr(),
                                                 int(rows_per_file), d, mu.data_ptr(), cov.data_ptr(), _stream()))
        return mu, cov

    # ----------------------------------------------------------------- Frechet
    def frechet(self, mu1, cov1, mu2, cov2, iters: int = 0) -> torch.Tensor:
        """fp64 cuda tensors -> fp64 [8] cuda: FAD, tr sqrt, residual, iters, |dmu|^2, trC1, trC2."""
        d = mu1.shape[0]
        for t in (mu1, cov1, mu2, cov2):
            assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        out = torch.zeros(8, dtype=torch.float64, device=mu1.device)
        _check(lib().fad_frechet(self._h, mu1.data_ptr(), cov1.data_ptr(), mu2.data_ptr(), cov2.data_ptr(),
                                 d, iters, out.data_ptr(), _stream()))
        return out


class Baseline:
    """Device-resident baseline statistics with the matrix square root precomputed."""

    def __init__(self, eng: "Engine", mu, cov):
        dev = eng.torch_device
        self.eng = eng
        def to_dev(a):
            if isinstance(a, torch.Tensor):
                return a.to(dev, torch.float64).contiguous()
            return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        self.mu = to_dev(mu)
        cov = to_dev(cov)
        self.d = self.mu.shape[0]
        self.sqrt = torch.empty((self.d, self.d), dtype=torch.float64, device=dev)
        self.scal = torch.empty(2, dtype=torch.float64, device=dev)
        _check(lib().fad_sqrt_psd(eng._h, cov.data_ptr(), self.d, 0, self.sqrt.data_ptr(), self.scal.data_ptr(), _stream()))

    def frechet(self, mu2: torch.Tensor, cov2: torch.Tensor) -> torch.Tensor:
        """fp64 device tensors -> fp64 [8] device (same layout as Engine.frechet)."""
        out = torch.zeros(8, dtype=torch.float64, device=self.mu.device)
        _check(lib().fad_frechet_presqrt(self.eng._h, self.mu.data_ptr(), self.sqrt.data_ptr(), self.scal.data_ptr(),
                                         mu2.data_ptr(), cov2.data_ptr(), self.d, 0, out.data_ptr(), _stream()))
        return out

    def frechet_batched(self, emb: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """emb fp16 [N, d], offsets int64 [n_items + 1] (both cuda) -> fp64 [n_items, 8]: every item's
        FAD against this baseline in one lock-step launch sequence (fad_frechet_batched)."""
        assert emb.dtype == torch.float16 and emb.is_cuda and emb.is_contiguous() and emb.shape[1] == self.d
        assert offsets.dtype == torch.int64 and offsets.is_cuda
        n_items = offsets.shape[0] - 1
        out = torch.zeros((n_items, 8), dtype=torch.float64, device=emb.device)
        _check(lib().fad_frechet_batched(self.eng._h, self.mu.data_ptr(), self.sqrt.data_ptr(), self.scal.data_ptr(),
                                         emb.data_ptr(), offsets.data_ptr(), n_items, self.d, 0, out.data_ptr(), _stream()))
        return out


_engines: dict = {}


def engine(device: int | None = None, max_examples: int | None = None) -> Engine:
    """Process-wide engine per device (created lazily)."""
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _engines:
        me = max_examples or int(os.environ.get("FADTK_MAX_EXAMPLES", "2048"))
        _engines[dev] = Engine(dev, me)
    return _engines[dev]
