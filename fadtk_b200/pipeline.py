"""In-memory hot path: PCM -> embeddings -> statistics -> Frechet distance, no filesystem.

The reference moves everything between stages through files (.wav -> .npy -> mu/cov.npy,
SURVEY.md section 1).  ``cache_embedding_files`` / ``FrechetAudioDistance`` keep that contract;
this module is the same arithmetic with the stages chained on one CUDA stream, used when the
caller already holds PCM in memory (and by bench.py).  Under torch.distributed every rank feeds
its own shard of clips and ONE all-reduce of the packed statistics precedes the Frechet chain.

Semantics = the reference's directory flow with one clip per file: embeddings are rounded to
fp16 (model_loader.py:47-48), per-file means are rounded to fp16 before the merge
(utils.py:13-46) - mirrored here from per-clip means, see utils.mirror_file_mean_rounding.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native, dist


class EvalSetFAD:
    """FAD of equal-length PCM16 clips against fixed baseline statistics.

    ``model`` (the engine must already hold that model's weights):
    "vggish" (16 kHz, 128-d, one row per 0.96 s); "clap-laion-audio" / "clap-laion-music" (48 kHz, 512-d,
    one row per started second); "encodec-emb" (24 kHz, 128-d, one row per 320 samples); "whisper-<size>"
    (16 kHz, two rows per clip); "w2v2-", "hubert-", "wavlm-", "MERT-" names (16 / 24 kHz, 768- or 1024-d,
    one row per 20 ms; a trailing "-<k>" selects hidden_states[k], default the last layer).
    """

    def __init__(self, engine: _native.Engine, mu_base: torch.Tensor, cov_base: torch.Tensor,
                 clip_samples: int, clips_per_chunk: int = 1000, mirror_file_means: bool = True,
                 model: str = "vggish"):
        self.eng = engine
        self.dev = engine.torch_device
        self.model = model
        self.mu_base = mu_base.to(self.dev, torch.float64).contiguous()
        self.cov_base = cov_base.to(self.dev, torch.float64).contiguous()
        self.clip_samples = int(clip_samples)
        self.clips_per_chunk = int(clips_per_chunk)
        self.w2v_layer = None
        if model == "vggish":
            self.d = 128
            self.rows_per_clip = int(_native.lib().fad_vggish_num_examples(self.clip_samples))
        elif model in ("clap-laion-audio", "clap-laion-music"):
            self.d = 512
            self.rows_per_clip = -(-self.clip_samples // 48000)
        elif model == "encodec-emb":
            self.d = 128
            r = self.clip_samples
            for ratio in (2, 4, 5, 8):                       # ceil at every down-sampling conv
                r = -(-r // ratio)
            self.rows_per_clip = r
        elif model.split("-")[0] in ("w2v2", "hubert", "wavlm", "MERT"):
            self.d = 1024 if "large" in model else 768
            self.rows_per_clip = _native.Engine.w2v_frames(self.clip_samples)
            tail = model.rsplit("-", 1)[-1]
            self.w2v_layer = int(tail) if tail.isdigit() else (24 if "large" in model else 12)
        elif model.startswith("whisper-"):
            from .model_loader import WhisperModel
            self.d = WhisperModel.DIMS[model.split("-", 1)[1]]
            self.rows_per_clip = 2                           # last_hidden_state of the two start tokens
        else:
            raise ValueError(model)
        self.mirror = mirror_file_means
        self.shift = None
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._staging = None
        self._plans = {}
        self._baseline = None
        dist.enable_native_allreduce(self.eng)               # multi-GPU: the statistics all-reduce runs through the C ABI

    def _plan(self, n_clips: int):
        if n_clips not in self._plans:
            off = np.arange(n_clips + 1, dtype=np.int64) * self.clip_samples
            if self.model == "vggish":
                ex, _ = self.eng.vggish_plan(off)
                self._plans[n_clips] = (torch.from_numpy(ex).to(self.dev),)
            elif self.model == "encodec-emb" or self.w2v_layer is not None:
                self._plans[n_clips] = ()
            elif self.model.startswith("whisper-"):
                self._plans[n_clips] = (torch.from_numpy(off[:-1].copy()).to(self.dev),
                                        torch.full((n_clips,), self.clip_samples, dtype=torch.int32, device=self.dev))
            else:
                self._plans[n_clips] = (self.eng.clap_plan_to_device(self.eng.clap_plan_frames(off)),)
        return self._plans[n_clips]

    def embed(self, pcm_dev: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """int16 [n_clips, clip_samples] (device) -> fp16 [n_clips * rows_per_clip, d]"""
        plan = self._plan(pcm_dev.shape[0])
        flat = pcm_dev.reshape(-1)
        if self.model == "vggish":
            return self.eng.vggish_forward(flat, plan[0], out)
        if self.w2v_layer is not None:
            emb = self.eng.w2v_forward(pcm_dev.contiguous(), self.w2v_layer).reshape(-1, self.d)
        elif self.model == "encodec-emb":
            emb = self.eng.encodec_forward(pcm_dev.contiguous()).reshape(-1, self.d)
        elif self.model.startswith("whisper-"):
            emb = self.eng.whisper_forward(flat, plan[0], plan[1]).reshape(-1, self.d)
        else:
            emb = self.eng.clap_forward(flat, plan[0])
        if out is not None:
            out.copy_(emb)
            return out
        return emb

    def _shared_shift(self, emb: torch.Tensor) -> torch.Tensor:
        if self.shift is None:
            # every rank must use the same vector: rank 0's first-rows mean
            s = emb[:4096].float().mean(0)
            if dist.is_distributed():
                torch.distributed.broadcast(s, src=0)
            self.shift = s.to(torch.float16)
        return self.shift

    def score(self, emb: torch.Tensor) -> torch.Tensor:
        """fp16 [n_clips * rows_per_clip, d] (this rank's shard) -> fp64[8] (device)."""
        d, r = self.d, self.rows_per_clip
        shift = self._shared_shift(emb)
        n_acc = self.eng.stats_acc_len(d)
        # ONE buffer = ONE all-reduce: [rows | exact per-clip means | fp16-rounded per-clip means], all packed accumulators
        buf = torch.zeros(n_acc * (3 if self.mirror else 1), dtype=torch.float64, device=self.dev)
        acc = buf[:n_acc]
        self.eng.stats_accumulate(emb, shift, acc)
        if self.mirror:
            # the reference's per-file statistics (one clip = one file): np.mean of an fp16 file is fp16 (utils.py:14)
            m64, m16 = self.eng.file_means(emb, r)
            self.eng.stats_accumulate_f64(m64, buf[n_acc:2 * n_acc])
            self.eng.stats_accumulate_f64(m16, buf[2 * n_acc:])
        dist.allreduce_sum_(buf)                              # the only cross-GPU exchange
        if self.mirror:
            mu, cov = self.eng.stats_finalize_mirrored(acc, buf[n_acc:2 * n_acc], buf[2 * n_acc:], shift, r, d)
        else:
            mu, cov = self.eng.stats_finalize(acc, shift, d)
        if self._baseline is None:                            # sqrt(C_base) once per baseline, not per eval set
            self._baseline = _native.Baseline(self.eng, self.mu_base, self.cov_base)
        return self._baseline.frechet(mu.contiguous(), cov.contiguous())

    def run_device(self, pcm_dev: torch.Tensor) -> torch.Tensor:
        """pcm_dev int16 [n_clips, clip_samples] resident in HBM -> fp64[8] result (device)."""
        n_clips = pcm_dev.shape[0]
        cpc = min(self.clips_per_chunk, n_clips)
        if cpc == n_clips:
            return self.score(self.embed(pcm_dev))
        emb_all = torch.empty((n_clips * self.rows_per_clip, self.d), dtype=torch.float16, device=self.dev)
        for s in range(0, n_clips, cpc):
            c = min(cpc, n_clips - s)
            self.embed(pcm_dev[s:s + c], emb_all[s * self.rows_per_clip:(s + c) * self.rows_per_clip])
        return self.score(emb_all)

    def run_host(self, pcm_host: torch.Tensor) -> float:
        """pcm_host: PINNED int16 [n_clips, clip_samples].  H2D copies (double-buffered on a copy
        stream) overlap the forward of the previous chunk; returns the FAD as a Python float
        (device -> host read of the result)."""
        assert pcm_host.is_pinned() and pcm_host.dtype == torch.int16
        n_clips = pcm_host.shape[0]
        cpc = min(self.clips_per_chunk, n_clips)
        if self._staging is None or self._staging[0].shape[0] < cpc:
            self._staging = [torch.empty((cpc, self.clip_samples), dtype=torch.int16, device=self.dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream(self.dev)
        emb_all = torch.empty((n_clips * self.rows_per_clip, self.d), dtype=torch.float16, device=self.dev)
        for i, s in enumerate(range(0, n_clips, cpc)):
            b = i & 1
            c = min(cpc, n_clips - s)
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(self._free[b])
                self._staging[b][:c].copy_(pcm_host[s:s + c], non_blocking=True)
                self._ready[b].record(self._copy_stream)
            main.wait_event(self._ready[b])
            self.embed(self._staging[b][:c], emb_all[s * self.rows_per_clip:(s + c) * self.rows_per_clip])
            self._free[b].record(main)
        return float(self.score(emb_all)[0].item())
