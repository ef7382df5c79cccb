"""In-memory hot path: PCM -> VGGish embeddings -> statistics -> Frechet distance, no filesystem.

The reference moves everything between stages through files (.wav -> .npy -> mu/cov.npy,
SURVEY.md section 1).  ``cache_embedding_files`` / ``FrechetAudioDistance`` keep that contract;
this module is the same arithmetic with the stages chained on one CUDA stream, used when the
caller already holds PCM in memory (and by bench.py).  Under torch.distributed every rank feeds
its own shard of clips and the packed statistics are all-reduced once before the Frechet chain.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native, dist
from .utils import DeviceStatistics


class EvalSetFAD:
    """FAD of equal-length PCM16 clips against fixed baseline statistics."""

    def __init__(self, engine: _native.Engine, mu_base: torch.Tensor, cov_base: torch.Tensor,
                 clip_samples: int, clips_per_chunk: int = 1000):
        self.eng = engine
        self.dev = engine.torch_device
        self.mu_base = mu_base.to(self.dev, torch.float64).contiguous()
        self.cov_base = cov_base.to(self.dev, torch.float64).contiguous()
        self.clip_samples = int(clip_samples)
        self.clips_per_chunk = int(clips_per_chunk)
        self.rows_per_clip = int(_native.lib().fad_vggish_num_examples(self.clip_samples))
        self.shift = None
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._staging = None

    def _plan(self, n_clips: int) -> torch.Tensor:
        off = np.arange(n_clips + 1, dtype=np.int64) * self.clip_samples
        ex, _ = self.eng.vggish_plan(off)
        return torch.from_numpy(ex).to(self.dev, non_blocking=True)

    def _finish(self, st: DeviceStatistics) -> torch.Tensor:
        st.allreduce()                                   # one NCCL all-reduce of d^2+2d+1 doubles
        mu, cov = st.finalize()
        return self.eng.frechet(self.mu_base, self.cov_base, mu, cov)

    def _stats(self, emb_first: torch.Tensor) -> DeviceStatistics:
        st = DeviceStatistics(128, self.eng)
        if self.shift is None:
            # shared shift: every rank must use the same vector, take rank 0's first-chunk mean
            s = emb_first[:4096].float().mean(0)
            if dist.is_distributed():
                torch.distributed.broadcast(s, src=0)
            self.shift = s.to(torch.float16)
        st.wide = False
        st.shift = self.shift
        st.acc = self.eng.stats_new(128)
        return st

    def run_device(self, pcm_dev: torch.Tensor) -> torch.Tensor:
        """pcm_dev int16 [n_clips, clip_samples] resident in HBM -> fp64[8] result (device)."""
        n_clips = pcm_dev.shape[0]
        flat = pcm_dev.reshape(-1)
        ex = self._plan(n_clips)
        emb = self.eng.vggish_forward(flat, ex)
        st = self._stats(emb)
        self.eng.stats_accumulate(emb, st.shift, st.acc)
        return self._finish(st)

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
        ex_chunk = self._plan(cpc)
        emb_all = torch.empty((n_clips * self.rows_per_clip, 128), dtype=torch.float16, device=self.dev)
        st = None
        for i, s in enumerate(range(0, n_clips, cpc)):
            b = i & 1
            c = min(cpc, n_clips - s)
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(self._free[b])
                self._staging[b][:c].copy_(pcm_host[s:s + c], non_blocking=True)
                self._ready[b].record(self._copy_stream)
            main.wait_event(self._ready[b])
            ex = ex_chunk if c == cpc else self._plan(c)
            out = emb_all[s * self.rows_per_clip:(s + c) * self.rows_per_clip]
            self.eng.vggish_forward(self._staging[b][:c].reshape(-1), ex, out)
            self._free[b].record(main)
        st = self._stats(emb_all)
        self.eng.stats_accumulate(emb_all, st.shift, st.acc)
        res = self._finish(st)
        return float(res[0].item())
