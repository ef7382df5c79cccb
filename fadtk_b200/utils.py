"""Statistics + cache-path helpers (mirror of fadtk/utils.py).

``calculate_embd_statistics_online`` keeps the reference's name and return value but not its
mechanism: instead of one process per file pickling a d x d fp64 scatter matrix back to a
sequential Chan merge (utils.py:13-46), all rows go through one shifted E^T E tensor-core
contraction on the GPU (csrc/stats.cuh) whose packed fp64 result is additive - across batches
and, with one all-reduce, across GPUs.
"""
from __future__ import annotations

import subprocess
from pathlib import Path
from typing import Iterable, Union

import numpy as np
import torch

PathLike = Union[str, Path]

_ROWS_PER_UPLOAD = 1 << 20


def _split_hi_lo(x: torch.Tensor) -> torch.Tensor:
    """fp32/fp64 rows -> fp16 [n, 2d] = [hi | lo] with hi + lo == x to ~2^-22 relative."""
    hi = x.to(torch.float16)
    lo = (x - hi.to(x.dtype)).to(torch.float16)
    return torch.cat([hi, lo], dim=1).contiguous()


class DeviceStatistics:
    """Running (n, sum, outer-product-sum) of embedding rows on one GPU.

    ``add`` accepts numpy or torch arrays of shape [n, d].  fp16 rows (what the reference caches,
    model_loader.py:47-48) go straight to the tensor-core kernel; wider dtypes are split into
    fp16 hi/lo halves and contracted as a [n, 2d] matrix so nothing is rounded away.
    """

    def __init__(self, d: int, engine=None, reduce_ranks: bool = False):
        """``reduce_ranks``: this accumulator will be summed over the ranks (``allreduce``); every rank must then shift
        by the SAME vector, so the first ``add`` broadcasts rank 0's (a collective call)."""
        from . import _native
        self.eng = engine or _native.engine()
        self.d = d
        self.reduce_ranks = reduce_ranks
        self.wide = None
        self.shift = None
        self.acc = None

    def _setup(self, first: torch.Tensor):
        self.wide = first.dtype != torch.float16
        dd = 2 * self.d if self.wide else self.d
        head = first[: min(first.shape[0], 4096)].to(self.eng.torch_device)
        if self.wide:
            s = torch.zeros(dd, dtype=torch.float16, device=self.eng.torch_device)
            s[: self.d] = head.double().mean(0).to(torch.float16)
            self.shift = s
        else:
            self.shift = head.float().mean(0).to(torch.float16)
        if self.reduce_ranks:
            from . import dist
            if dist.is_distributed():
                s32 = self.shift.float()
                torch.distributed.broadcast(s32, src=0)
                self.shift = s32.to(torch.float16)
        self.acc = self.eng.stats_new(dd)

    def add(self, rows):
        t = torch.from_numpy(np.ascontiguousarray(rows)) if isinstance(rows, np.ndarray) else rows
        if t.shape[0] == 0:
            return
        if self.acc is None:
            self._setup(t)
        for s in range(0, t.shape[0], _ROWS_PER_UPLOAD):
            part = t[s:s + _ROWS_PER_UPLOAD]
            if not part.is_cuda:
                # pinned source (what the native embedders hand back): asynchronous DMA.  Pageable source: one staged copy
                # by the driver - cheaper than pinning a copy first (cudaHostAlloc + memcpy + DMA)
                part = part.to(self.eng.torch_device, non_blocking=part.is_pinned())
            part = _split_hi_lo(part) if self.wide else part.contiguous()
            self.eng.stats_accumulate(part, self.shift, self.acc)

    def add_gather(self, emb_dev: torch.Tensor, idx_dev: torch.Tensor):
        if self.acc is None:
            self._setup(emb_dev)
        assert not self.wide
        self.eng.stats_accumulate_gather(emb_dev, idx_dev, self.shift, self.acc)

    def allreduce(self):
        """Sum the packed accumulator over all ranks (NCCL over NVLink); no-op single-process."""
        from . import dist
        dist.enable_native_allreduce(self.eng)               # the C ABI's own NCCL communicator when running on GPUs
        dist.allreduce_sum_(self.acc)

    def count(self) -> int:
        return 0 if self.acc is None else int(self.acc[0].item())

    def finalize(self):
        """-> (mu fp64 [d], cov fp64 [d, d]) cuda tensors; cov is zero when n < 2 (utils.py:42-43)."""
        if self.acc is None:
            raise AssertionError("No files provided")
        if not self.wide:
            return self.eng.stats_finalize(self.acc, self.shift, self.d)
        d = self.d
        mu2, cov2 = self.eng.stats_finalize(self.acc, self.shift, 2 * d)
        mu = mu2[:d] + mu2[d:]
        cov = cov2[:d, :d] + cov2[:d, d:] + cov2[d:, :d] + cov2[d:, d:]
        return mu.contiguous(), cov.contiguous()


def statistics_of_arrays(arrays: Iterable[np.ndarray], d: int | None = None, reduce_ranks: bool = False):
    """mean / covariance of the concatenation of ``arrays`` -> numpy fp64 (mu [d], cov [d, d])."""
    st = None
    for a in arrays:
        if st is None:
            st = DeviceStatistics(d or a.shape[-1], reduce_ranks=reduce_ranks)
        st.add(a)
    if st is None:
        raise AssertionError("No files provided")
    if reduce_ranks:
        st.allreduce()
    mu, cov = st.finalize()
    return mu.cpu().numpy(), cov.cpu().numpy()


def _weighted_scatter(means: np.ndarray, counts: np.ndarray, centre: np.ndarray) -> np.ndarray:
    """sum_f n_f (m_f - c)(m_f - c)^T"""
    dm = means - centre[None, :]
    return (dm * counts[:, None]).T @ dm


def mirror_file_mean_rounding(mu_exact, cov_exact, n, file_means16, file_means64, counts):
    """Reproduce what the reference's online merge computes (fadtk/utils.py:13-46).

    ``_process_file`` returns ``np.mean`` of an fp16 array - an fp16 value - while its scatter
    matrix comes from ``np.cov`` with its own fp64 mean.  The Chan merge is algebraically exact for
    whatever means it is fed, so the reference's result is
        mu  = sum_f n_f m16_f / n
        S   = sum_f S_f + sum_f n_f (m16_f - mu)(m16_f - mu)^T
    whereas the exact scatter is  sum_f S_f + sum_f n_f (m_f - mu_exact)(m_f - mu_exact)^T.
    The difference is a rank-F correction built from the per-file means only.
    """
    counts = np.asarray(counts, dtype=np.float64)
    m16 = np.asarray(file_means16, dtype=np.float64)
    m64 = np.asarray(file_means64, dtype=np.float64)
    mu_ref = (m16 * counts[:, None]).sum(0) / n
    if n < 2:
        return mu_ref, np.zeros_like(cov_exact)
    s_ref = cov_exact * (n - 1) - _weighted_scatter(m64, counts, mu_exact) + _weighted_scatter(m16, counts, mu_ref)
    return mu_ref, s_ref / (n - 1)


def calculate_embd_statistics_online(files: list[PathLike]) -> tuple[np.ndarray, np.ndarray]:
    """Mean and covariance of the embeddings stored in ``files`` (fadtk/utils.py:19-46).

    :param files: npy files holding ndarrays of shape (n_frames, n_features)

    All rows go through one exact shifted Gram contraction on the GPU; the reference's habit of
    rounding every per-file mean to fp16 before merging (it moves FAD by ~1e-4 on small sets) is
    then mirrored from the per-file means.  A file with exactly ONE frame makes the reference's
    covariance all-NaN (``np.cov`` of one row with ddof = 1 is NaN, ``* (n - 1)`` keeps it, the merge
    spreads it - utils.py:16, 36-45; BASELINE config 1's 1-s VGGish clips hit this) and the score then
    fails in ``calc_frechet_distance``; that is mirrored: mu as the reference computes it, cov = NaN.
    ``FADTK_SINGLE_FRAME_FILES=keep`` instead lets such a file contribute its row (an extension).
    """
    assert len(files) > 0, "No files provided"
    from . import _io_native
    files = list(files)
    st = None
    m_in, m64, counts = [], [], []
    quirk = True
    for chunk in _io_native.plan_embedding_chunks(files, _BYTES_PER_READ):
        # one native batched read per chunk of files (libfadtk_io.so, at most ~1 GB), one Gram accumulation per chunk
        flat, off = _io_native.load_embedding_files(chunk)
        if st is None:
            st = DeviceStatistics(flat.shape[-1])
        st.add(flat)
        quirk = quirk and flat.dtype == np.float16
        a, b, c = per_file_means(flat, off)
        m_in.append(a), m64.append(b), counts.append(c)
    mu, cov = st.finalize()
    mu, cov = mu.cpu().numpy(), cov.cpu().numpy()
    if not quirk:
        return mu, cov
    counts = np.concatenate(counts)
    mu_ref, cov_ref = mirror_file_mean_rounding(mu, cov, float(counts.sum()), np.concatenate(m_in), np.concatenate(m64), counts)
    return mu_ref, poison_single_frame_files(cov_ref, counts)


def poison_single_frame_files(cov: np.ndarray, counts) -> np.ndarray:
    """The reference's result when some file holds a single frame: an all-NaN covariance (utils.py:16)."""
    import os
    n_single = int((np.asarray(counts) == 1).sum())
    if n_single == 0 or os.environ.get("FADTK_SINGLE_FRAME_FILES", "") == "keep":
        return cov
    import logging
    logging.getLogger("fadtk_b200").warning(
        f"{n_single} embedding file(s) hold a single frame: the reference's per-file np.cov (fadtk/utils.py:16) is NaN "
        "for them and its merge makes the whole covariance NaN - mirrored. Use longer clips, concatenate the embeddings "
        "(calc_embd_statistics), or set FADTK_SINGLE_FRAME_FILES=keep to let such files contribute their row.")
    return np.full_like(cov, np.nan)


_BYTES_PER_READ = 1 << 30


def per_file_means(flat: np.ndarray, off: np.ndarray):
    """Per-file means of a ragged concatenation (file i = flat[off[i]:off[i+1]]), empty files dropped:
    -> (np.mean(file, axis=0) in the array's own dtype - fp16 in, fp16 out, as _process_file computes it
    (fadtk/utils.py:14) -, the same mean in fp64, row counts).  Files of equal length are reduced together;
    the result is bit-identical to calling np.mean file by file."""
    rows = np.diff(off)
    keep = np.nonzero(rows > 0)[0]
    m_in = np.empty((len(keep), flat.shape[1]), dtype=flat.dtype)
    m64 = np.empty((len(keep), flat.shape[1]), dtype=np.float64)
    pos = {int(i): k for k, i in enumerate(keep)}
    for r in np.unique(rows[keep]):
        idx = keep[rows[keep] == r]
        starts = off[idx]
        if len(idx) > 1 and np.all(np.diff(starts) == r):         # contiguous run of equal-length files: a plain view
            block = flat[starts[0]:starts[0] + len(idx) * r].reshape(len(idx), r, -1)
        else:
            block = flat[(starts[:, None] + np.arange(r)[None, :]).reshape(-1)].reshape(len(idx), r, -1)
        k = [pos[int(i)] for i in idx]
        with np.errstate(all="ignore"):
            m_in[k] = np.mean(block, axis=1)
        m64[k] = block.mean(axis=1, dtype=np.float64)
    return m_in, m64, rows[keep].astype(np.int64)


def pack_statistics_numpy(rows: np.ndarray, shift: np.ndarray) -> np.ndarray:
    """Host definition of the packed accumulator the GPU kernels produce (include/fadtk_b200.h):
    [n | sum(x - shift) exact | sum y y^T | sum y] with y = x - shift carried as an fp16 hi/lo
    pair.  Used to document and test the wire format of the cross-GPU all-reduce; the product
    path never calls it."""
    d = rows.shape[1]
    x = rows.astype(np.float16)
    y32 = x.astype(np.float32) - shift.astype(np.float32)          # exact
    hi = y32.astype(np.float16)
    lo = (y32 - hi.astype(np.float32)).astype(np.float16)
    y = hi.astype(np.float64) + lo.astype(np.float64)
    exact = y32.astype(np.float64)
    acc = np.zeros(1 + 2 * d + d * d)
    acc[0] = rows.shape[0]
    acc[1:1 + d] = exact.sum(0)
    acc[1 + d:1 + d + d * d] = (y.T @ y).ravel()
    acc[1 + d + d * d:] = y.sum(0)
    return acc


def finalize_packed_numpy(acc: np.ndarray, shift: np.ndarray):
    """mu, cov from a packed accumulator (same algebra as csrc/stats.cuh stats_finalize_kernel)."""
    d = shift.shape[0]
    n = acc[0]
    sum_x, outer, sum_y = acc[1:1 + d], acc[1 + d:1 + d + d * d].reshape(d, d), acc[1 + d + d * d:]
    mu = shift.astype(np.float64) + (sum_x / n if n > 0 else 0.0)
    cov = np.zeros((d, d)) if n < 2 else (outer - np.outer(sum_y, sum_y) / n) / (n - 1)
    return mu, cov


def find_sox_formats(sox_path: str) -> list[str]:
    """File formats supported by SoX (fadtk/utils.py:49-57); empty when SoX is absent."""
    try:
        out = subprocess.check_output((sox_path, "-h")).decode()
        head = "AUDIO FILE FORMATS: "
        i = out.index(head) + len(head)
        return out[i:out.index("\n", i)].split()
    except Exception:
        return []


def get_cache_embedding_path(model: str, audio_dir: PathLike) -> Path:
    """<dir>/embeddings/<model>/<stem>.npy for an audio file (fadtk/utils.py:60-68)."""
    audio_dir = Path(audio_dir)
    return audio_dir.parent / "embeddings" / model / audio_dir.with_suffix(".npy").name
