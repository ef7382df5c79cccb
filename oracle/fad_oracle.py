"""numpy restatement of the reference's statistics and Frechet arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function names the reference
lines it follows; ``tests/test_oracle_golden.py`` pins every one of them against
outputs of the reference's own functions (tests/golden/, made by make_golden.py).

dtype behaviour that moves the answer at the 1e-5 level is kept on purpose:

* ``np.mean`` of an fp16 array returns fp16 (fad.py:48, utils.py:16);
* ``np.cov`` of an fp16 array is computed in fp64 (fad.py:48, utils.py:16);
* the returned Frechet value uses the eigen-decomposition route
  ``V sqrt(D) V^-1`` of the *non-symmetric* product C1 C2 (fad.py:91-92,108,119-120);
  ``scipy.linalg.sqrtm`` (fad.py:88) only feeds a log warning.
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np
from numpy.lib.scimath import sqrt as _complex_sqrt
from scipy import linalg as _la


class InfResult(NamedTuple):
    """Mirror of FADInfResults (fad.py:35-39)."""
    score: float
    slope: float
    r2: float
    points: list


def embd_statistics(rows: np.ndarray):
    """mean / covariance of an [n, d] embedding matrix.  Reference: fad.py:42-48."""
    if rows.shape[0] < 2:
        raise AssertionError(
            f"FAD requires at least two embedding window frames, you have {rows.shape}.")
    mu = rows.mean(axis=0)                  # dtype follows the input: fp16 stays fp16
    cov = np.cov(rows, rowvar=False)        # always promoted to fp64, ddof = 1
    return mu, cov


def file_partial(rows: np.ndarray):
    """Per-file sufficient statistics.  Reference: utils.py:13-16.

    Returns (mean [d] in the input dtype, scatter matrix [d, d] fp64, n).  A file
    with a single row gives an all-NaN scatter matrix (np.cov with ddof=1 divides
    by zero) - the reference has the same behaviour and it poisons the merge.
    """
    n = rows.shape[0]
    with np.errstate(all="ignore"):
        scatter = np.cov(rows, rowvar=False) * (n - 1)
    return rows.mean(axis=0), scatter, n


def merge_statistics(partials: Sequence[tuple]):
    """Sequential Chan merge of per-file partials.  Reference: utils.py:30-46."""
    first_mean = partials[0][0]
    d = first_mean.shape[-1]
    mu = np.zeros(d)
    scatter = np.zeros((d, d))
    n = 0
    for m_f, s_f, n_f in partials:
        delta = m_f - mu
        mu += n_f / (n + n_f) * delta
        scatter += s_f + np.outer(delta, delta) * n * n_f / (n + n_f)
        n += n_f
    if n < 2:
        return mu, np.zeros_like(scatter)
    return mu, scatter / (n - 1)


def online_statistics(per_file_rows: Sequence[np.ndarray]):
    """utils.py:19-46 applied to in-memory per-file embedding arrays."""
    if len(per_file_rows) == 0:
        raise AssertionError("No files provided")
    return merge_statistics([file_partial(r) for r in per_file_rows])


def trace_sqrt_product(cov1: np.ndarray, cov2: np.ndarray, eps: float = 1e-6) -> float:
    """tr sqrt(C1 C2) exactly as the reference evaluates it.  fad.py:88-108."""
    prod = cov1.dot(cov2)
    evals, evecs = _la.eig(prod)
    root = (evecs * _complex_sqrt(evals)) @ _la.inv(evecs)
    if not np.isfinite(root).all():
        # fad.py:94-99 - regularise both covariances and fall back to sqrtm
        bump = np.eye(cov1.shape[0]) * eps
        root = _la.sqrtm((cov1 + bump).dot(cov2 + bump))
    if np.iscomplexobj(root):
        if not np.allclose(np.diagonal(root).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(root.imag))))
        root = root.real
    return float(np.trace(root))


def frechet_distance(mu1, cov1, mu2, cov2, eps: float = 1e-6):
    """||mu1-mu2||^2 + tr C1 + tr C2 - 2 tr sqrt(C1 C2).  Reference: fad.py:51-120."""
    mu1 = np.atleast_1d(mu1)
    mu2 = np.atleast_1d(mu2)
    cov1 = np.atleast_2d(cov1)
    cov2 = np.atleast_2d(cov2)
    if mu1.shape != mu2.shape:
        raise AssertionError("Training and test mean vectors have different lengths")
    if cov1.shape != cov2.shape:
        raise AssertionError("Training and test covariances have different dimensions")
    diff = mu1 - mu2                        # fp16 - fp16 stays fp16 (fad.py:83)
    return diff.dot(diff) + np.trace(cov1) + np.trace(cov2) \
        - 2 * trace_sqrt_product(cov1, cov2, eps)


def inf_sample_sizes(n_rows: int, steps: int = 25, min_n: int = 500):
    """fad.py:325-328."""
    return [int(n) for n in np.linspace(min_n, n_rows, steps)]


def score_inf(mu_base, cov_base, rows: np.ndarray, steps: int = 25, min_n: int = 500,
              rng=np.random) -> InfResult:
    """FAD-infinity extrapolation.  Reference: fad.py:304-351.

    Draws the bootstrap indices from ``rng.choice`` exactly like fad.py:333 (global
    numpy RNG, with replacement), so ``np.random.seed(k)`` reproduces the reference.
    """
    sizes = inf_sample_sizes(len(rows), steps, min_n)
    points = []
    for n in sizes:
        pick = rng.choice(rows.shape[0], size=n, replace=True)
        mu_e, cov_e = embd_statistics(rows[pick])
        points.append([n, frechet_distance(mu_base, cov_base, mu_e, cov_e)])
    ys = np.array(points)[:, 1]
    xs = 1 / np.array(sizes)
    slope, intercept = np.polyfit(xs, ys, 1)
    resid = ys - (slope * xs + intercept)
    r2 = 1 - np.sum(resid ** 2) / np.sum((ys - np.mean(ys)) ** 2)
    return InfResult(score=intercept, slope=slope, r2=r2, points=points)


def score_individual(mu_base, cov_base, named_rows: Sequence[tuple]):
    """Per-song scores sorted by |score|.  Reference: fad.py:353-395.

    ``named_rows`` is a sequence of (name, rows).  Songs whose statistics raise are
    dropped, as the reference swallows the exception (fad.py:380-391).
    """
    out = []
    for name, rows in named_rows:
        try:
            mu_e, cov_e = embd_statistics(rows)
            out.append((name, frechet_distance(mu_base, cov_base, mu_e, cov_e)))
        except Exception:
            continue
    return sorted(out, key=lambda p: np.abs(p[1]))
