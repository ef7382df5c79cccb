"""Fit of the single-MUFU erf used by the GELU epilogue of fadtk_b200/csrc/conv_gemm.cuh.

Test infrastructure (offline tool): erf(z) = 1 - 2^(z q(z)) on z in [0, 4.3], q a degree-6 polynomial
fitted to log2(erfc(z)) / z by Lawson-reweighted least squares with weight erfc(z) ln 2 (so the
residual is the ABSOLUTE error of erf).  Prints the coefficients (lowest order first) and the
max error of an fp32 Horner evaluation.  Run: python -m oracle.fit_gelu
"""
import numpy as np
from scipy.special import erf, erfc

ZMAX, TERMS = 4.3, 7


def fit(terms: int = TERMS, zmax: float = ZMAX):
    z = np.linspace(0.0, zmax, 20001)
    target = np.log2(erfc(z))
    w = erfc(z) * np.log(2.0)
    V = np.stack([z ** (k + 1) for k in range(terms)], 1)
    lw = np.ones_like(z)
    for _ in range(60):
        c, *_ = np.linalg.lstsq(V * (w * lw)[:, None], target * w * lw, rcond=None)
        err = np.abs((V @ c - target) * w)
        lw = lw * (err / err.max() + 1e-3)
        lw /= lw.max()
    return c


def erf_fp32(z: np.ndarray, c: np.ndarray) -> np.ndarray:
    zf, cf = z.astype(np.float32), c.astype(np.float32)
    q = np.full_like(zf, cf[-1])
    for k in range(len(cf) - 2, -1, -1):
        q = (q * zf + cf[k]).astype(np.float32)
    return 1.0 - np.exp2((q * zf).astype(np.float32).astype(np.float64))


if __name__ == "__main__":
    c = fit()
    z = np.linspace(0.0, ZMAX, 200001)
    print("coefficients (z^0..):", ", ".join(f"{v:.8e}" for v in c))
    print("max |erf error| fp32:", np.abs(erf_fp32(z, c) - erf(z)).max())
