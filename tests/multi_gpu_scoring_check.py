"""Manual multi-GPU check (not collected by pytest): `--indiv` and `--inf` under torchrun must give the
same results as a single process.  Usage on a box with >= 2 GPUs:

    python tests/multi_gpu_scoring_check.py prepare /tmp/mg
    python -m fadtk_b200 vggish /tmp/mg/base.npz /tmp/mg/ev /tmp/mg/one.csv --indiv
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 -m fadtk_b200 vggish /tmp/mg/base.npz /tmp/mg/ev /tmp/mg/two.csv --indiv
    python tests/multi_gpu_scoring_check.py compare /tmp/mg
"""
import sys
from pathlib import Path

import numpy as np


def prepare(root: Path):
    rng = np.random.default_rng(3)
    d = 128
    mix = rng.standard_normal((d, d)) / np.sqrt(d)
    base = (rng.standard_normal((4000, d)) @ mix)
    np.savez(root / "base.npz", **{"vggish.mu": base.mean(0), "vggish.cov": np.cov(base, rowvar=False)})
    (root / "ev" / "embeddings" / "vggish").mkdir(parents=True, exist_ok=True)
    for i in range(37):
        rows = ((rng.standard_normal((20 + 7 * i, d)) @ mix) * (0.6 + 0.02 * i) + 0.01 * i).astype(np.float16)
        (root / "ev" / f"s{i:03d}.wav").write_bytes(b"")
        np.save(root / "ev" / "embeddings" / "vggish" / f"s{i:03d}.npy", rows)


def compare(root: Path):
    a = [ln.split(",") for ln in (root / "one.csv").read_text().splitlines()]
    b = [ln.split(",") for ln in (root / "two.csv").read_text().splitlines()]
    assert [r[0] for r in a] == [r[0] for r in b], "file order differs"
    sa, sb = np.array([float(r[1]) for r in a]), np.array([float(r[1]) for r in b])
    assert np.array_equal(sa, sb), np.abs(sa - sb).max()
    print(f"multi-GPU --indiv identical: {len(a)} rows")


if __name__ == "__main__":
    {"prepare": prepare, "compare": compare}[sys.argv[1]](Path(sys.argv[2]))
