"""Acceptance run against the reference's published per-song scores (mirror of ``python -m fadtk.test``,
fadtk/test/__main__.py): embed the sample clips, score every clip and the whole set against ``fma_pop``, and
require the largest deviation from ``samples_FAD_scores.csv`` to stay below 5 % of the mean score per model.

It needs what the reference's own run needs and this repository cannot ship: the sample audio and the score table
(``--samples``, ``--scores``; the reference keeps them in fadtk/test/), the ``fma_pop`` statistics
(``FADTK_STATS_DIR``) and real model checkpoints (``FADTK_*_CKPT``, see weights*.py) - with the seeded synthetic
weights the embeddings are not the published models'.

    python -m fadtk_b200.test --samples <dir> --scores <samples_FAD_scores.csv> [--models vggish clap-laion-audio] [--out <dir>]
"""
from __future__ import annotations

import argparse
import csv
import sys
import traceback
from pathlib import Path

import numpy as np

TOLERANCE_PERCENT = 5.0            # fadtk/test/__main__.py:95


def reference_scores(table: Path) -> dict:
    """{model_name_with_underscores: {"samples/<stem>": score}} from the FAD_<model>_fma_pop columns."""
    out: dict = {}
    with open(table, newline="") as f:
        for row in csv.DictReader(f):
            for col, v in row.items():
                if col.startswith("FAD_") and col.endswith("_fma_pop") and v not in ("", None):
                    out.setdefault(col[len("FAD_"):-len("_fma_pop")], {})[row["song_id"]] = float(v)
    return out


def song_id(path: str) -> str:
    """'<...>/samples/mg-1634.opus' (either slash) -> 'samples/mg-1634' - the key the score table uses."""
    parts = path.replace("\\", "/").split("/")
    return "/".join(parts[-2:]).split(".")[0]


def compare(ours: dict, theirs: dict) -> dict:
    """Deviation statistics of one model: ours / theirs map song ids to scores (ours may hold a subset)."""
    ids = sorted(ours)
    a = np.array([ours[i] for i in ids], dtype=np.float64)
    b = np.array([theirs[i] for i in ids], dtype=np.float64)
    mad = float(np.abs(a - b).max())
    mean = float(a.mean())
    pct = mad / mean * 100.0
    return {"mse": float(((a - b) ** 2).mean()), "max_abs_diff": mad, "mean": mean, "mad%": pct, "pass": bool(pct < TOLERANCE_PERCENT)}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--samples", type=Path, required=True)
    ap.add_argument("--scores", type=Path, required=True)
    ap.add_argument("--models", nargs="*", default=None)
    ap.add_argument("--out", type=Path, default=Path("fad_scores"))
    ap.add_argument("-w", "--workers", type=int, default=8)
    args = ap.parse_args(argv)

    from ..fad import FrechetAudioDistance, log
    from ..fad_batch import cache_embedding_files
    from ..model_loader import get_all_models

    ref = reference_scores(args.scores)
    print("Models with reference data:", sorted(ref))
    table = []
    for model in get_all_models():
        key = model.name.replace("-", "_")
        if key not in ref or (args.models and model.name not in args.models):
            continue
        if model.name.startswith("MERT") and model.name[-1] not in "148M":     # the reference's own layer subset (:29-31)
            continue
        log.info(f"Computing FAD scores for {model.name}")
        out_csv = args.out / f"{model.name}.csv"
        try:
            if not out_csv.is_file():
                cache_embedding_files(args.samples, model, workers=args.workers)
                fad = FrechetAudioDistance(model, audio_load_worker=args.workers, load_model=False)
                fad.score_individual("fma_pop", args.samples, out_csv)
                whole = fad.score("fma_pop", args.samples)
                with open(out_csv, "a") as f:
                    f.write(f"\n/samples/all,{whole}")
        except Exception as e:                                 # noqa: BLE001 - report and fail like the reference (:45-58)
            traceback.print_exc()
            log.error(f"Error when computing FAD scores for {model.name}: {e}")
            return 1
        with open(out_csv, newline="") as f:
            ours = {song_id(r[0]): float(r[1]) for r in csv.reader(f) if len(r) == 2}
        row = {"model": model.name, **compare(ours, ref[key])}
        table.append(row)
        log.info(row)
    if not table:
        print("no model with reference data was run", file=sys.stderr)
        return 1
    with open(args.out / "comparison.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(table[0]))
        w.writeheader()
        w.writerows(table)
    if not all(r["pass"] for r in table):
        log.error("Some models failed the test")
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
