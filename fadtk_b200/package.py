"""Pack the statistics of one dataset directory into a baseline ``.npz`` (mirror of ``python -m fadtk.package``,
fadtk/package.py): keys ``<model>.mu`` / ``<model>.cov``, the file ``FrechetAudioDistance.load_stats`` accepts as a
baseline path or - dropped into ``$FADTK_STATS_DIR`` - as a baseline *name* like the reference's ``fma_pop``
(fad.py:249-266).

    python -m fadtk_b200.package <directory> <out.npz> [-m vggish clap-laion-audio ...] [-w 8]

Without ``-m`` every model of the registry is embedded, as the reference does (143 forward passes over the
directory); models whose statistics are already cached under ``<directory>/stats/<model>/`` cost nothing.
"""
from __future__ import annotations

import sys
from argparse import ArgumentParser
from pathlib import Path

import numpy as np


def pack_statistics(directory, out, models, workers: int = 8, embed: bool = True) -> Path:
    """``models``: ModelLoader instances.  Embeds what is missing (``embed``), then writes the ``.npz``."""
    from .fad import FrechetAudioDistance
    from .fad_batch import cache_embedding_files

    data = {}
    for model in models:
        if embed and not (Path(directory) / "stats" / model.name).exists():
            cache_embedding_files(directory, model, workers=workers)
        fad = FrechetAudioDistance(model, audio_load_worker=workers, load_model=False)
        mu, cov = fad.load_stats(Path(directory))
        data[f"{model.name}.mu"] = mu
        data[f"{model.name}.cov"] = cov
    out = Path(out)
    out.parent.mkdir(parents=True, exist_ok=True)
    np.savez(out, **data)
    return out if out.suffix == ".npz" else out.with_name(out.name + ".npz")     # np.savez appends the suffix


def main(argv=None) -> int:
    from .model_loader import get_all_models

    registry = {m.name: m for m in get_all_models()}
    ap = ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("directory", type=str)
    ap.add_argument("out", type=str)
    ap.add_argument("-m", "--models", nargs="*", choices=list(registry), default=None)
    ap.add_argument("-w", "--workers", type=int, default=8)
    ap.add_argument("-s", "--sox-path", type=str, default="/usr/bin/sox")
    ap.add_argument("-y", "--yes", action="store_true", help="do not ask when the output name lacks .npz")
    args = ap.parse_args(argv)
    if Path(args.out).suffix != ".npz" and not args.yes:
        print("The output file you specified is not a npz file, are you sure? (y/N)")
        if input().lower() != "y":
            return 1
    chosen = [registry[n] for n in (args.models or registry)]
    if not args.models:                                        # the default (every model) skips entries without a forward pass
        from .model_loader import UnbuiltModel
        skipped = [m.name for m in chosen if isinstance(m, UnbuiltModel)]
        if skipped:
            print(f"skipping {', '.join(skipped)}: not built in this package (name them with -m to force)")
        chosen = [m for m in chosen if not isinstance(m, UnbuiltModel)]
    path = pack_statistics(args.directory, args.out, chosen, workers=args.workers)
    print(f"statistics of {len(chosen)} model(s) written to {path}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
