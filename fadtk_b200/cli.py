"""Argument parsing and the three run modes behind the ``fadtk`` / ``fadtk.embeds`` command lines.

The interface is the reference's (fadtk/__main__.py:17-30, fadtk/embeds.py:14-19) - positionals ``model baseline
eval [csv]``, ``-w/--workers``, ``-s/--sox-path``, ``--inf``, ``--indiv``; ``-m/--models``, ``-d/--dirs`` - so scripts
written for ``fadtk`` run unchanged; the implementation is table driven and multi-GPU aware (torchrun: every rank
embeds its shard, rank 0 reports).
"""
from __future__ import annotations

import time
from argparse import ArgumentParser
from pathlib import Path

from . import dist

_COMMON = (
    (("-w", "--workers"), dict(type=int, default=8, help="host threads for file I/O and audio conversion")),
    (("-s", "--sox-path"), dict(type=str, default="/usr/bin/sox", help="accepted for compatibility; conversion runs on the GPU")),
)
_SCORE_ARGS = (
    (("model",), dict(type=str, help="embedding model (a registry name)")),
    (("baseline",), dict(type=str, help="baseline set: a directory, a statistics .npz, or a built-in statistics name")),
    (("eval",), dict(type=str, help="evaluation set: a directory or a statistics .npz")),
    (("csv",), dict(type=str, nargs="?", help="append the result row here; with --indiv: where the per-song table goes "
                                              "(default fad-individual-results.csv)")),
    (("--inf",), dict(action="store_true", help="FAD-inf: extrapolate the score to an infinite evaluation set")),
    (("--indiv",), dict(action="store_true", help="score every evaluation file on its own against the baseline")),
)
_EMBED_ARGS = (
    (("-m", "--models"), dict(type=str, nargs="+", required=True, help="registry names")),
    (("-d", "--dirs"), dict(type=str, nargs="+", required=True, help="audio directories")),
)
CSV_HEADER = "model,baseline,eval,score,inf_r2,time\n"       # fadtk/__main__.py:64


def _parser(prog: str, table, registry) -> ArgumentParser:
    ap = ArgumentParser(prog=prog)
    for flags, kw in table + _COMMON:
        kw = dict(kw)
        if flags[-1] in ("model", "--models"):
            kw["choices"] = list(registry)
        ap.add_argument(*flags, **kw)
    return ap


def _registry():
    from .model_loader import get_all_models
    return {m.name: m for m in get_all_models()}


def _embed_directories(model, paths, workers):
    from .fad_batch import cache_embedding_files
    for p in paths:
        if Path(p).is_dir():
            cache_embedding_files(p, model, workers=workers)


def _append_row(csv_path: str, row) -> None:
    out = Path(csv_path)
    out.parent.mkdir(parents=True, exist_ok=True)
    if not out.is_file():
        out.write_text(CSV_HEADER)
    with open(out, "a") as f:
        f.write(",".join(str(v) for v in row) + "\n")


def score_main(argv=None) -> int:
    """``python -m fadtk_b200 model baseline eval [csv] [--inf | --indiv]``"""
    from .fad import FrechetAudioDistance, log
    registry = _registry()
    args = _parser("fadtk", _SCORE_ARGS, registry).parse_args(argv)
    dist.init_from_env()
    model = registry[args.model]
    _embed_directories(model, (args.baseline, args.eval), args.workers)
    per_rank_work = args.inf or args.indiv                     # those two shard their own work over the ranks
    if dist.rank() != 0 and not per_rank_work:
        dist.shutdown()
        return 0

    fad = FrechetAudioDistance(model, audio_load_worker=args.workers, load_model=False)
    r2 = None
    if args.inf:                                               # --inf wins when both are given (fadtk/__main__.py:45-50)
        assert Path(args.eval).is_dir(), "FAD-inf requires a directory as the evaluation dataset"
        result = fad.score_inf(args.baseline, sorted(Path(args.eval).glob("*.*")))
        if dist.rank() != 0:
            dist.shutdown()
            return 0
        print("FAD-inf Information:", result)
        score, r2 = result.score, result.r2
    elif args.indiv:
        assert Path(args.eval).is_dir(), "Individual FAD requires a directory as the evaluation dataset"
        table = Path(args.csv or "fad-individual-results.csv")
        fad.score_individual(args.baseline, args.eval, table)
        if dist.rank() == 0:
            log.info(f"Individual FAD scores saved to {table}")
        dist.shutdown()
        return 0
    else:
        score = fad.score(args.baseline, args.eval)

    log.info("FAD computed.")
    if args.csv:
        _append_row(args.csv, (model.name, args.baseline, args.eval, score, r2, time.time()))
        log.info(f"FAD score appended to {args.csv}")
    log.info(f"The FAD {model.name} score between {args.baseline} and {args.eval} is: {score}")
    dist.shutdown()
    return 0


def embeds_main(argv=None) -> int:
    """``python -m fadtk_b200.embeds -m MODEL [...] -d DIR [...]``"""
    from .fad import log
    registry = _registry()
    args = _parser("fadtk.embeds", _EMBED_ARGS, registry).parse_args(argv)
    dist.init_from_env()
    for name in args.models:
        for d in args.dirs:
            log.info(f"Caching embeddings for {d} using {name}")
            _embed_directories(registry[name], (d,), args.workers)
    dist.shutdown()
    return 0
