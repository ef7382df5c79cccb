"""``fadtk.embeds`` command line (mirror of fadtk/embeds.py:5-27):

    python -m fadtk_b200.embeds -m MODEL [MODEL ...] -d DIR [DIR ...] [-w N] [-s sox]
"""
from argparse import ArgumentParser

from . import dist
from .fad import log
from .fad_batch import cache_embedding_files
from .model_loader import get_all_models


def main():
    models = {m.name: m for m in get_all_models()}

    agupa = ArgumentParser(prog="fadtk.embeds")
    agupa.add_argument('-m', '--models', type=str, choices=list(models.keys()), nargs='+', required=True)
    agupa.add_argument('-d', '--dirs', type=str, nargs='+', required=True)
    agupa.add_argument('-w', '--workers', type=int, default=8)
    agupa.add_argument('-s', '--sox-path', type=str, default='/usr/bin/sox')
    args = agupa.parse_args()

    dist.init_from_env()
    for model_name in args.models:
        model = models[model_name]
        for d in args.dirs:
            log.info(f"Caching embeddings for {d} using {model.name}")
            cache_embedding_files(d, model, workers=args.workers)


if __name__ == "__main__":
    main()
