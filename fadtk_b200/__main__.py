"""``fadtk`` command line (mirror of fadtk/__main__.py:9-70): same positionals, flags and outputs.

    python -m fadtk_b200 <model> <baseline> <eval> [csv] [-w N] [-s sox] [--inf | --indiv]

Run under ``torchrun --nproc-per-node N`` to shard the embedding stage over N GPUs; scoring
and CSV output happen on rank 0.
"""
import time
from argparse import ArgumentParser
from pathlib import Path

from . import dist
from .fad import FrechetAudioDistance, log
from .fad_batch import cache_embedding_files
from .model_loader import get_all_models


def main():
    models = {m.name: m for m in get_all_models()}

    agupa = ArgumentParser(prog="fadtk")
    agupa.add_argument('model', type=str, choices=list(models.keys()), help="The embedding model to use")
    agupa.add_argument('baseline', type=str, help="The baseline dataset")
    agupa.add_argument('eval', type=str, help="The directory to evaluate against")
    agupa.add_argument('csv', type=str, nargs='?',
                       help="The CSV file to append results to. "
                            "If this argument is not supplied, single-value results will be printed to stdout, "
                            "and for --indiv, the results will be saved to 'fad-individual-results.csv'")
    agupa.add_argument('-w', '--workers', type=int, default=8)
    agupa.add_argument('-s', '--sox-path', type=str, default='/usr/bin/sox')
    agupa.add_argument('--inf', action='store_true', help="Use FAD-inf extrapolation")
    agupa.add_argument('--indiv', action='store_true',
                       help="Calculate FAD for individual songs and store the results in the given file")
    args = agupa.parse_args()

    dist.init_from_env()
    model = models[args.model]
    baseline, eval = args.baseline, args.eval

    # 1. embeddings for directory arguments
    for d in [baseline, eval]:
        if Path(d).is_dir():
            cache_embedding_files(d, model, workers=args.workers)
    if dist.rank() != 0 and not (args.inf or args.indiv):     # plain FAD: rank 0 scores; --inf / --indiv shard their work
        dist.shutdown()
        return

    # 2. FAD
    fad = FrechetAudioDistance(model, audio_load_worker=args.workers, load_model=False)
    if args.inf:
        assert Path(eval).is_dir(), "FAD-inf requires a directory as the evaluation dataset"
        score = fad.score_inf(baseline, sorted(Path(eval).glob('*.*')))
        if dist.rank() != 0:
            dist.shutdown()
            return
        print("FAD-inf Information:", score)
        score, inf_r2 = score.score, score.r2
    elif args.indiv:
        assert Path(eval).is_dir(), "Individual FAD requires a directory as the evaluation dataset"
        csv_path = Path(args.csv or 'fad-individual-results.csv')
        fad.score_individual(baseline, eval, csv_path)
        if dist.rank() == 0:
            log.info(f"Individual FAD scores saved to {csv_path}")
        dist.shutdown()
        exit(0)
    else:
        score = fad.score(baseline, eval)
        inf_r2 = None

    # 3. results
    log.info("FAD computed.")
    if args.csv:
        Path(args.csv).parent.mkdir(parents=True, exist_ok=True)
        if not Path(args.csv).is_file():
            Path(args.csv).write_text('model,baseline,eval,score,inf_r2,time\n')
        with open(args.csv, 'a') as f:
            f.write(f'{model.name},{baseline},{eval},{score},{inf_r2},{time.time()}\n')
        log.info(f"FAD score appended to {args.csv}")

    log.info(f"The FAD {model.name} score between {baseline} and {eval} is: {score}")
    dist.shutdown()


if __name__ == "__main__":
    main()
