"""``python -m fadtk_b200 <model> <baseline> <eval> [csv] [-w N] [-s sox] [--inf | --indiv]`` - the ``fadtk`` command
line (fadtk/__main__.py); see cli.score_main.  Under ``torchrun --nproc-per-node N`` the embedding stage is sharded
over N GPUs and rank 0 reports."""
import sys

from .cli import score_main


def main():
    return score_main()


if __name__ == "__main__":
    sys.exit(main())
