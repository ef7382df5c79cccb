"""``python -m fadtk_b200.embeds -m MODEL [MODEL ...] -d DIR [DIR ...] [-w N] [-s sox]`` - the ``fadtk.embeds``
command line (fadtk/embeds.py); see cli.embeds_main."""
import sys

from .cli import embeds_main


def main():
    return embeds_main()


if __name__ == "__main__":
    sys.exit(main())
