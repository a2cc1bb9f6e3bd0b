import sys

from .driver import cli

sys.exit(cli())
