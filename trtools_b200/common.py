"""stderr helpers (reference trtools/utils/common.py:7-36)."""
import sys


def WARNING(msg):
    sys.stderr.write(msg.strip() + "\n")


def MSG(msg, debug=False):
    if debug:
        sys.stderr.write(msg.strip() + "\n")
