#!/usr/bin/env python3
"""Drop-in for the reference's 0-prepare-coex.py (same flags, same working directory and files); the permuted locus
sets, the global correlation list and -- unless -po -- the whole run go through the GPU library.
See coexpression_b200/prepare.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from coexpression_b200.prepare import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main(coexappdir=os.path.dirname(os.path.abspath(__file__))))
