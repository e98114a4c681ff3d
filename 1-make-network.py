#!/usr/bin/env python3
"""Drop-in for the reference's 1-make-network.py (same flags: -mf -wd [-t -r]); the arithmetic
runs on the GPU through coexpression_b200/csrc/coexpression_v2.so.  See coexpression_b200/make_network.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from coexpression_b200.make_network import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
