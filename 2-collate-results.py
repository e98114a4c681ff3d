#!/usr/bin/env python3
"""Drop-in for the arithmetic and main table of the reference's 2-collate-results.py (same flag: -wd).
See coexpression_b200/collate.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from coexpression_b200.collate import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
