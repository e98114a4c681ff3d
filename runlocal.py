#!/usr/bin/env python3
"""Stand-in for the reference's runlocal.py (`runlocal.py -wd <working_files_dir>`, started by 0-prepare-coex.py:237-245): where the
reference keeps up to four 1-make-network.py processes alive and polls perm_results_store every 30 s (runlocal.py:36-98), this
reduces every map file of permutation_store/ in ONE process with the table resident on the GPU and then collates."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from coexpression_b200 import collate, make_network  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-wd", "--working_files_dir", type=str, default="null", help="filepath")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--lean", action="store_true", help="B200 extension: the two P x P result objects for the real data only")
    a = ap.parse_args()
    make_network.run_batch(a.working_files_dir, None, a.device, lean=a.lean)
    collate.collate(a.working_files_dir)
