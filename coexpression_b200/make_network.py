"""B200 drop-in for the reference's per-permutation driver (reference 1-make-network.py).

Same command line (`-mf <mapfile> -wd <working_files_dir> [-t N] [-r]`), same inputs
(settings.txt, expression table, feature BED, global correlation list, windowBed map file) and
the same eight JSON objects in `perm_results_store/` (1-make-network.py:523-540), so that
2-collate-results.py and runlocal.py's completion polling keep working.

`run_batch` is the B200-first entry: every map file of `permutation_store/` is reduced in ONE
process with the table resident on the GPU (the reference starts one process per permutation
and re-reads / re-ranks / re-marshals the table in each, runlocal.py:50-55).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

from . import hostio
from .api import CoexContext, Options
from .coexfunctions import (read_correlation_list, read_expression_file, read_snp_map, readfeaturecoordinates,
                            readsettings)
from .results import perm_objects


def options_from_settings(st):
    """settings.txt keys read at 1-make-network.py:35-49 (values arrive as floats / bools, Q9)."""
    seed = st.get("seedfile", "none")
    if isinstance(seed, str) and seed != "none" and os.path.exists(seed):      # the reference's own test, 1-make-network.py:85-86
        # the "obselete" -seed feature (1-make-network.py:82-97, 169-171, 484-505) adds the seed promoters to snp_map and
        # prunes them after the densities: it is out of scope here, and silently ignoring it
        # would write different perm_results_store objects than the reference does
        raise NotImplementedError("settings.txt names a seedfile (%r): seed-promoter pruning (reference 1-make-network.py:"
                                  "82-97,484-505) is not supported by the B200 drop-in; rerun 0-prepare without -seed" % (seed,))
    return Options(
        correlationmeasure=st.get("correlationmeasure", "Spearman"),
        anticorrelation=bool(st.get("anticorrelation", False)),
        noiter=bool(st.get("noiter", False)),
        useaverage=bool(st.get("useaverage", False)),
        nsp=float(st.get("nsp", 1.0)),
        pval_to_join=float(st.get("pval_to_join", 0.1)),
        selectionthreshold=float(st.get("selectionthreshold", 0)),
        selectionthreshold_is_j=bool(st.get("selectionthreshold_is_j", False)),
        soft_separation=int(float(st.get("soft_separation", 100000))))


def result_path(store, mapfile, label):
    parts = os.path.split(mapfile)[-1].replace(".bed", "").split(".")
    return os.path.join(store, ".".join(parts[:-1] + [label] + parts[-1:]))


class Session:
    """Expression table resident on one GPU + everything a permutation needs besides its hits."""

    def __init__(self, settings, device=0, verbose=False, cache=False):
        self.settings = settings
        self.verbose = verbose
        self.options = options_from_settings(settings)
        # cache=True: binary cache of the parsed table next to the text file (row N4, coexfunctions.read_expression_file)
        self.table, self.header = read_expression_file(settings["expression_file"], cache=cache)
        addresses, _, _ = readfeaturecoordinates(settings["feature_coordinates"])
        self.addresses = addresses
        glist = read_correlation_list(settings["correlationfile"])
        names = self.table.names
        chrom = np.array([addresses[nm][0] if nm in addresses else "" for nm in names])
        start = np.array([addresses[nm][1] if nm in addresses else 0 for nm in names], dtype=np.int64)
        end = np.array([addresses[nm][2] if nm in addresses else 0 for nm in names], dtype=np.int64)
        self.has_address = np.array([nm in addresses for nm in names])
        self.ctx = CoexContext(device)
        self.ctx.set_expression(self.table.values)
        self.ctx.prepare()
        self.ctx.set_features(chrom, start, end, names)
        self.ctx.set_global_list(glist)

    def close(self):
        self.ctx.close()

    def hits_of(self, mapfile):
        snp_map = read_snp_map(mapfile, self.table, self.verbose)
        rows = np.array([self.table.index[p] for p in snp_map], dtype=np.int32)
        if len(rows) > 1 and not self.has_address[rows].all():
            missing = [self.table.names[r] for r in rows if not self.has_address[r]]
            raise KeyError(missing[0])          # the reference fails in join_nearby (coexfunctions.py:285)
        return snp_map, rows

    def run(self, mapfiles, store, full_dump=True, max_scores_bytes=8 << 30, lean=False):
        """Reduce the map files in chunks whose score matrices fit `max_scores_bytes` and write each file's result objects
        (1-make-network.py:523-540).  full_dump: all eight objects; lean: the two P x P objects (`individualscores`,
        `indjoined`) only for the real data, permutation number 0 -- 2-collate-results.py reads nothing but `indmeasure`
        of a permuted set (:67-76) --, which also keeps the permuted sets' score blocks on the device."""
        os.makedirs(store, exist_ok=True)
        parsed = [self.hits_of(mf) for mf in mapfiles]
        scored = [full_dump and (not lean or permutation_number(mf) == 0) for mf in mapfiles]
        written = []
        i = 0
        while i < len(mapfiles):
            j, used = i, 0
            while j < len(mapfiles) and scored[j] == scored[i]:
                need = 8 * len(parsed[j][1]) ** 2
                if j > i and used + need > max_scores_bytes:
                    break
                used += need
                j += 1
            res = self.ctx.network_batch([parsed[k][1] for k in range(i, j)], self.options, want_scores=scored[i])
            for k in range(i, j):
                written += write_result_objects(store, mapfiles[k], res, k - i, self.table.names, parsed[k][0], scored[i])
            i = j
        return written


def permutation_number(mapfile):
    """<feature label>.<k>.bed -> k (0 = the real data, 0-prepare-coex.py:357,598-613); -1 if the name has no number."""
    try:
        return int(os.path.split(mapfile)[-1].replace(".bed", "").split(".")[-1])
    except ValueError:
        return -1


def write_result_objects(store, mapfile, res, k, names, snp_map, with_scores):
    """The files `<stem>.<object>.<k>` of one permutation, same names, same text as the reference's json.dump calls.  The
    six small objects go through json.dump; the two P x P ones are written straight from the score matrix by the host
    library (hostio.write_json_matrix: the same bytes json.dump would write, without building 420 000 Python floats)."""
    objs = perm_objects(res, k, names, with_scores=False)
    if with_scores:
        p = res.perm(k)
        hit_names = [names[r] for r in p["hit_rows"]]
        G = len(p["label"])
        if G == 0 or len(hit_names) < 2:
            objs["individualscores"] = {}
            objs["indjoined"] = {}
        else:
            S = res.score_matrix(k)
            w = np.asarray(p["winner"], dtype=np.int64)
            objs["individualscores"] = (S, hit_names)
            objs["indjoined"] = (S[np.ix_(w, w)], [hit_names[i] for i in p["label"]])
    objs["snp_map"] = snp_map
    written = []
    for label, var in objs.items():
        path = result_path(store, mapfile, label)
        if isinstance(var, tuple):
            hostio.write_json_matrix(path, var[0], var[1], skip_diag=True)
        else:
            with open(path, "w") as o:
                json.dump(var, o)
        written.append(path)
    return written


def run_batch(working_files_dir, mapfiles=None, device=0, verbose=False, full_dump=True, cache=False, lean=False):
    settings = readsettings(working_files_dir)
    perm_dir = os.path.join(working_files_dir, "permutation_store")
    if mapfiles is None:
        mapfiles = sorted((os.path.join(perm_dir, f) for f in os.listdir(perm_dir) if f.endswith(".bed")),
                          key=lambda p: int(p.split(".")[-2]))
    sess = Session(settings, device, verbose, cache)
    try:
        return sess.run(mapfiles, os.path.join(working_files_dir, "perm_results_store"), full_dump, lean=lean)
    finally:
        sess.close()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-mf", "--mapfile", type=str, default="null", help="filepath")
    ap.add_argument("-wd", "--working_files_dir", type=str, default="null", help="filepath")
    ap.add_argument("-t", "--numthreads", type=int, default=1, help="accepted and ignored, as in the reference (Q10)")
    ap.add_argument("-r", "--recordedges", action="store_true", default=False,
                    help="accepted; the reference crashes on it (Q11) so nothing is recorded")
    ap.add_argument("--all", action="store_true", help="B200 extension: reduce every map file of permutation_store/")
    ap.add_argument("--lean", action="store_true",
                    help="B200 extension: write `individualscores` / `indjoined` for the real data (permutation 0) only; "
                         "2-collate-results.py reads nothing else of a permuted set than `indmeasure`")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--cache", action="store_true", default=bool(os.environ.get("COEX_TABLE_CACHE")),
                    help="B200 extension: keep a binary cache of the parsed expression table next to the text file")
    args = ap.parse_args(argv)
    settings = readsettings(args.working_files_dir)
    verbose = bool(settings.get("verbose", False))
    if args.all:
        run_batch(args.working_files_dir, None, args.device, verbose, cache=args.cache, lean=args.lean)
    else:
        run_batch(args.working_files_dir, [args.mapfile], args.device, verbose, cache=args.cache, lean=args.lean)
    return 0


if __name__ == "__main__":
    sys.exit(main())
