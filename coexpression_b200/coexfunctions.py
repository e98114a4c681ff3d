"""Python-3 mirror of the reference's shared helpers that sit on the hot path's boundary
(reference coexfunctions.py).  Same names, argument meaning and file formats; arithmetic that
the reference does per pair in Python / scipy is NOT here -- it runs on the GPU through
coexpression_b200.api.  Nothing in this module imports the oracle.

  readsettings ............ coexfunctions.py:49-74   settings.txt -> dict (float / bool coercion)
  read_expression_file .... coexfunctions.py:201-233 (drop NA columns, first duplicate label wins; optional binary cache, row N4)
  readfeaturecoordinates .. coexfunctions.py:235-274
  read_correlation_list ... 1-make-network.py:76-79
  read_snp_map ............ 1-make-network.py:131-177
  empiricalp .............. coexfunctions.py:379-383
  fdr_correction .......... coexfunctions.py:333-358
"""
from __future__ import annotations

import json
import os
from bisect import bisect_left

import numpy as np

version = "v1.3-b200"


def getversion():
    return version


def str2bool(v):
    if v.lower() == "false":
        return False
    if v.lower() == "true":
        return True
    return v


def readsettings(wfd):
    setfile = os.path.join(wfd, "settings.txt")
    if not os.path.exists(setfile):
        print("SERIOUS ERROR: the following file should exist, but doesn't: \n{}".format(setfile))
    with open(setfile) as f:
        settings = json.load(f)
    for key in settings:                       # every value goes through float() first (quirk Q9)
        try:
            settings[key] = float(settings[key])
        except (TypeError, ValueError):
            pass
    for key in settings:
        try:
            settings[key] = str2bool(settings[key])
        except AttributeError:
            pass
    return settings


class ExpressionTable:
    """What the reference's ExpressionDict holds (coexfunctions.py:160-198), as plain arrays:
    `names` in file order (the row order of the resident matrix, quirk Q1) and `values`."""

    def __init__(self, names, values, header):
        self.names = list(names)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.header = list(header)
        self.index = {nm: i for i, nm in enumerate(self.names)}

    def __len__(self):
        return len(self.names)

    def __contains__(self, key):
        return key in self.index

    def __getitem__(self, key):
        return list(self.values[self.index[key]])


def _expression_cache_key(filename):
    """Identity of the text table for the binary cache: size, mtime and a hash of its first and last MiB."""
    import hashlib
    st = os.stat(filename)
    h = hashlib.sha256()
    h.update(("%d:%d:" % (st.st_size, st.st_mtime_ns)).encode())
    with open(filename, "rb") as f:
        h.update(f.read(1 << 20))
        if st.st_size > (2 << 20):
            f.seek(st.st_size - (1 << 20))
            h.update(f.read(1 << 20))
    return h.hexdigest()


def read_expression_file(filename, cache=False):
    """Header row + label column, whitespace separated; columns with any NA are dropped and of
    duplicated labels the first occurrence is kept (coexfunctions.py:217-231).

    Row N4: every reference process re-parses the ~3 GB text table with pandas
    (coexfunctions.py:217).  With cache=True the parsed result (AFTER the NA-column and duplicate-label
    rules, so the semantics cannot drift) is kept next to the table as `<file>.b200cache/` -- float64
    row-major `values.npy` opened memory-mapped, labels, header, and the identity key of the text file
    (size, mtime, hash of the first and last MiB); a stale or foreign cache is ignored and rewritten."""
    cdir = filename + ".b200cache"
    key = None
    if cache:
        key = _expression_cache_key(filename)
        try:
            with open(os.path.join(cdir, "key.json")) as f:
                meta = json.load(f)
            if meta.get("key") == key and meta.get("format") == 1:
                values = np.load(os.path.join(cdir, "values.npy"), mmap_mode="r")
                with open(os.path.join(cdir, "names.txt")) as f:
                    names = f.read().split("\n")[:-1]
                with open(os.path.join(cdir, "header.txt")) as f:
                    header = f.read().split("\n")[:-1]
                if values.shape == (len(names), len(header)) and values.dtype == np.float64:
                    return ExpressionTable(names, values, header), header
        except (OSError, ValueError, KeyError):
            pass
    import pandas as pd
    df = pd.read_table(filename, header=0)
    label_col = df.columns[0]
    df = df.dropna(axis=1)
    df = df.drop_duplicates(subset=label_col)
    df = df.set_index(label_col)
    table = ExpressionTable(list(df.index), df.values, list(df.columns.values))
    if cache:
        try:
            os.makedirs(cdir, exist_ok=True)
            np.save(os.path.join(cdir, "values.npy"), table.values)
            with open(os.path.join(cdir, "names.txt"), "w") as f:
                f.write("".join(str(nm) + "\n" for nm in table.names))
            with open(os.path.join(cdir, "header.txt"), "w") as f:
                f.write("".join(str(h) + "\n" for h in table.header))
            with open(os.path.join(cdir, "key.json"), "w") as f:       # written last: a torn cache has no key
                json.dump({"key": key, "format": 1, "rows": len(table.names), "samples": len(table.header)}, f)
        except OSError:
            pass                                                       # read-only source directory: no cache
    return table, list(df.columns.values)


def readfeaturecoordinates(featfile):
    """-> (addresses {label: [chrom, start, end]}, feature_dict {chrom: {start: end}}, labels).
    Header lines are skipped until column 2 parses as an int; the label is the LAST column."""
    pa, fd, sl = {}, {}, []
    started = False
    with open(featfile) as f:
        for raw in f:
            line = raw.strip().split("\t")
            if not started:
                try:
                    int(line[2])
                    started = True
                except (IndexError, ValueError):
                    continue
            label, chrom = line[-1], line[0]
            start, end = int(line[1]), int(line[2])
            pa[label] = [chrom, start, end]
            sl.append(label)
            st, en = min(start, end), max(start, end)
            per = fd.setdefault(chrom, {})
            if st not in per or per[st] < en:
                per[st] = en
    return pa, fd, sl


def read_correlation_list(correlationfile):
    with open(correlationfile) as f:
        return np.array([float(x) for x in f if x.strip() != ""], dtype=np.float64)


def read_snp_map(mapfile, table, verbose=False):
    """windowBed join lines -> ordered {promoter: {"p": -9999, "snps": "a|b"}}; promoters not in
    the expression table are skipped (1-make-network.py:137-164).  `p` stays -9999 because the
    reference's lookup table `snp_p_values` does not exist in that script (bare excepts)."""
    snp_map = {}
    with open(mapfile) as f:
        for raw in f:
            line = raw.strip().split("\t")
            if len(line) < 4:
                continue
            prom = line[-1]
            if prom not in table:
                if verbose:
                    print("skipping", prom, "as not in expression file")
                continue
            snp = line[3]
            if prom in snp_map:
                snp_map[prom]["snps"] += "|" + snp
            else:
                snp_map[prom] = {"p": -9999, "snps": snp}
    return snp_map


def make_correlation_list(ctx, names, correlationfile, correlationmeasure="Spearman", gpl_len=1000000, seed=None,
                          py2_format=False):
    """Row N2: the global correlation list (reference 0-prepare-coex.py:621-716), the second caller of the
    pair-list ABI.  Existing entries are kept (sub-sampled to gpl_len), int(1.1 * missing) random row pairs are
    drawn, only pairs whose labels differ in their first five characters ("different chromosome", :682) are
    correlated -- on the GPU, through the resident table (ctx.pairs == getPearson) -- NaN -> -99, and the
    ascending list is written one value per line.  Returns the list as a numpy array.
    py2_format: write the values as the Python-2.7 reference does -- `"%s" % r` is str(float) there, 12 significant digits
    (0-prepare-coex.py:707-709) -- instead of Python 3's shortest round-trip repr; a reader (`float(line)`) takes either, but a
    list truncated to 12 digits bisects 1-2 entries differently, so this matters only for file-level equality with a deployed
    Python-2 reference."""
    import random
    rnd = random.Random(seed)
    existing = []
    if os.path.exists(correlationfile):
        with open(correlationfile) as f:
            lines = f.readlines()
        if len(lines) > gpl_len:
            lines = rnd.sample(lines, gpl_len)
        try:
            existing = [float(x) for x in lines if x.strip() != ""]
        except ValueError:
            existing = []
    new_to_add = int(min(gpl_len, gpl_len - len(existing)) * 1.1)
    if new_to_add > 0:
        n = len(names)
        ia, ib = [], []
        for _ in range(new_to_add):
            a, b = rnd.randint(0, n - 1), rnd.randint(0, n - 1)
            if names[a][:5] != names[b][:5]:
                ia.append(a)
                ib.append(b)
        r = ctx.pairs(np.array(ia, dtype=np.int32), np.array(ib, dtype=np.int32), ranked=(correlationmeasure == "Spearman"))
        existing += [(-99 if v != v else float(v)) for v in r]
        existing.sort()
        with open(correlationfile, "w") as o:
            for v in existing:
                o.write(("%.12g\n" % v) if py2_format else ("%s\n" % v))
    return np.array(existing, dtype=np.float64)


def empiricalp(thisvalue, empiricalcollection):
    if len(empiricalcollection) == 0:
        return 1
    empiricalcollection.sort()
    return 1 - float(bisect_left(empiricalcollection, thisvalue)) / len(empiricalcollection)


def fdr_correction(pvals, alpha=0.05, method="indep"):
    """Benjamini-Hochberg ('indep'/'poscorr') or Benjamini-Yekutieli ('negcorr') adjusted p."""
    pvals = np.asarray(pvals, dtype=np.float64)
    shape = pvals.shape
    flat = pvals.ravel()
    order = np.argsort(flat)
    ranked = flat[order]
    m = len(ranked)
    ecdf = np.arange(1, m + 1) / float(m) if m else np.zeros(0)
    if method in ("i", "indep", "p", "poscorr"):
        pass
    elif method in ("n", "negcorr"):
        harmonic = 0.0
        for k in range(1, m + 1):
            harmonic += 1.0 / k
        ecdf = ecdf / harmonic
    else:
        raise ValueError("Method should be 'indep' and 'negcorr'")
    reject = ranked < ecdf * alpha
    if reject.any():
        reject[:int(np.nonzero(reject)[0].max())] = True
    corrected = np.minimum.accumulate((ranked / ecdf)[::-1])[::-1] if m else ranked
    corrected = np.minimum(corrected, 1.0)
    back = np.argsort(order)
    return reject[back].reshape(shape), corrected[back].reshape(shape)
