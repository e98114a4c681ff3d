"""Golden vectors from the reference's OWN Python, executed in this container (tests/golden/run_reference_py3.py).

Run HERE, where /root/reference exists:
    make -C oracle && PYTHONHASHSEED=0 python tests/golden/make_reference_golden.py

Writes tests/golden/reference_run_<case>.json.gz: the inputs (table, features, global list, map-file promoters) and the
seven result objects `1-make-network.py` dumped for them (perm_results_store/<feat>.<object>.<k>), plus
tests/golden/reference_functions.json: known answers of the reference's `coexfunctions` helpers (get_gwad, get_address,
listitem, merge_ranges, empiricalp, fdr_correction, read_expression_file, readfeaturecoordinates) called directly,
tests/golden/reference_global_list.json.gz (row N2: lines 623-716 of 0-prepare-coex.py executed) and
tests/golden/reference_permutations.json (row N1: lines 471-613 executed, windowBed replaced by an echo of its -a file).  tests/test_reference_pin.py checks oracle/network_oracle.py
and oracle/permute_oracle.py against them -- nothing there reads /root/reference.
"""
import gzip
import json
import math
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from coexpression_b200 import synth  # noqa: E402
from oracle import network_oracle as no  # noqa: E402

OBJECTS = ("indmeasure", "individualscores", "prom_to_join", "indjoined", "joining_instructions", "winners",
           "allpromoterscores", "snp_map")

# name: (table seed, N, n, loci, window, settings overrides)
CASES = {
    "spearman_default": (44, 300, 24, 24, 30000, {}),
    "spearman_anticor": (45, 300, 24, 24, 30000, {"anticorrelation": "True"}),
    "spearman_noiter_j001": (46, 260, 31, 22, 40000, {"noiter": "True", "pval_to_join": 0.01}),
    "spearman_useavg_m": (47, 260, 20, 22, 40000, {"useaverage": "True", "selectionthreshold_is_j": "True"}),
    "pearson_default": (48, 240, 24, 20, 30000, {"correlationmeasure": "Pearson"}),
    "spearman_global_nsp0": (49, 240, 24, 20, 30000, {"nsp": 0.0}),
    "spearman_wide_join": (50, 300, 16, 26, 60000, {"pval_to_join": 0.6, "soft_separation": 400000}),
}


def run_case(name, seed, N, n, n_loci, window, overrides):
    t = synth.make_table(N, n, seed=seed)
    measure = overrides.get("correlationmeasure", "Spearman")
    with tempfile.TemporaryDirectory() as tmp:
        wd = os.path.join(tmp, "work")
        os.makedirs(os.path.join(wd, "permutation_store"))
        os.makedirs(os.path.join(wd, "perm_results_store"))
        expr, bed = os.path.join(tmp, "t.expression"), os.path.join(tmp, "f5ep.bed")
        clist = os.path.join(tmp, "t.expression.%s" % ("spearmanlist" if measure == "Spearman" else "correlationlist"))
        synth.write_expression_file(expr, t)
        synth.write_feature_bed(bed, t)
        data = no.rank_rows(t.X) if measure == "Spearman" else t.X
        ga, gb = synth.global_pair_indices(t.names, 4000, seed=seed + 100)
        g = no.get_pearson(data, ga, gb, "reference")
        glist = np.sort(np.where(np.isnan(g), -99.0, g))
        with open(clist, "w") as o:
            for v in glist:
                o.write("%s\n" % (int(v) if v == -99 else repr(float(v))))
        m = synth.Mapper(t)
        loci = synth.make_loci(t, n_loci, seed=seed + 200)
        sets = synth.circular_shifts(loci, m.genome_len, 1, seed=seed + 300)
        mapfiles = []
        for k, gw in enumerate(sets):
            mf = os.path.join(wd, "permutation_store", "f5ep.%d.bed" % k)
            synth.write_map_file(mf, t, m, gw, window, prefix="rs" if k == 0 else "rand_rs")
            with open(mf, "a") as o:                       # lines the reader must skip (1-make-network.py:139-146)
                o.write("1\t5\t6\trsNOTABLE\t1\t1\t2\tx\t0\t+\tchr1:1..2,+not_in_table\n")
                o.write("too\tshort\tline\n")
            mapfiles.append(mf)
        cfg = os.path.join(tmp, "app.cfg")
        with open(cfg, "w") as o:
            o.write("[directorypaths]\nsourcefilesdir=%s\nresdir=%s\npathtobedtools=/nonexistent\n" % (tmp, tmp))
        settings = {"verbose": "False", "feature_coordinates": bed, "correlationfile": clist, "expression_file": expr,
                    "supplementary_label": "x", "correlationmeasure": "Spearman", "soft_separation": 100000, "nsp": 1.0,
                    "seedfile": "none", "pval_to_join": 0.1, "useaverage": "False", "anticorrelation": "False",
                    "noiter": "False", "selectionthreshold": 0, "selectionthreshold_is_j": "False", "config_file": cfg}
        settings.update(overrides)
        with open(os.path.join(wd, "settings.txt"), "w") as o:
            json.dump(settings, o)
        runs = []
        for k, mf in enumerate(mapfiles):
            r = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_py3.py"), os.path.join(tmp, "app"), wd, mf],
                               capture_output=True, text=True, env=dict(os.environ, PYTHONHASHSEED="0"))
            assert r.returncode == 0, r.stderr[-3000:]
            store = os.path.join(wd, "perm_results_store")
            runs.append({obj: json.load(open(os.path.join(store, "f5ep.%s.%d" % (obj, k)))) for obj in OBJECTS})
        out = {
            "case": name, "settings": {k: v for k, v in settings.items() if k not in
                                       ("feature_coordinates", "correlationfile", "expression_file", "config_file")},
            "names": list(t.names), "X": [[float(v) for v in row] for row in t.X],
            "addresses": {nm: list(a) for nm, a in t.addresses().items()},
            "glist": [float(v) for v in glist], "runs": runs, "mapfiles": [open(mf).read() for mf in mapfiles],
            "how": "unmodified /root/reference/coexfunctions.py + mechanically 2to3-rewritten /root/reference/1-make-network.py, "
                   "executed by tests/golden/run_reference_py3.py against the unmodified reference C build",
        }
        path = os.path.join(HERE, "reference_run_%s.json.gz" % name)
        with gzip.open(path, "wt") as o:
            json.dump(out, o)
        print(name, "P =", [len(r["snp_map"]) for r in runs], "groups =", [len(r["prom_to_join"]) for r in runs],
              os.path.getsize(path), "bytes")


def function_vectors():
    from run_reference_py3 import import_reference_coexfunctions
    cf = import_reference_coexfunctions()
    rng = np.random.default_rng(7)
    chromlist = ["chr1", "chr2", "chr3", "chrX"]
    lengths = {"chr1": 1000, "chr2": 750, "chr3": 10, "chrX": 333}
    chromrange, acc = {}, 0
    for c in chromlist:
        chromrange[c] = [acc, acc + lengths[c]]
        acc += lengths[c]
    gw = []
    for _ in range(200):
        c = chromlist[int(rng.integers(0, 4))]
        a = int(rng.integers(-5, lengths[c] + 5))
        gw.append([c, a, cf.get_gwad(c, a, chromrange)])
    ad = []
    for g in list(range(-3, 12)) + [int(v) for v in rng.integers(0, 3 * acc, 200)] + [acc, acc + 1, 2 * acc, 2 * acc + 1]:
        ad.append([g, list(cf.get_address(g, chromrange, chromlist))])
    mr = []
    for _ in range(40):
        k = int(rng.integers(1, 12))
        starts = np.sort(rng.integers(0, 2000, k))
        rngs = np.stack([starts, starts + rng.integers(1, 60, k)], axis=1)
        ss = int(rng.choice([0, 10, 100, 500]))
        merged = [[int(a), int(b)] for a, b in cf.merge_ranges(rngs, ss)]
        mr.append([rngs.tolist(), ss, merged])
    ep = []
    for _ in range(60):
        coll = [float(v) for v in np.round(rng.normal(0, 3, int(rng.integers(0, 30))), 1)]
        v = float(np.round(rng.normal(0, 3), 1))
        ep.append([v, coll, cf.empiricalp(v, list(coll))])
    fd = []
    for method in ("indep", "negcorr"):
        for _ in range(15):
            p = [float(v) for v in rng.random(int(rng.integers(1, 40))) ** 3]
            rej, corr = cf.fdr_correction(p, 0.05, method)
            fd.append([p, method, [bool(x) for x in rej], [float(x) for x in corr]])
    # listitem (coexfunctions.py:319-325) rounds with the built-in round(): half away from zero in Python 2.7, half to
    # even in Python 3.  Only indices on which the two agree are recorded (no exact halves): the half-away rule of the
    # restatement is covered by its own known-answer test (tests/test_permute.py).
    li = []
    for _ in range(300):
        n_items = int(rng.integers(1, 400))
        index = int(rng.integers(-3 * n_items, 10_000_000_000))
        a = float(index) / n_items
        b = a - int(round(a, 0))
        c = b * n_items
        if abs(abs(a - math.floor(a)) - 0.5) < 1e-9 or abs(abs(c - math.floor(c)) - 0.5) < 1e-9:
            continue
        li.append([n_items, index, cf.listitem(list(range(n_items)), index)])
    # file readers (coexfunctions.py:201-274): an expression table with an NA column, a duplicated label and odd number
    # formats; a feature BED with header lines, a reversed interval and two features sharing a start
    expr_text = ("sample\ts0\ts1\tbad\ts2\n"
                 "chr1:10..20,+\t0.0\t1.5\t1\t2e-3\n"
                 "chr1:30..40,-\t3\t0.25\tNA\t7.125\n"
                 "chr2:5..9,+\t1e2\t0\t3\t0.1\n"
                 "chr1:10..20,+\t9\t9\t9\t9\n"
                 "chrX:1..2,-\t-0.0\t4.75\t5\t12345.678\n")
    bed_text = ("track name=features\n#chrom\tstart\tend\tname\n"
                "chr1\t10\t20\tx\t0\t+\tchr1:10..20,+\n"
                "chr1\t40\t30\tx\t0\t-\tchr1:30..40,-\n"
                "chr2\t5\t9\tx\t0\t+\tchr2:5..9,+\n"
                "chr2\t5\t12\tx\t0\t+\tchr2:5..12,+\n"
                "chrX\t1\t2\tx\t0\t-\tchrX:1..2,-\n")
    with tempfile.TemporaryDirectory() as tmp:
        expr_path, bed_path = os.path.join(tmp, "e.expression"), os.path.join(tmp, "f.bed")
        open(expr_path, "w").write(expr_text)
        open(bed_path, "w").write(bed_text)
        ed, header = cf.read_expression_file(expr_path)
        names = list(dict.keys(ed))
        readers = {"expr_text": expr_text, "names": names, "header": [str(h) for h in header],
                   "values": [[float(v) for v in ed[nm]] for nm in names], "bed_text": bed_text}
        r_pa, r_fd, r_sl = cf.readfeaturecoordinates(bed_path)
        readers.update(pa=r_pa, fd={c: {str(k): v for k, v in d.items()} for c, d in r_fd.items()}, sl=r_sl)
    with open(os.path.join(HERE, "reference_functions.json"), "w") as o:
        json.dump({"readers": readers, "chromlist": chromlist, "chromrange": chromrange, "get_gwad": gw, "get_address": ad, "listitem": li,
                   "merge_ranges": mr, "empiricalp": ep, "fdr_correction": fd,
                   "how": "direct calls of the unmodified /root/reference/coexfunctions.py under Python 3"}, o)
    print("reference_functions.json", len(gw), len(ad), len(mr), len(ep), len(fd))


def collate_vectors():
    """Row N3: the reference's 2-collate-results.py executed on a perm_results_store that its own 1-make-network.py wrote
    (one real hit set + three permuted ones); the main table `<label>.individual_scores` and `allscores.txt` it writes are
    the known answers for coexpression_b200/collate.py."""
    t = synth.make_table(260, 20, seed=71)
    with tempfile.TemporaryDirectory() as tmp:
        wd = os.path.join(tmp, "work")
        os.makedirs(os.path.join(wd, "permutation_store"))
        os.makedirs(os.path.join(wd, "perm_results_store"))
        expr, bed = os.path.join(tmp, "t.expression"), os.path.join(tmp, "f5ep.bed")
        clist = os.path.join(tmp, "t.expression.spearmanlist")
        synth.write_expression_file(expr, t)
        synth.write_feature_bed(bed, t)
        ga, gb = synth.global_pair_indices(t.names, 4000, seed=171)
        g = no.get_pearson(no.rank_rows(t.X), ga, gb, "reference")
        glist = np.sort(np.where(np.isnan(g), -99.0, g))
        with open(clist, "w") as o:
            for v in glist:
                o.write("%s\n" % (int(v) if v == -99 else repr(float(v))))
        m = synth.Mapper(t)
        loci = synth.make_loci(t, 18, seed=271)
        sets = synth.circular_shifts(loci, m.genome_len, 3, seed=371)
        cfg = os.path.join(tmp, "app.cfg")
        with open(cfg, "w") as o:
            o.write("[directorypaths]\nsourcefilesdir=%s\nresdir=%s\npathtobedtools=/nonexistent\nf5resource=http://x/\n" % (tmp, tmp))
        settings = {"verbose": "False", "feature_coordinates": bed, "correlationfile": clist, "expression_file": expr,
                    "supplementary_label": "golden", "correlationmeasure": "Spearman", "soft_separation": 100000, "nsp": 1.0,
                    "seedfile": "none", "pval_to_join": 0.1, "useaverage": "False", "anticorrelation": "False",
                    "noiter": "False", "selectionthreshold": 0, "selectionthreshold_is_j": "False", "config_file": cfg,
                    "useremail": "none", "snp_count": 18, "genome_length": int(m.genome_len)}
        with open(os.path.join(wd, "settings.txt"), "w") as o:
            json.dump(settings, o)
        env = dict(os.environ, PYTHONHASHSEED="0")
        for k, gw in enumerate(sets):
            mf = os.path.join(wd, "permutation_store", "f5ep.%d.bed" % k)
            synth.write_map_file(mf, t, m, gw, 30000, prefix="rs" if k == 0 else "rand_rs")
            r = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_py3.py"), os.path.join(tmp, "app"), wd, mf],
                               capture_output=True, text=True, env=env)
            assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_py3.py"), "collate", os.path.join(tmp, "app"), wd],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-3000:]
        store = os.path.join(wd, "perm_results_store")
        files = {f: open(os.path.join(store, f)).read() for f in sorted(os.listdir(store))
                 if f.endswith(".0") or ".indmeasure." in f}
        out = {"store": files, "settings": {k: v for k, v in settings.items() if k in ("supplementary_label", "verbose")},
               "individual_scores": open(os.path.join(wd, "golden.individual_scores")).read(),
               "allscores": open(os.path.join(wd, "allscores.txt")).read(), "feature_bed": open(bed).read(),
               # the reporting files of 2-collate-results.py:165-430 (html table, circos network, BioLayout file, stats, file list,
               # and the qq_plot file, of which the reference only ever writes the header: `fullrange_perm_indm` is undefined)
               "reports": {f: open(os.path.join(wd, f)).read() for f in ("golden.html", "golden.network.php", "golden.layout",
                                                                          "golden_stats.html", "golden.filelist", "golden.qq_plot")},
               "map0": open(os.path.join(wd, "permutation_store", "f5ep.0.bed")).read(),
               "full_settings": dict(settings, config_file="<cfg>"), "f5resource": "http://x/",
               "how": "unmodified coexfunctions.py + mechanically 2to3-rewritten 1-make-network.py and 2-collate-results.py"}
    with gzip.open(os.path.join(HERE, "reference_collate.json.gz"), "wt") as o:
        json.dump(out, o)
    print("collate:", len(out["individual_scores"].splitlines()) - 1, "groups,", len(files), "store files")


def permutation_vectors():
    """Row N1: the permuted locus positions, by EXECUTING the reference's permutation block (0-prepare-coex.py, from
    `writequeue('Starting to make permutations` to the `windowBed` call of every set) for the circular, backcirc and
    background modes with a seeded `random`.  `windowbed` is pointed at a two-line script that copies its `-a` file (the
    temporary BED of the shifted loci) to stdout, so every `permutation_store/<feat>.<k>.bed` the block writes holds the
    positions it handed to windowBed -- the join itself is bedtools' and stays unpinned."""
    import random
    import subprocess
    import types
    from subprocess import PIPE, STDOUT
    from run_reference_py3 import convert_py2, import_reference_coexfunctions, REF
    cf = import_reference_coexfunctions()
    with open(os.path.join(REF, "0-prepare-coex.py")) as f:
        lines = f.read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.startswith("writequeue('Starting to make permutations"))
    last = next(i for i, l in enumerate(lines) if i > first and l.strip().startswith("print output  # so that errors"))
    assert (first, last) == (470, 612), (first, last)    # file lines 471..613 of the reference as surveyed
    code = compile(convert_py2("\n".join(lines[first:last + 1])), "0-prepare-coex.py[471:613]", "exec")
    chromlist = ["chr1", "chr2", "chr3", "chrX"]
    lengths = {"chr1": 5000, "chr2": 3000, "chr3": 800, "chrX": 1200}
    chromrange, acc = {}, 0
    for c in chromlist:
        chromrange[c] = [acc, acc + lengths[c]]
        acc += lengths[c]
    rng = np.random.default_rng(21)
    snps = []
    for i in range(40):
        c = chromlist[int(rng.integers(0, 4))]
        snps.append((c, int(rng.integers(1, lengths[c] - 1)), "rs%d" % i))
    snp_bed_data = {}
    for c, a, name in snps:
        snp_bed_data.setdefault(c, {})[a] = name
    flat = [(c, a) for c in snp_bed_data for a in snp_bed_data[c]]      # the iteration order of the block's loops
    backlines = []
    for i in range(600):
        c = chromlist[int(rng.integers(0, 4))]
        a = int(rng.integers(1, lengths[c] - 1))
        backlines.append("%s\t%d\t%d\tbg%d\trsb%d\n" % (c, a, a + 1, i, i))
    out = []
    for mode, seed in (("circular", 31), ("backcirc", 32), ("background", 33)):
        with tempfile.TemporaryDirectory() as tmp:
            store = os.path.join(tmp, "permutation_store")
            os.makedirs(store)
            cap = os.path.join(tmp, "capture.py")
            with open(cap, "w") as o:
                o.write("import sys\nsys.stdout.write(open(sys.argv[sys.argv.index('-a') + 1]).read())\n")
            bgfile = os.path.join(tmp, "background.bed")
            with open(bgfile, "w") as o:
                o.writelines(backlines)
            ns = {name: getattr(cf, name) for name in dir(cf) if not name.startswith("__")}
            random.seed(seed)
            numperms = 6
            ns.update(writequeue=lambda *a, **k: None, queuefile=os.path.join(tmp, "q"), permutationmode=mode, verbose=False,
                      snps_mapped=os.path.join(store, "f5ep.0.bed"), primary_dict={"a": 1, "b": 2}, random=random,
                      numperms=numperms, storage_dir_permutations=store, feature_label="f5ep", backgroundfile=bgfile,
                      snp_bed_data=snp_bed_data, chromrange=chromrange, chromlist=chromlist, snp_count=len(flat),
                      doitagain=False, working_files_dir=tmp, windowbed="%s %s" % (sys.executable, cap), window=1234,
                      args=types.SimpleNamespace(feature_coordinates=os.path.join(tmp, "feat.bed")),
                      subprocess=subprocess, PIPE=PIPE, STDOUT=STDOUT)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                exec(code, ns)
            shifts = list(ns["shifts"])
            sets = []
            for k, shift in enumerate(shifts):
                path = ns["mapfiles"][shift]
                rows = [l.rstrip("\n").split("\t") for l in open(path)]
                sets.append({"shift": int(shift), "file": os.path.basename(path),
                             "positions": [["chr" + r[0].replace("chr", ""), int(r[1]), int(r[2]), r[3]] for r in rows]})
        out.append({"mode": mode, "seed": seed, "numperms": numperms, "snps": [list(x) for x in flat],
                    "backlines": backlines if mode != "circular" else [], "sets": sets})
        print("permutation block", mode, "shifts", shifts[:3], "...", [len(s["positions"]) for s in sets])
    with open(os.path.join(HERE, "reference_permutations.json"), "w") as o:
        json.dump({"chromlist": chromlist, "chromrange": chromrange, "runs": out,
                   "how": "lines 471-613 of /root/reference/0-prepare-coex.py, mechanically 2to3-rewritten, executed with a seeded "
                          "`random`; windowbed replaced by a script that echoes its -a file"}, o)


def global_list_vectors():
    """Row N2: the global correlation list, by EXECUTING lines 623-716 of the reference's 0-prepare-coex.py (the block that
    draws random row pairs, calls getPearson and writes <expr>.spearmanlist) in a namespace that holds what the lines
    before it would have set up.  `random` is seeded, so coexfunctions.make_correlation_list (same call sequence on a
    random.Random(seed)) must draw the same pairs."""
    import ctypes
    import random
    import types
    from run_reference_py3 import convert_py2, import_reference_coexfunctions, REF
    cf = import_reference_coexfunctions()
    with open(os.path.join(REF, "0-prepare-coex.py")) as f:
        lines = f.read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.startswith('if verbose: print ("reading correlation lists:"'))
    last = next(i for i, l in enumerate(lines) if "Correlation Centile at" in l)
    assert (first, last) == (622, 715), (first, last)    # file lines 623..716 of the reference as surveyed
    block = "\n".join(lines[first:last + 1])
    code = compile(convert_py2(block), "0-prepare-coex.py[623:716]", "exec")
    out = []
    for case, (measure, seed, pre_existing) in {"spearman_fresh": ("Spearman", 5, 0), "pearson_fresh": ("Pearson", 6, 0),
                                                "spearman_topup": ("Spearman", 7, 700)}.items():
        t = synth.make_table(220, 18, seed=60 + seed)
        with tempfile.TemporaryDirectory() as tmp:
            expr = os.path.join(tmp, "t.expression")
            synth.write_expression_file(expr, t)
            corrfile = os.path.join(tmp, "t.expression.list")
            existing = []
            if pre_existing:
                rng = np.random.default_rng(seed)
                existing = sorted(float(np.round(v, 6)) for v in rng.uniform(-0.9, 0.9, pre_existing))
                with open(corrfile, "w") as o:
                    for v in existing:
                        o.write("%s\n" % v)
            ns = {name: getattr(cf, name) for name in dir(cf) if not name.startswith("__")}
            ns.update({name: getattr(ctypes, name) for name in ("c_double", "c_int", "byref", "cdll")})
            random.seed(seed)
            ns.update(verbose=False, correlationmeasure=measure, correlationfile=corrfile, gpl_len=2000,
                      permutationmode="circular", args=types.SimpleNamespace(expression_file=expr), random=random,
                      coexpression=ctypes.cdll.LoadLibrary(os.path.join(ROOT, "oracle", "_ref", "coexpression_v2.so")))
            exec(code, ns)
            with open(corrfile) as f:
                values = [float(x) for x in f if x.strip()]
        out.append({"case": case, "measure": measure, "seed": seed, "gpl_len": 2000, "existing": existing,
                    "names": list(t.names), "X": [[float(v) for v in row] for row in t.X], "values": values})
        print("global list", case, len(values), "values")
    with gzip.open(os.path.join(HERE, "reference_global_list.json.gz"), "wt") as o:
        json.dump({"runs": out, "how": "lines 623-716 of /root/reference/0-prepare-coex.py, mechanically 2to3-rewritten, executed with a "
                                       "seeded `random` against the unmodified coexfunctions.py and the unmodified reference C build"}, o)


if __name__ == "__main__":
    assert "reference" in no.available_kinds(), "build oracle/_ref/coexpression_v2.so first (make -C oracle)"
    function_vectors()
    global_list_vectors()
    permutation_vectors()
    collate_vectors()
    if len(sys.argv) > 1 and sys.argv[1] == "functions":
        sys.exit(0)
    for name, spec in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run_case(name, *spec)
