"""CPU-only: the Python-3 mirror of the reference's file readers / naming (no GPU, no oracle arithmetic)."""
import json
import os

import numpy as np

from coexpression_b200 import synth
from coexpression_b200.coexfunctions import (empiricalp, fdr_correction, read_correlation_list, read_expression_file,
                                             read_snp_map, readfeaturecoordinates, readsettings)
from coexpression_b200.make_network import options_from_settings, result_path


def test_expression_reader_semantics(tmp_path):
    t = synth.make_table(40, 6, seed=2)
    path = tmp_path / "x.expression"
    synth.write_expression_file(path, t)
    # append a duplicated label (first occurrence wins) and a column with an NA (dropped)
    lines = open(path).read().splitlines()
    lines[0] += "\tbad"
    for k in range(1, len(lines)):
        lines[k] += "\t1.0" if k != 3 else "\tNA"
    lines.append(lines[1].split("\t")[0] + "\t" + "\t".join(["9.0"] * 7))
    open(path, "w").write("\n".join(lines) + "\n")
    table, header = read_expression_file(str(path))
    assert table.names == t.names                       # file order = row order of the resident matrix (Q1)
    assert header == [f"s{j}" for j in range(6)]        # NA column dropped (coexfunctions.py:221)
    assert np.array_equal(table.values, t.X)            # duplicate label: first kept (coexfunctions.py:225)
    assert table[t.names[5]] == list(t.X[5]) and t.names[7] in table and "nope" not in table


def test_expression_binary_cache_row_n4(tmp_path):
    """Row N4: the cached table equals the parsed one bit for bit (same NA-column / duplicate-label rules), is
    memory-mapped on the second read, and is rebuilt when the text file changes."""
    t = synth.make_table(50, 7, seed=4)
    path = tmp_path / "y.expression"
    synth.write_expression_file(path, t)
    lines = open(path).read().splitlines()
    lines[0] += "\tbad"
    for k in range(1, len(lines)):
        lines[k] += "\t2.5" if k != 4 else "\tNA"
    lines.append(lines[2].split("\t")[0] + "\t" + "\t".join(["7.0"] * 8))
    open(path, "w").write("\n".join(lines) + "\n")
    plain, h0 = read_expression_file(str(path))
    first, h1 = read_expression_file(str(path), cache=True)          # parses and writes <file>.b200cache/
    assert os.path.exists(str(path) + ".b200cache/key.json")
    second, h2 = read_expression_file(str(path), cache=True)         # served from the cache
    assert isinstance(second.values, np.memmap) or isinstance(second.values.base, np.memmap)
    for tab, hd in ((first, h1), (second, h2)):
        assert tab.names == plain.names and hd == h0 and tab.header == plain.header
        assert tab.values.dtype == np.float64 and np.array_equal(np.asarray(tab.values), plain.values)
        assert tab.index == plain.index
    # a changed text file invalidates the cache
    lines[1] = lines[1].split("\t")[0] + "\t" + "\t".join(["3.25"] * 8)
    open(path, "w").write("\n".join(lines) + "\n")
    os.utime(path, ns=(1, 1))                                         # even with an old mtime the size/hash key differs
    third, _ = read_expression_file(str(path), cache=True)
    assert np.all(np.asarray(third.values)[0] == 3.25)
    fourth, _ = read_expression_file(str(path), cache=True)
    assert np.array_equal(np.asarray(fourth.values), np.asarray(third.values))


def test_feature_reader_and_map_reader(tmp_path):
    t = synth.make_table(60, 4, seed=3)
    bed = tmp_path / "f.bed"
    synth.write_feature_bed(bed, t)
    body = open(bed).read()
    open(bed, "w").write("track name=x\nbrowser hide all\n" + body)               # header lines are skipped
    pa, fd, sl = readfeaturecoordinates(str(bed))
    assert sl == t.names and pa[t.names[3]] == [t.chrom[3], int(t.start[3]), int(t.end[3])]
    table, _ = (lambda p: read_expression_file(str(p)))(_write_expr(tmp_path, t))
    mp = tmp_path / "feat.3.bed"
    with open(mp, "w") as o:
        o.write(f"1\t10\t11\trsA\tchr1\t5\t9\t{t.names[9]}\n")
        o.write("short\tline\n")                                                   # < 4 columns: skipped
        o.write(f"1\t10\t11\trsB\tchr1\t5\t9\t{t.names[2]}\n")
        o.write(f"1\t10\t11\trsC\tchr1\t5\t9\t{t.names[9]}\n")                      # second SNP for the same promoter
        o.write("1\t10\t11\trsD\tchr1\t5\t9\tnot_in_table\n")                       # skipped (1-make-network.py:142-146)
    snp_map = read_snp_map(str(mp), table)
    assert list(snp_map) == [t.names[9], t.names[2]]                                # first-appearance order
    assert snp_map[t.names[9]] == {"p": -9999, "snps": "rsA|rsC"}
    assert result_path("/x/perm_results_store", str(mp), "indmeasure") == "/x/perm_results_store/feat.indmeasure.3"


def _write_expr(tmp_path, t):
    p = tmp_path / "e.expression"
    synth.write_expression_file(p, t)
    return p


def test_settings_roundtrip_quirk_q9(tmp_path):
    st = {"verbose": "False", "anticorrelation": "True", "nsp": "1.0", "pval_to_join": 0.01, "noiter": False,
          "useaverage": "false", "selectionthreshold": 0, "selectionthreshold_is_j": "True", "soft_separation": "100000",
          "correlationmeasure": "Pearson", "numperms": "100", "expression_file": "/a/b.expression"}
    json.dump(st, open(tmp_path / "settings.txt", "w"))
    got = readsettings(str(tmp_path))
    assert got["anticorrelation"] is True and got["verbose"] is False and got["numperms"] == 100.0 and got["nsp"] == 1.0
    assert got["expression_file"] == "/a/b.expression"
    o = options_from_settings(got)
    assert (o.correlationmeasure, o.anticorrelation, o.noiter, o.useaverage) == ("Pearson", True, False, False)
    assert o.selectionthreshold_is_j is True and o.soft_separation == 100000 and o.pval_to_join == 0.01


def test_correlation_list_and_collation_maths(tmp_path):
    p = tmp_path / "x.spearmanlist"
    open(p, "w").write("-99\n-0.5\n\n0.25\n0.75\n")
    assert list(read_correlation_list(str(p))) == [-99.0, -0.5, 0.25, 0.75]
    from oracle import network_oracle as no             # checker: same maths restated independently
    pv = [0.2, 0.01, 0.03, 0.04, 0.5]
    for method in ("indep", "negcorr"):
        a = fdr_correction(pv, 0.05, method); b = no.fdr_correction(pv, 0.05, method)
        assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1], rtol=0, atol=1e-15)
    pool = [3.0, 1.0, 2.0, 2.0]
    assert empiricalp(2.0, pool) == 0.75 and empiricalp(9.0, pool) == 0.0 and empiricalp(1.0, []) == 1


def test_collate_numbers_against_oracle(tmp_path):
    """Row N3: pooled null, empirical p, BY / BH and the positional pairing of the main table."""
    from coexpression_b200.collate import collate
    from oracle import network_oracle as no
    rng = np.random.default_rng(3)
    wd = tmp_path / "w"; store = wd / "perm_results_store"; store.mkdir(parents=True)
    groups = [f"chr1:{100 * i}..{100 * i + 10},+" for i in range(9)]
    real = {g: float(v) for g, v in zip(groups, np.round(rng.gamma(2.0, 3.0, 9), 3))}
    real[groups[4]] = real[groups[3]]                       # tied densities
    objs = {"indmeasure": real, "individualscores": {}, "prom_to_join": {g: [g] for g in groups},
            "indjoined": {g: {h: 1.0 for h in groups if h != g} for g in groups},
            "joining_instructions": {g: g for g in groups}, "winners": {g: g for g in groups},
            "allpromoterscores": {g: 0.0 for g in groups}, "snp_map": {g: {"p": -9999, "snps": f"rs{i}|rs{i}b"} for i, g in enumerate(groups)}}
    for name, obj in objs.items():
        json.dump(obj, open(store / f"f5ep.{name}.0", "w"))
    perms = []
    for k in range(1, 6):
        d = {f"g{j}": float(v) for j, v in enumerate(np.round(rng.gamma(2.0, 2.5, 7), 3))}
        perms.append(d)
        json.dump(d, open(store / f"f5ep.indmeasure.{k}", "w"))
    json.dump({"supplementary_label": "lab", "verbose": "False"}, open(wd / "settings.txt", "w"))
    got = collate(str(wd))
    ps, by, bh = no.collate(real, perms)
    assert got["p"] == ps and np.allclose(got["fdr_by"], by, atol=1e-15) and np.allclose(got["fdr_bh"], bh, atol=1e-15)
    assert got["num_completed_perms"] == 5
    rows = open(got["individual_scores_file"]).read().splitlines()
    assert len(rows) == 10 and rows[0].startswith("Top_promoter[SNPs_in_top_promoter]")
    order = sorted(real.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)
    first = rows[1].split("\t")
    assert first[2] == order[0][0] and float(first[5]) == order[0][1] / 9 and float(first[6]) == ps[0]
