"""The 0-prepare-coex.py drop-in (coexpression_b200/prepare.py): command-line contract against the reference's
(tests/golden/reference_prepare_flags.json, extracted from its source by tests/golden/make_prepare_flags.py), host
readers, and -- on a GPU -- the whole prepare -> make-network -> collate run on a synthetic supfiles directory."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, has_cuda

from coexpression_b200 import prepare


def test_flags_defaults_and_choices_equal_the_reference():
    with open(os.path.join(GOLDEN, "reference_prepare_flags.json")) as f:
        ref = json.load(f)
    parser = prepare.build_parser("/app")
    mine = {tuple(a.option_strings): a for a in parser._actions if a.option_strings and a.option_strings[0] != "-h"}
    seen = set()
    for fl in ref["flags"]:
        key = tuple(fl["names"])
        assert key in mine, key
        seen.add(key)
        act = mine[key]
        if fl.get("action") == "store_true":
            assert act.nargs == 0 and act.const is True and act.default is False, key
        if "default" in fl and fl["default"] != "<expr>":
            assert act.default == fl["default"], key
        if "choices" in fl:
            assert list(act.choices) == fl["choices"], key
        if "type" in fl:
            assert act.type.__name__ == fl["type"], key
    extra = set(mine) - seen
    assert all(k[0].startswith("--") for k in extra), extra      # extensions are long options only
    # `dest` names are the settings.txt keys 1-make-network.py / 2-collate read back
    for dest in ("expression_file", "feature_coordinates", "correlationmeasure", "anticorrelation", "noiter", "useaverage",
                 "nsp", "pval_to_join", "selectionthreshold", "selectionthreshold_is_j", "soft_separation", "seedfile",
                 "numperms", "verbose"):
        assert any(a.dest == dest for a in parser._actions), dest


def test_output_label_and_readers(tmp_path):
    parser = prepare.build_parser("/app")
    a = parser.parse_args(["-f", "/x/gwas.bed", "-x", "/s/tab.expression", "-a", "-i", "-w", "2000", "-j", "0.01", "-p", "backcirc"])
    assert prepare.output_label(a, a.snp_details_file) == "gwas_complete_BACKCIRC_ANTIin_NOiter_w2000_pj0.01_tab"
    a = parser.parse_args(["-f", "/x/gwas.bed", "-x", "/s/tab.expression", "-s", "0", "-u", "-m"])
    assert prepare.output_label(a, a.snp_details_file) == "gwas_s0.0_CIRCULAR_SIJ_useav_pj0.1_tab"
    cl = tmp_path / "chromlengths.txt"
    cl.write_text("".join("chr%s: %d\n" % (c, 1000 * (i + 1)) for i, c in enumerate(list(range(1, 23)) + ["X", "Y", "M"])))
    lengths, genome = prepare.read_chrom_lengths(str(cl))
    assert lengths["chr2"] == 2000 and genome == sum(1000 * (i + 1) for i in range(25))
    cr, chromlist = prepare.chromosome_ranges(lengths)
    assert chromlist[0] == "chr1" and chromlist[-1] == "chrM" and cr["chr2"] == [1000, 3000]
    snp = tmp_path / "in.bed"
    snp.write_text("#comment\n1\t10\t11\trs1\nchr2 20 21 rs2\n3\t30\t31\nGENE1\nchrUn\t5\t6\trs9\n1\t10\t11\trs1b\n")
    data, all_snps, bad = prepare.read_snps(str(snp), lengths)
    assert data == {"chr1": {10: "rs1b"}, "chr2": {20: "rs2"}, "chr3": {30: "3"}}       # same address: the later name wins
    assert all_snps == ["rs1", "rs2", "3", "rs1b"] and bad == ["GENE1"]


def _supfiles(tmp_path, N=3000, n=64, loci=80, seed=77):
    from coexpression_b200 import synth
    t = synth.make_table(N, n, seed=seed)
    src = tmp_path / "supfiles"
    src.mkdir()
    synth.write_expression_file(str(src / "tab.expression"), t)
    synth.write_feature_bed(str(src / "feat.bed"), t)
    with open(src / "chromlengths.txt", "w") as o:
        for c in ["chr%s" % x for x in list(range(1, 23)) + ["X", "Y"]]:
            o.write("%s: %d\n" % (c, t.chrom_len[c]))
        o.write("chrM: 16571\n")
    m = synth.Mapper(t)
    g = synth.make_loci(t, loci, seed + 1)
    ci, pos = m.gwad_to_chrom(g)
    with open(src / "gwas.bed", "w") as o:
        for k in range(loci):
            o.write("%s\t%d\t%d\trs%d\n" % (m.chrom_list[ci[k]].replace("chr", ""), pos[k], pos[k] + 1, k))
    cfg = tmp_path / "app.cfg"
    cfg.write_text("[directorypaths]\nsourcefilesdir = %s/\nresdir = %s/\npathtobedtools = ''\npathtopython = python\n"
                   "f5resource = none\n" % (src, tmp_path / "results"))
    return t, src, cfg


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_prepare_run_collate_end_to_end(tmp_path):
    """`0-prepare-coex.py -f gwas.bed -n 4 -w 3000 ...` on synthetic supfiles: directory layout and files of the reference,
    map files equal to the restated windowBed join of the same shifted loci, densities equal to the oracle's, collation run."""
    import random
    from oracle import network_oracle as no
    from oracle import permute_oracle as po
    t, src, cfg = _supfiles(tmp_path)
    argv = ["-f", str(src / "gwas.bed"), "-x", "SFD:tab.expression", "-q", "SFD:feat.bed", "-cl", "SFD:chromlengths.txt",
            "-cc", str(cfg), "-n", "4", "-w", "3000", "-gpl", "20000", "--rng_seed", "11"]
    assert prepare.main(argv + ["-gwd"]) == 0
    assert prepare.main(argv) == 0
    wd = str(tmp_path / "results") + "/gwas_complete_CIRCULAR_w3000_pj0.1_tab"
    for name in ("settings.txt", "jobfile.txt", "collationcommandfile.txt", "coex.queue", "permutation_store", "perm_results_store"):
        assert os.path.exists(os.path.join(wd, name)), name
    with open(os.path.join(wd, "settings.txt")) as f:
        st = json.load(f)
    with open(os.path.join(GOLDEN, "reference_prepare_flags.json")) as f:
        ref = json.load(f)
    for fl in ref["flags"]:
        assert fl["names"][1].lstrip("-") in st, fl["names"]
    for key in ref["settings_extra_keys"]:
        assert key in st, key
    assert st["correlationfile"] == str(src / "tab.expression") + ".spearmanlist" and os.path.exists(st["correlationfile"])
    jobs = open(os.path.join(wd, "jobfile.txt")).read().split("\n")
    assert len(jobs) == 5 and jobs[0].endswith("-mf %s/permutation_store/feat.0.bed -wd %s" % (wd, wd))
    # ---- map files: the promoters of every set == restated join of the loci shifted by the same RNG stream ----------
    lengths, _ = prepare.read_chrom_lengths(str(src / "chromlengths.txt"))
    data, all_snps, _ = prepare.read_snps(str(src / "gwas.bed"), lengths)
    cr, chromlist = prepare.chromosome_ranges(lengths)
    loci = [(c, ad) for c in data for ad in data[c]]
    rng = random.Random(11)
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(4)]
    order = [0, 1, 2, 3, 4]
    hits = []
    for k, s in zip(order, shifts):
        want = po.hit_rows(po.window_bed(po.circular_positions(loci, s, cr, chromlist), t.chrom, t.start, t.end, 3000))
        lines = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(wd, "permutation_store", "feat.%d.bed" % k))]
        assert [ln[-1] for ln in lines] == [t.names[r] for r in want], k
        assert all(len(ln) >= 4 for ln in lines)
        hits.append(want)
    assert len(hits[0]) > 10
    # ---- densities of every set against the oracle ----------------------------------------------------------------------
    glist = [float(x) for x in open(st["correlationfile"]) if x.strip()]
    assert glist == sorted(glist) and len(glist) > 15000
    R = no.rank_rows(t.X)
    for k in order:
        with open(os.path.join(wd, "perm_results_store", "feat.indmeasure.%d" % k)) as f:
            got = json.load(f)
        want = no.make_network(t.names, t.X, [t.names[r] for r in hits[k]], t.addresses(), glist, Xrank=R)
        assert set(got) == set(want["indmeasure"])
        for g_, v in want["indmeasure"].items():
            assert abs(got[g_] - v) <= 1e-6
    # ---- collation ran (2-collate-results.py outputs this repo writes) -----------------------------------------------------
    assert any(f.endswith(".individual_scores") for f in os.listdir(wd)) or os.path.exists(os.path.join(wd, "allscores.txt"))
    # ---- resume: a second call finds the permutations and reruns only the batch -------------------------------------------
    assert prepare.main(argv + ["-po"]) == 0
    # ---- --lean_results (B200 extension): the P x P objects of the real data only, same collation ----------------------------
    assert prepare.main(argv + ["-o", "gwas_lean", "--lean_results"]) == 0
    wl = str(tmp_path / "results") + "/gwas_lean"
    store = sorted(os.listdir(os.path.join(wl, "perm_results_store")))
    assert [f for f in store if "individualscores" in f or "indjoined" in f] == ["feat.individualscores.0", "feat.indjoined.0"]
    assert len(store) == 8 + 4 * 6
    for k in order:
        a_, b_ = (open(os.path.join(d_, "perm_results_store", "feat.indmeasure.%d" % k)).read() for d_ in (wd, wl))
        assert a_ == b_
    for name in ("allscores.txt",) + tuple(f for f in os.listdir(wd) if f.endswith(".individual_scores")):
        other = name.replace("gwas_complete_CIRCULAR_w3000_pj0.1_tab", "gwas_lean")
        assert open(os.path.join(wd, name)).read() == open(os.path.join(wl, other)).read(), name
