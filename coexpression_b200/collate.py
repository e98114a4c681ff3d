"""Row N3 of SURVEY.md section 8f: the arithmetic of the reference's 2-collate-results.py and its main table.

  pooled null ........ every permuted `indmeasure` value, k != 0          2-collate-results.py:67-76
  real measures ...... [indmeasure[g] for g in indjoined]                 :142
  empirical p ........ 1 - bisect_left(sorted(pool), x) / len             :161, coexfunctions.py:379-383
  FDR ................ Benjamini-Yekutieli ('negcorr') and BH ('indep')   :162-163, coexfunctions.py:333-358
  table .............. `<label>.individual_scores`: groups by (score, name) descending, the i-th largest score is
                       paired with the i-th smallest p (:176-244); reported score = indmeasure / len(indmeasure) (:242)

  allscores.txt ...... one line per hit: address, group, allpromoterscores, group density, group FDR (:264-283)

  reports ............ the remaining files of :165-430 (`write_reports`): `<label>.html` (results table), `<label>.network.php`
                       (circos network), `<label>.layout` (BioLayout / d3), `<label>_stats.html`, `<label>.filelist` and the
                       `<label>.qq_plot` file -- of which the reference only ever writes the header line (`fullrange_perm_indm`
                       is an undefined name inside a bare try/except, :148-158).  Text for text equal to what the executed
                       reference writes (tests/golden/reference_collate.json.gz), with two documented freedoms: the
                       `//NODECLASS` lines of the layout file follow the iteration order of a Python set in the reference
                       (hash order; written sorted here), and `round()` is Python 3's as in the executed reference.  The
                       reference's metadata file is read through an undefined name (`promoter_details`, :113-119), i.e. never.
Same command line: `2-collate-results.py -wd <working_files_dir> [-t N]`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

from .coexfunctions import empiricalp, fdr_correction, readfeaturecoordinates, readsettings


def load_store(store):
    files = os.listdir(store)
    perm_files = sorted(f for f in files if ".indmeasure." in f and not f.endswith(".0"))
    pool = []
    for f in perm_files:
        with open(os.path.join(store, f)) as fh:
            pool += list(json.load(fh).values())
    stems = [f for f in files if f.endswith(".indmeasure.0")]
    if len(stems) != 1:
        raise RuntimeError("Error: expected exactly one real-data result set in {}".format(store))
    stem = stems[0].replace(".indmeasure.0", "")
    real = {}
    for name in ("indmeasure", "individualscores", "prom_to_join", "indjoined", "joining_instructions", "winners",
                 "allpromoterscores", "snp_map"):
        with open(os.path.join(store, f"{stem}.{name}.0")) as fh:
            real[name] = json.load(fh)
    return pool, real, len(perm_files)


def collate_numbers(pool, real):
    """-> dict(realmeasures, p (ascending), fdr_by, fdr_bh) exactly as 2-collate-results.py:142,161-163."""
    realmeasures = [real["indmeasure"][g] for g in real["indjoined"]]
    pool = list(pool)
    p = sorted(empiricalp(x, pool) for x in realmeasures)
    return {"realmeasures": realmeasures, "p": p,
            "fdr_by": list(fdr_correction(p, 0.05, "negcorr")[1]) if p else [],
            "fdr_bh": list(fdr_correction(p)[1]) if p else []}


def write_individual_scores(path, real, numbers):
    indmeasure, winners, ptj = real["indmeasure"], real["winners"], real["prom_to_join"]
    ji, snp_map = real["joining_instructions"], real["snp_map"]
    lines = ["Top_promoter[SNPs_in_top_promoter]\tpromoter_snps_in_region\tchosen_promoter_id\tpromoters\toriginal_p\t"
             "corrected_coexpression_score\traw_p(perm)\tBenjamini-Hochberg\n"]
    fdrstore = {}
    for i, (key, value) in enumerate(sorted(indmeasure.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)):
        label = ji[key]
        snps = []
        for prom in ptj[label]:
            for snp in (snp_map[prom]["snps"].split("|") if prom in snp_map else ["no_snps.seed_promoter"]):
                if snp not in snps:
                    snps.append(snp)
        winner = winners[label]
        winning_snps = snp_map[winner]["snps"] if winner in snp_map else ""
        lines.append("%s [%s]\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (
            winner, winning_snps, "|".join(snps), winner, "|".join(ptj[label]), 1,
            value / len(indmeasure), numbers["p"][i], numbers["fdr_bh"][i]))
        fdrstore[key] = numbers["fdr_bh"][i]
    with open(path, "w") as o:
        o.writelines(lines)
    return fdrstore


def write_allscores(path, real, fdrstore, prom_addresses):
    """2-collate-results.py:264-283: sorted by start, then (stably) by chromosome string."""
    ji, indmeasure = real["joining_instructions"], real["indmeasure"]
    rows = [[prom_addresses[prom][0], prom_addresses[prom][1], prom_addresses[prom][2], prom, ji[prom],
             real["allpromoterscores"][prom], indmeasure[ji[prom]], fdrstore[ji[prom]]] for prom in ji]
    rows.sort(key=lambda x: int(x[1]))
    rows.sort(key=lambda x: x[0])
    with open(path, "w") as o:
        o.write("chrom\tstart\tend\tname\tgroup_name\traw_coex\tgroup_coex\tgroup_FDR\n")
        for row in rows:
            o.write("%s\n" % ("\t".join(str(x) for x in row)))


def _sum(values):
    """Python 2's `sum` of floats: plain left-to-right double additions (Python >= 3.12 compensates)."""
    s = 0
    for v in values:
        s = s + v
    return s


def remove_overlap(thisdict):
    """coexfunctions.py:437-461: {chrom: {start: end}} -> merged, non-overlapping ranges."""
    newdict = {}
    for chrom in thisdict:
        newdict[chrom] = {}
        starts = sorted(thisdict[chrom].keys())
        last_start = starts[0]
        last_end = thisdict[chrom][last_start]
        for this_start in starts[1:]:
            this_end = thisdict[chrom][this_start]
            if this_start <= last_end:
                last_start = min(this_start, last_start)
                last_end = max(this_end, last_end)
            else:
                newdict[chrom][last_start] = last_end
                last_start = this_start
                last_end = this_end
        newdict[chrom][last_start] = last_end
    return newdict


def write_reports(working_files_dir, settings, real, numbers, num_completed_perms, f5resource, feature_dict):
    """The reporting files of 2-collate-results.py:144-158, 165-262 (html), 285-322 (network), 324-380 (layout), 382-430 (stats,
    file list).  -> list of the files written."""
    label = str(settings.get("supplementary_label", "coex"))
    indmeasure, winners, ptj = real["indmeasure"], real["winners"], real["prom_to_join"]
    ji, snp_map, indjoined, individualscores = real["joining_instructions"], real["snp_map"], real["indjoined"], real["individualscores"]
    wd = working_files_dir
    written = []

    def out(name, text):
        with open(os.path.join(wd, name), "w") as o:
            o.write(text)
        written.append(name)

    qqfile = label + ".qq_plot"
    out(qqfile, "rand\t%s\n" % qqfile)                                     # :146-148, the rest of the block raises NameError
    # ---- html table (:170-262) ----
    html = ["<table id=\"results\">\n<tr>\n<th>Top promoter (click for details in new tab)</th>\n<th>SNPs in top_promoter</th>\n"
            "<th>Corrected coexpression score</th>\n<th>FDR</th>\n</tr>"]
    pseudonyms, highlight = {}, []
    for i, (key, value) in enumerate(sorted(indmeasure.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)):
        group_label = ji[key]
        winner = winners[group_label]
        pseudonyms[winner] = winner                                          # (no metadata: `promoter_details` is undefined, :113)
        winning_snps = snp_map[winner]["snps"] if winner in snp_map else ""
        winning_snps = winning_snps.replace("||", "|")
        if winning_snps[-1:] == "|":
            winning_snps = winning_snps[:-1]
        fdr = numbers["fdr_bh"][i]
        if fdr < 0.05:
            html.append("<tr class=\"sig\">\n")
            highlight.append(winner)
            highlight.extend(ptj[group_label])
        else:
            html.append("<tr>\n")
        html.append("<td><a href=\"%s%s\" target=\"_blank\">%s</a></td><td>%s</td><td>%s</td><td>%s</td></tr>\n"
                    % (f5resource, winner, winner, winning_snps.replace("|", ", "), round(value / len(indmeasure), 1), round(fdr, 2)))
    html.append("</table>")
    out(label + ".html", "".join(html))
    # ---- circos network (:285-322): mean of the group-level edges of every pair of hits of different groups ----
    already = {}
    for prom1 in individualscores:
        for prom2 in individualscores[prom1]:
            if prom1 != prom2:
                g1, g2 = ji[prom1], ji[prom2]
                if g1 != g2 and g1 in indjoined and g2 in indjoined[g1]:
                    already.setdefault("\t".join(sorted([prom1, prom2])), []).append(indjoined[g1][g2])
    net = ['<? header("Access-Control-Allow-Origin: https://baillielab.net")?>\nsource\ttarget\tvalue\n']
    for c in already:
        net.append("%s\t%s\n" % (c, float(_sum(already[c])) / len(already[c])))
    out(label + ".network.php", "".join(net))
    # ---- BioLayout / d3 layout file (:324-380) ----
    lf = ["//BioLayout Express 3D Version 2.1  Layout File\n"]
    included = []
    for g1 in indjoined:
        for g2 in indjoined[g1]:
            w1, w2 = winners[ji[g1]], winners[ji[g2]]
            edgelist = [indjoined[g1][g2]]
            if g2 in indjoined and g1 in indjoined[g2]:
                edgelist.append(indjoined[g2][g1])
            edge = float(_sum(edgelist)) / len(edgelist)                      # the two directions differ under nsp
            if edge >= -1:
                lf.append('"%s"\t"%s"\t%s\n' % (pseudonyms[w1], pseudonyms[w2], edge))
                included.append(w1)
                included.append(w2)
    for entry in sorted(set(included)):                                      # (the reference: set iteration order)
        lf.append('//NODECLASS\t"%s"\t%s\t"significantly_coexpressed"\n' % (pseudonyms[entry], "TRUE" if entry in highlight else "FALSE"))
    lf.append('//EDGECOLOR\t"#FF0033"\n')
    lf.append("//EDGESIZE\t2\n")
    lf.append('//NODECLASSCOLOR\t"TRUE"\t"significantly_coexpressed"\t"#0000FF"\n')
    lf.append('//NODECLASSCOLOR\t"FALSE"\t"significantly_coexpressed"\t"#BAAAD1"\n')
    lf.append('//CURRENTCLASSSET\t"significantly_coexpressed"\n')
    lf.append('//DEFAULTSEARCH\t"%s"\n' % f5resource)
    out(label + ".layout", "".join(lf))
    # ---- stats of the real data (:382-418) ----
    feature_label = str(settings["feature_coordinates"]).split("/")[-1].split(".")[0]
    snps_mapped = os.path.join(wd, "permutation_store", feature_label + ".0.bed")
    proms, tss_snps = set(), set()
    with open(snps_mapped) as f:
        for raw in f:
            line = raw.strip().split("\t")
            proms.add(line[-1])
            tss_snps.add(line[3])
    snp_count, genome_length = int(settings["snp_count"]), int(settings["genome_length"])
    hpm_genome = snp_count * 1000000.0 / genome_length * 1.0
    range_length = 0
    merged = remove_overlap(feature_dict)
    for chrom in merged:
        for start in merged[chrom]:
            range_length += merged[chrom][start] - start
    try:
        hpm = (float(len(proms)) * 1000000) / float(range_length)
    except ZeroDivisionError:
        hpm = "div0"
    statsfile = label + "_stats.html"
    out(statsfile, "snps_searched\t%s\ntss_snps\t%s\nhpm_genome\t%s\nhpm_tss\t%s\nnumperms\t%s\nnsp\t%s\nallproms\t%s\ndistinct\t%s\n"
        % (snp_count, len(tss_snps), hpm_genome, hpm, num_completed_perms, settings.get("nsp"), len(real["allpromoterscores"]), len(ptj)))
    out(label + ".filelist", "qqfile\t%s.svg\nlayoutfile\t%s\nstatsfile\t%s\nfullresults\t%s\nhtmlresults\t%s\nemail\t%s\n"
        % (qqfile, label + ".layout", statsfile, label + ".individual_scores", label + ".html", settings.get("useremail")))
    return written


def collate(working_files_dir):
    settings = readsettings(working_files_dir)
    pool, real, n_perm = load_store(os.path.join(working_files_dir, "perm_results_store"))
    numbers = collate_numbers(pool, real)
    label = str(settings.get("supplementary_label", "coex"))
    out = os.path.join(working_files_dir, label + ".individual_scores")
    numbers["fdrstore"] = write_individual_scores(out, real, numbers)
    feat = settings.get("feature_coordinates")
    if feat and os.path.exists(str(feat)):
        prom_addresses, feature_dict, _ = readfeaturecoordinates(str(feat))
        write_allscores(os.path.join(working_files_dir, "allscores.txt"), real, numbers["fdrstore"], prom_addresses)
        f5resource = ""
        cfg = settings.get("config_file")
        if cfg and os.path.exists(str(cfg)):                                 # 2-collate-results.py:52-57
            import configparser
            config = configparser.RawConfigParser(allow_no_value=True)
            config.read(str(cfg))
            if config.has_option("directorypaths", "f5resource"):
                f5resource = config.get("directorypaths", "f5resource")
        try:
            numbers["reports"] = write_reports(working_files_dir, settings, real, numbers, n_perm, f5resource, feature_dict)
        except (KeyError, OSError) as e:                                     # e.g. settings.txt of a hand-made store without snp_count
            numbers["reports_error"] = "%s: %s" % (type(e).__name__, e)
    numbers["num_completed_perms"] = n_perm
    numbers["individual_scores_file"] = out
    return numbers


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-wd", "--working_files_dir", type=str, default="null", help="filepath")
    ap.add_argument("-t", "--numthreads", type=int, default=4, help="accepted and ignored")
    args = ap.parse_args(argv)
    res = collate(args.working_files_dir)
    print("collated {} groups against {} permutations -> {}".format(len(res["p"]), res["num_completed_perms"],
                                                                   res["individual_scores_file"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
