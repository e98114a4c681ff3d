"""Row N3 of SURVEY.md section 8f: the arithmetic of the reference's 2-collate-results.py and its main table.

  pooled null ........ every permuted `indmeasure` value, k != 0          2-collate-results.py:67-76
  real measures ...... [indmeasure[g] for g in indjoined]                 :142
  empirical p ........ 1 - bisect_left(sorted(pool), x) / len             :161, coexfunctions.py:379-383
  FDR ................ Benjamini-Yekutieli ('negcorr') and BH ('indep')   :162-163, coexfunctions.py:333-358
  table .............. `<label>.individual_scores`: groups by (score, name) descending, the i-th largest score is
                       paired with the i-th smallest p (:176-244); reported score = indmeasure / len(indmeasure) (:242)

  allscores.txt ...... one line per hit: address, group, allpromoterscores, group density, group FDR (:264-283)

The HTML / BioLayout / network writers of the reference (:165-430) are reporting only and are not rebuilt.
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


def collate(working_files_dir):
    settings = readsettings(working_files_dir)
    pool, real, n_perm = load_store(os.path.join(working_files_dir, "perm_results_store"))
    numbers = collate_numbers(pool, real)
    label = str(settings.get("supplementary_label", "coex"))
    out = os.path.join(working_files_dir, label + ".individual_scores")
    numbers["fdrstore"] = write_individual_scores(out, real, numbers)
    feat = settings.get("feature_coordinates")
    if feat and os.path.exists(str(feat)):
        write_allscores(os.path.join(working_files_dir, "allscores.txt"), real, numbers["fdrstore"],
                        readfeaturecoordinates(str(feat))[0])
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
