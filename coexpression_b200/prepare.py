"""B200 drop-in for the reference's 0-prepare-coex.py: same command line (flags of 0-prepare-coex.py:23-73), same
working directory (`<resdir><label>/` with settings.txt, permutation_store/<feat>.<k>.bed, perm_results_store/,
jobfile.txt, collationcommandfile.txt, coex.queue -- 0-prepare-coex.py:156-181, 443-453, 889-900) and the same global
correlation list file next to the expression table (:114-118, 621-716).

What changes is who does the work:
  * the K permuted locus sets and their `windowBed -w` joins (0-prepare-coex.py:471-613, one subprocess per
    permutation) are ONE device call, CoexContext.permute_map (csrc/permute.cuh); the map files are still written in
    windowBed's line format so that `1-make-network.py -mf ... -wd ...`, resume (-z) and 2-collate keep working;
  * the global correlation list goes through the resident-table pair kernel (coexfunctions.make_correlation_list);
  * unless -po, the run itself is `make_network.run_batch` (every map file in one process, table resident on the GPU)
    followed by `collate.collate`, instead of runlocal.py's pool of 1-make-network.py processes.

Not supported (fails loudly): `-p post` (rejection sampling around join_nearby driven by an unseeded RNG,
0-prepare-coex.py:723-871 -- host control flow, not a data path) and bonus gene / promoter names in the input file,
which force that mode (:406-430); `-seed` (obsolete, 1-make-network.py:82-97).
"""
from __future__ import annotations

import argparse
import configparser
import json
import os
import random
import sys

import numpy as np

from .coexfunctions import getversion, make_correlation_list, read_expression_file, readfeaturecoordinates

version = getversion()


def build_parser(coexappdir):
    """The flags of reference 0-prepare-coex.py:23-73, names, defaults and types unchanged."""
    parser = argparse.ArgumentParser()
    parser.add_argument('-f', '--snp_details_file', default='unset', help='input filepath')
    parser.add_argument('-po', '--prepareonly', action="store_true", default=False, help='Stop after setup script')
    parser.add_argument('-n', '--numperms', type=int, default=100, help='window size')
    parser.add_argument('-p', '--permutationmode', type=str, choices=['circular', 'post', 'backcirc', 'background'],
                        default='circular', help='permutation mode')
    parser.add_argument('-s', '--nsp', type=float, default=1.0,
                        help='node-specific pvalue in use; specify precision (1=perfect, 0=globaldistribution is used.)')
    parser.add_argument('-t', '--numthreads', type=int, default=1, help='numthreads used')
    parser.add_argument('-w', '--window', type=int, default=0, help='window size')
    parser.add_argument('-b', '--backgroundfile', default='unset', help='backgroundfile')
    parser.add_argument('-q', '--feature_coordinates', default="SFD:f5ep300_100.bed",
                        help='SORTED bedfile for the expression map. address of each locus is in expression file.')
    parser.add_argument('-x', '--expression_file', default="SFD:f5ep_ptt_condense.expression",
                        help='expression file for the expression map')
    parser.add_argument('-me', '--metadatafile', default="SFD:METADATA_U22_ENHANCERS", help='metadatafile')
    parser.add_argument('-cl', '--chrom_lengths_file', default="SFD:chromlengths.txt", help='chrom_lengths_file')
    parser.add_argument('-cc', '--config_file', default=os.path.join(coexappdir, 'app.cfg'),
                        help='use a custom config file instead of app.cfg')
    parser.add_argument('-o', '--outputdirname', default="auto", help='fullname of outputdir')
    parser.add_argument('-e', '--useremail', default="no_email", help='user email')
    parser.add_argument('-y', '--version_requested', default=version, help='version of the software the user is expecting')
    parser.add_argument('-gwd', '--get_wd_only', action="store_true", default=False, help='returns the working directory path only')
    parser.add_argument('-v', '--verbose', action="store_true", default=False, help='increases verbosity')
    parser.add_argument('-j', '--pval_to_join', type=float, default=0.1, help='pval_to_join')
    parser.add_argument('-u', '--useaverage', action="store_true", default=False,
                        help='use the average coexpression score (eliminates position enrichment signal)')
    parser.add_argument('-z', '--doitagain', action="store_true", default=False, help='create a whole new permutations set')
    parser.add_argument('-cm', '--correlationmeasure', type=str, choices=['Spearman', 'Pearson'], default='Spearman',
                        help='filepath')
    parser.add_argument('-gpl', '--gpl_len', type=int, default=1000000,
                        help='number of correlations to include in the global correlation lists.')
    parser.add_argument('-ss', '--soft_separation', type=int, default=100000,
                        help='correlating promoters in this range will be joined.')
    parser.add_argument('-seed', '--seedfile', default="none", help='seed file')
    parser.add_argument('-a', '--anticorrelation', action="store_true", default=False, help='anticorrelation')
    parser.add_argument('-i', '--noiter', action="store_true", default=False, help='noiter')
    parser.add_argument('-r', '--recordedges', action="store_true", default=False, help='recordedges')
    parser.add_argument('-m', '--selectionthreshold_is_j', action="store_true", default=False,
                        help='match selectionthreshold_is_j')
    parser.add_argument('-st', '--selectionthreshold', type=float, default=0, help='selectionthreshold')
    # ---- B200 extensions (absent from the reference; none of them changes a default behaviour) ----
    parser.add_argument('--device', type=int, default=0, help='B200 extension: CUDA device')
    parser.add_argument('--rng_seed', type=int, default=None,
                        help='B200 extension: seed of the permutation RNG (the reference is unseeded, 0-prepare-coex.py:233-234)')
    parser.add_argument('--lean_results', action="store_true", default=False,
                        help='B200 extension: write the two P x P result objects (individualscores, indjoined) for the real data '
                             'only; 2-collate-results.py reads nothing else of a permuted set than indmeasure')
    parser.add_argument('--cache', action="store_true", default=False,
                        help='B200 extension: binary cache of the parsed expression table')
    return parser


def writequeue(line, qfile):
    """coexfunctions.py:511-518: timestamped progress lines."""
    import datetime
    with open(qfile, "a") as o:
        o.write("%s\t%s\n" % (datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"), line))


def read_chrom_lengths(path):
    """0-prepare-coex.py:254-261: `chrN: length` lines."""
    lengths, genome = {}, 0
    with open(path) as f:
        for raw in f:
            line = raw.strip().split(": ")
            if len(line) < 2:
                continue
            lengths[line[0]] = int(float(line[1]))
            genome += int(float(line[1]))
    return lengths, genome


def read_snps(path, chrom_lengths, verbose=False):
    """0-prepare-coex.py:265-309.  -> (snp_bed_data {chrom: {address: snp}} in insertion order, all_snps, badlines)."""
    snp_bed_data, all_snps, badlines = {}, [], []
    with open(path) as f:
        lines = f.readlines()
    for line in lines:
        if line.startswith("#"):
            continue
        if "\t" in line:
            line = line.strip().split("\t")
        elif " " in line:
            line = line.strip().split(" ")
        else:
            line = line.strip().split()
        try:
            address = int(float(line[1]))
        except (IndexError, ValueError):
            if verbose:
                print("skipping line of snp_details file", line)
            if line:
                badlines.append(line[0])
            continue
        if len(line) > 1:
            try:
                snp = line[3]
            except IndexError:
                snp = "_".join(line[:1])
            chrom = "chr%s" % line[0].replace("chr", "")
            if chrom not in chrom_lengths:
                if verbose:
                    print("skipping", chrom, "as length not known")
                continue
            if snp not in all_snps:
                all_snps.append(snp)
            snp_bed_data.setdefault(chrom, {})[address] = snp
    return snp_bed_data, all_snps, badlines


def chromosome_ranges(chrom_lengths):
    """0-prepare-coex.py:319-330: chr1..22, X, Y, M laid end to end."""
    chromrange, chromlist, start = {}, [], 0
    for chromnum in list(range(1, 23)) + ['X', "Y", "M"]:
        chrom = "chr%s" % chromnum
        chromlist.append(chrom)
        thislen = chrom_lengths[chrom]
        chromrange[chrom] = [start, start + thislen]
        start += thislen
    return chromrange, chromlist


def output_label(a, snp_details_file):
    """0-prepare-coex.py:120-144."""
    if a.outputdirname != "auto":
        return a.outputdirname
    label = os.path.split(snp_details_file)[1].replace(".bed", "")
    label += "_complete" if a.nsp == 1 else "_s%s" % a.nsp
    label += {"post": "_POST", "backcirc": "_BACKCIRC", "circular": "_CIRCULAR", "background": "_BACKGROUND"}[a.permutationmode]
    if a.selectionthreshold_is_j:
        label += "_SIJ"
    if a.anticorrelation:
        label += "_ANTIin"
    if a.noiter:
        label += "_NOiter"
    if a.useaverage:
        label += "_useav"
    if a.window > 0:
        label += "_w%s" % a.window
    label += "_pj%s" % a.pval_to_join
    label += "_%s" % os.path.split(a.expression_file)[-1].replace('.expression', '')
    return label


def write_map_file(path, hit_rows, hit_loci, which, loci, snp_names, table_names, addresses, rand_prefix):
    """One windowBed-style line per promoter, in the promoters' first-appearance order: A columns
    `chrom(without chr) address address+1 snp` + B columns `chrom start end label`; the consumer reads column 3 and
    the last column (1-make-network.py:137-147).  which = 1: the LAST locus mapping to the promoter, as the reference's
    rewrite of the real-data file keeps it (0-prepare-coex.py:396-399, 432-438); 0: the first (permuted sets, whose SNP
    names nothing reads)."""
    with open(path, "w") as o:
        for row, lo in zip(hit_rows, hit_loci):
            m = int(lo[which])
            chrom, address = loci[m]
            label = table_names[int(row)]
            bc, bs, be = addresses[label]
            o.write("%s\t%s\t%s\t%s%s\t%s\t%s\t%s\t%s\n" % (chrom.replace("chr", ""), address, int(address) + 1, rand_prefix,
                                                       snp_names[m], bc, bs, be, label))


def main(argv=None, coexappdir=None):
    coexappdir = coexappdir or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    parser = build_parser(coexappdir)
    a = parser.parse_args(argv)
    verbose = a.verbose
    if verbose and not a.get_wd_only:
        for name, value in vars(a).items():
            print((name, value))
    # ---- app.cfg (0-prepare-coex.py:82-94) ----------------------------------------------------------------------
    config = configparser.RawConfigParser(allow_no_value=True)
    config.read(a.config_file)
    sourcefilesdir = config.get('directorypaths', 'sourcefilesdir')
    resdir = config.get('directorypaths', 'resdir')
    pathtopython = config.get('directorypaths', 'pathtopython')
    if not os.path.isabs(sourcefilesdir):
        sourcefilesdir = os.path.join(coexappdir, sourcefilesdir)
    if not os.path.isabs(resdir):
        resdir = os.path.join(coexappdir, resdir)
    for key in ("feature_coordinates", "expression_file", "metadatafile", "chrom_lengths_file"):
        v = getattr(a, key)
        if v.startswith("SFD:"):
            setattr(a, key, os.path.join(sourcefilesdir, v.replace("SFD:", "")))
    snp_details_file = a.snp_details_file
    if snp_details_file == 'unset':
        snp_details_file = os.path.join(sourcefilesdir, 'test.bed')
    expfilename = a.expression_file.split("/")[-1]
    ext = ".correlationlist" if a.correlationmeasure == "Pearson" else ".spearmanlist"
    correlationfile = os.path.join(sourcefilesdir, expfilename) + ext
    label = output_label(a, snp_details_file)
    supplementary_label = 'coex'
    working_files_dir = "%s%s" % (resdir, label)
    jobfile = os.path.join(working_files_dir, "jobfile.txt")
    queuefile = os.path.join(working_files_dir, "coex.queue")
    collationcommandfile = os.path.join(working_files_dir, "collationcommandfile.txt")
    storage_dir_permutations = os.path.join(working_files_dir, "permutation_store")
    storage_dir_perm_results = os.path.join(working_files_dir, "perm_results_store")
    settingsfile = os.path.join(working_files_dir, "settings.txt")
    path_to_makenet = os.path.join(coexappdir, "1-make-network.py")
    path_to_collationfile = os.path.join(coexappdir, "2-collate-results.py")
    if a.get_wd_only or verbose:
        print(working_files_dir)
    if a.get_wd_only:
        return 0
    if version != a.version_requested:
        print("Running version {} but version {} was explicitly requested".format(version, a.version_requested))
        return 1
    if a.permutationmode == 'post':
        raise SystemExit("-p post is not supported by the B200 drop-in (see coexpression_b200/prepare.py); use circular, "
                         "backcirc or background")
    if a.seedfile != "none":
        raise SystemExit("-seed is obsolete in the reference and not supported by the B200 drop-in")
    for d in (resdir, working_files_dir, storage_dir_permutations, storage_dir_perm_results):
        os.makedirs(d, exist_ok=True)
    runcommand = "{} {} -y {} -f {} -n {} -p {} -s {} -e {} -t {} -x {} -q {} -w {} -o {}".format(
        pathtopython, os.path.join(coexappdir, '0-prepare-coex.py'), version, snp_details_file, a.numperms, a.permutationmode,
        a.nsp, a.useremail, a.numthreads, a.expression_file, a.feature_coordinates, a.window, a.outputdirname)
    runcommand += " -j %s" % a.pval_to_join
    for flag, on in (("-v", verbose), ("-a", a.anticorrelation), ("-r", a.recordedges), ("-m", a.selectionthreshold_is_j),
                     ("-u", a.useaverage)):
        if on:
            runcommand += " " + flag
    feature_label = a.feature_coordinates.split("/")[-1].split(".")[0]
    snps_mapped = os.path.join(storage_dir_permutations, feature_label + ".0.bed")

    def start_run():
        if a.prepareonly:
            if verbose:
                print("Stopping after preparation")
            return 0
        from . import collate, make_network
        make_network.run_batch(working_files_dir, None, a.device, verbose, cache=a.cache, lean=a.lean_results)
        collate.collate(working_files_dir)
        return 0

    if os.path.exists(collationcommandfile) and not a.doitagain:            # 0-prepare-coex.py:246-251
        permfilecount = len(os.listdir(storage_dir_permutations))
        if permfilecount >= a.numperms:
            print("\n{} existing permutations found in:\n {}\nStopping preparations (use -z to repeat preparation)\n".format(
                permfilecount, storage_dir_permutations))
            return start_run()

    chrom_lengths, genome_length = read_chrom_lengths(a.chrom_lengths_file)
    print("genome read, length =", genome_length)
    snp_bed_data, all_snps, badlines = read_snps(snp_details_file, chrom_lengths, verbose)
    snp_count = len(all_snps)
    chromrange, chromlist = chromosome_ranges(chrom_lengths)
    prom_addresses, _, sorted_feature_list = readfeaturecoordinates(a.feature_coordinates)
    if badlines:
        known = [nm for nm in badlines if nm in prom_addresses]
        if known or os.path.exists(a.metadatafile):
            raise SystemExit("the input names genes / promoters (%s ...): the reference then switches to -p post "
                             "(0-prepare-coex.py:406-418), which the B200 drop-in does not support" % badlines[0])

    # ---- the table, resident on the GPU for everything that follows ------------------------------------------------
    from .api import CoexContext
    table, _ = read_expression_file(a.expression_file, cache=a.cache)
    names = table.names
    missing = [nm for nm in names if nm not in prom_addresses]
    if missing:                                                              # 0-prepare-coex.py:341-345
        print("bed coordinate datafile \n({}) is missing for some feature IDs in expression file \n({}):\n {}".format(
            a.feature_coordinates, a.expression_file, ', '.join(missing[:20])))
        return 1
    bed_line = {lab: i for i, lab in enumerate(sorted_feature_list)}
    rng = random.Random(a.rng_seed) if a.rng_seed is not None else random.Random()
    loci = [(chrom, address) for chrom in snp_bed_data for address in snp_bed_data[chrom]]      # the reference's A-file order
    snp_names = [snp_bed_data[c][ad] for c, ad in loci]
    ctx = CoexContext(a.device)
    try:
        ctx.set_expression(table.values)
        ctx.prepare()
        ctx.set_features(np.array([prom_addresses[nm][0] for nm in names]),
                         np.array([prom_addresses[nm][1] for nm in names], dtype=np.int64),
                         np.array([prom_addresses[nm][2] for nm in names], dtype=np.int64), names)
        ctx.set_feature_order(np.array([bed_line[nm] for nm in names], dtype=np.int32))
        # ---- real data (0-prepare-coex.py:347-369, 394-438) ---------------------------------------------------------
        rows0, loci0 = ctx.permute_map("circular", [c for c, _ in loci], [ad for _, ad in loci], [0], chromlist, chromrange,
                                       a.window, with_loci=True)
        write_map_file(snps_mapped, rows0[0], loci0[0], 1, loci, snp_names, names, prom_addresses, "")
        numhits = len(rows0[0])
        settingsdict = {k: v for k, v in vars(a).items() if k not in ("device", "rng_seed", "cache", "lean_results")}
        settingsdict.update(version=version, genome_length=genome_length, snp_count=snp_count, numhits=numhits,
                            runcommand=runcommand, correlationfile=correlationfile, supplementary_label=supplementary_label)
        # (selectionthreshold goes in as given: the reference dumps vars(args), and 1-make-network.py:95-96 applies -m itself)
        with open(settingsfile, "w") as o:
            json.dump(settingsdict, o, indent=4, sort_keys=True)
        if numhits < 2:
            writequeue('Not enough of these SNPs mapped to FANTOM5 promoters.', queuefile)
            print("Not enough snps mapped - if this isn't expected, could there be a chr in the bedfile?")
            return 1
        # ---- permuted sets (0-prepare-coex.py:471-613) ----------------------------------------------------------------
        writequeue('Starting to make permutations ({})'.format(a.permutationmode), queuefile)
        shifts = []
        for i in range(a.numperms):
            shifts.append(i + 1 if a.permutationmode == 'background' else int(rng.random() * 10000000000))
        mapfiles = [os.path.join(storage_dir_permutations, feature_label + ".%s.bed" % (i + 1)) for i in range(a.numperms)]
        todo = [i for i in range(a.numperms) if a.doitagain or not os.path.exists(mapfiles[i])]
        background = picks = None
        if a.permutationmode in ('background', 'backcirc'):
            with open(a.backgroundfile) as f:
                backlines = [ln.strip().split('\t') for ln in f.readlines()]
        if a.permutationmode == 'backcirc':
            backdict = {}
            for line in backlines:                                           # 0-prepare-coex.py:499-518
                chrom = 'chr' + line[0].replace('chr', '')
                address = int(float(line[1]))
                key = "%s_%s" % (chrom, address)
                if key in backdict or chrom not in chromrange:
                    continue
                if (chromrange[chrom][0] + address) < chromrange[chrom][1] or address < 0:
                    backdict[key] = chromrange[chrom][0] + address
            for chrom, address in loci:                                      # :520-528: the real hits too
                backdict.setdefault("%s_%s" % (chrom, address), chromrange[chrom][0] + address)
            sb = [k for k, _ in sorted(backdict.items(), key=lambda kv: (kv[1], kv[0]))]
            background = [(k.split("_")[0], int(k.split("_")[1])) for k in sb]
        if a.permutationmode == 'background':
            background = [("chr" + ln[0].replace("chr", ""), int(ln[2])) for ln in backlines]       # :568-570
            picks = [rng.sample(range(len(backlines)), snp_count) if i in todo else [0] * snp_count for i in range(a.numperms)]
        if todo:
            sh = [shifts[i] for i in todo]
            pk = [picks[i] for i in todo] if picks is not None else None
            rows, lo = ctx.permute_map(a.permutationmode, [c for c, _ in loci], [ad for _, ad in loci], sh, chromlist,
                                       chromrange, a.window, background=background, picks=pk, with_loci=True)
            for j, i in enumerate(todo):
                # (A columns of a permuted line: the locus' ORIGINAL position with the reference's `rand_` name prefix; no
                #  consumer reads them -- 1-make-network.py:137-147 takes column 3 and the last column)
                write_map_file(mapfiles[i], rows[j], lo[j], 0, loci, snp_names, names, prom_addresses, "rand_")
        # ---- global correlation list (0-prepare-coex.py:621-716) ------------------------------------------------------
        make_correlation_list(ctx, names, correlationfile, a.correlationmeasure, a.gpl_len, seed=a.rng_seed)
    finally:
        ctx.close()
    # ---- job files (0-prepare-coex.py:889-900) ---------------------------------------------------------------------
    jobs = ["{} {} -mf {} -wd {}".format(pathtopython, path_to_makenet, mf, working_files_dir) for mf in mapfiles + [snps_mapped]]
    with open(jobfile, 'w') as o:
        o.write("{}".format('\n'.join(jobs[::-1])))                         # the real data first
    with open(collationcommandfile, 'w') as o:
        o.write("{} {} -wd {}".format(pathtopython, path_to_collationfile, working_files_dir))
    return start_run()


if __name__ == "__main__":
    sys.exit(main())
