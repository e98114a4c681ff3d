"""TEST INFRASTRUCTURE -- CPU restatement of the reference's permutation generator and locus -> promoter
mapping.  Only tests/ may import this module.

Follows, with explicit orders where the reference iterates Python-2 dicts:
  get_gwad ............. reference coexfunctions.py:102-112   chromosome + address -> genome-wide address
  get_address .......... reference coexfunctions.py:115-128   genome-wide address -> chromosome + address (note the
                         `while gwad > genome_max` wrap and the half-open test `start < gwad <= end`)
  listitem ............. reference coexfunctions.py:319-325   index wrap used by the 'backcirc' mode.  Python 2
                         `round()` rounds halves AWAY from zero, Python 3 to even: restated explicitly.
                         get_gwad / get_address / listitem are PINNED against direct calls of the reference's own
                         functions (tests/golden/reference_functions.json, tests/test_reference_pin.py; listitem on
                         the indices where Python 2 and 3 rounding agree, the half-away rule by its own test).
                         The three modes below are PINNED against the reference's own permutation block
                         (0-prepare-coex.py:471-613) executed with a seeded RNG and a stand-in for the windowBed binary that
                         captures the temporary BED of every set (tests/golden/reference_permutations.json).
  circular mode ........ reference 0-prepare-coex.py:548-563
  backcirc mode ........ reference 0-prepare-coex.py:579-593  (sorted background by (gwad, key), :582)
  background mode ...... reference 0-prepare-coex.py:564-578  (random.sample of background lines; the draw is an input here)
  temporary BED ........ reference 0-prepare-coex.py:598-607  `chrom(without chr) address address+1 name`
  windowBed -w W ....... bedtools (v2.x, README.md:24; NOT under /root/reference and not installed here): published
                         semantics restated -- A is widened by W on both sides (start clamped at 0) and every B with
                         a.start' < b.end and b.start < a.end' is reported.  PARITY UNPINNED against a real bedtools
                         binary (none in this image); the output ORDER of bedtools (bin traversal) is irrelevant to the
                         reference, whose consumer puts the promoters into a dict (1-make-network.py:131-177).
Orders fixed here (and in the CUDA path): loci in input order, features of one locus in ascending table row,
promoters of a permutation in first-appearance order.
"""
from __future__ import annotations

import math

import numpy as np


def chromrange_from_lengths(chromlist, chrom_lengths):
    """0-prepare-coex.py:319-330: consecutive [start, end] in chromlist order."""
    cr, start = {}, 0
    for c in chromlist:
        cr[c] = [start, start + int(chrom_lengths[c])]
        start += int(chrom_lengths[c])
    return cr


def get_gwad(chrom, address, chromrange):
    if chrom not in chromrange:
        return None
    if (chromrange[chrom][0] + address) < chromrange[chrom][1] or address < 0:
        return chromrange[chrom][0] + address
    return None                       # the reference falls off the end of the function


def get_address(gwad, chromrange, chromlist):
    genome_max = chromrange[chromlist[-1]][1]
    while gwad > genome_max:
        gwad = gwad - genome_max
    for chrom in chromlist:
        start, end = chromrange[chrom]
        if gwad > start and gwad <= end:
            return chrom, gwad - start
    return "chr1", 0


def round_half_away(x):
    """Python 2 round(x, 0) == C round(): nearest, halves away from zero (|x| - floor|x| is exact)."""
    y = math.floor(abs(x))
    if abs(x) - y >= 0.5:
        y += 1
    return math.copysign(y, x)


def listitem_index(n, index):
    a = float(index) / n
    b = a - int(round_half_away(a))
    c = b * n
    return int(round_half_away(c))     # may be negative: Python indexes from the end


def circular_positions(snps, shift, chromrange, chromlist):
    """snps: list of (chrom, address).  -> list of (chrom, address) in input order (None where get_gwad fails)."""
    out = []
    for chrom, address in snps:
        g = get_gwad(chrom, address, chromrange)
        if g is None:
            out.append(None)
            continue
        out.append(get_address(g + shift, chromrange, chromlist))
    return out


def sorted_background(backdict):
    """0-prepare-coex.py:582: keys sorted by (gwad, key)."""
    return [k for k, _ in sorted(backdict.items(), key=lambda kv: (kv[1], kv[0]))]


def backcirc_positions(snps, shift, sortedbackground):
    """snps: list of (chrom, address); every one of them is a key `chrom_address` of the background."""
    pos = {k: i for i, k in enumerate(sortedbackground)}
    out = []
    n = len(sortedbackground)
    for chrom, address in snps:
        idx = listitem_index(n, pos["%s_%s" % (chrom, address)] + shift)
        new = sortedbackground[idx].split("_")
        out.append((new[0], int(new[1])))
    return out


def background_positions(backlines, picks):
    """0-prepare-coex.py:564-578: backlines = the background file's lines split on tabs, picks = the line numbers
    random.sample drew.  chrom = column 0 without `chr` (the `chr` is put back here: window_bed matches feature
    chromosomes WITH the prefix, as the feature BED is written without it on both sides), address = column 2."""
    out = []
    for linenum in picks:
        line = backlines[linenum]
        out.append(("chr" + line[0].replace("chr", ""), int(line[2])))
    return out


def window_bed(positions, feat_chrom, feat_start, feat_end, window):
    """positions: list of (chrom, address) or None; features as parallel arrays (chrom strings WITH the `chr`
    prefix).  A = [address, address + 1).  -> list of (locus index, feature row)."""
    feat_chrom = np.asarray(feat_chrom)
    feat_start = np.asarray(feat_start, dtype=np.int64)
    feat_end = np.asarray(feat_end, dtype=np.int64)
    out = []
    for k, p in enumerate(positions):
        if p is None:
            continue
        chrom, a = p
        a0 = max(0, a - window)
        a1 = a + 1 + window
        ok = (feat_chrom == chrom) & (feat_end > a0) & (feat_start < a1)
        for r in np.nonzero(ok)[0]:
            out.append((k, int(r)))
    return out


def hit_rows(pairs):
    seen, rows = set(), []
    for _, r in pairs:
        if r not in seen:
            seen.add(r)
            rows.append(r)
    return rows
