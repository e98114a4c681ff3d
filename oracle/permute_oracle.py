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
  windowBed -w W ....... bedtools (v2.x, README.md:24; NOT under /root/reference and not installed here): restated from the
                         published bedtools2 source -- windowBed.cpp BedWindow::FindWindowOverlaps (A widened by W on both
                         sides, start clamped at 0; a B record is reported iff min(a.end', b.end) - max(a.start', b.start) > 0)
                         and bedFile.cpp BedFile::allHits / bedFile.h getBin (UCSC binning: 16 kb bins, 7 levels of x8; the
                         B records of one A record come out by level from the finest up, bin ascending, B-file order inside
                         a bin).  PARITY UNPINNED against a real bedtools binary (none in this image).  The ORDER is irrelevant
                         to the Python-2 reference, whose consumer puts the promoters into a hash-ordered dict
                         (1-make-network.py:131-177), but it is the order an insertion-ordered (Python 3) run of the reference
                         gives to a permutation's promoters, so this repo adopts it.
Orders fixed here (and in the CUDA path): loci in input order, features of one locus in bedtools' reporting order,
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


def bedtools_bin(start, end):
    """UCSC binning scheme as bedtools keeps its B file in (bedtools2 src/utils/bedFile/bedFile.h: _binFirstShift = 14,
    _binNextShift = 3, _binLevels = 7; getBin): (level, bin number inside the level) of the smallest bin containing
    [start, end - 1]."""
    s, e = start >> 14, (max(end, start + 1) - 1) >> 14
    level = 0
    while level < 6 and s != e:
        s >>= 3
        e >>= 3
        level += 1
    return level, s


def window_bed(positions, feat_chrom, feat_start, feat_end, window, bed_index=None):
    """`bedtools window -w W -a A -b B` restated (bedtools2 windowBed.cpp BedWindow::FindWindowOverlaps + bedFile.cpp
    BedFile::allHits).  positions: list of (chrom, address) or None = the A records in file order, A = [address,
    address + 1) widened by W on both sides (start clamped at 0); features as parallel arrays (chrom strings WITH the
    `chr` prefix) = the B records, bed_index[row] = line number of the feature in the B file (default: row order).
    A B record is reported iff min(a.end', b.end) - max(a.start', b.start) > 0.  ORDER of the B records of one A record:
    allHits walks the bin levels from the finest (16 kb) upwards, the bins of a level in ascending order and the
    records of a bin in B-file order, i.e. ascending (level, bin, line) -- a feature straddling a 16 kb boundary is
    reported after every finer-bin feature of the window.  The order matters: it is the order of the map file's lines,
    hence of the permutation's promoters (1-make-network.py:131-177) and of quirk Q1's skipped column.
    UNPINNED: no bedtools binary in the build image; restated from the published source.  -> list of (locus index,
    feature row)."""
    feat_chrom = np.asarray(feat_chrom)
    feat_start = np.asarray(feat_start, dtype=np.int64)
    feat_end = np.asarray(feat_end, dtype=np.int64)
    if bed_index is None:
        bed_index = np.arange(len(feat_start))
    out = []
    for k, p in enumerate(positions):
        if p is None:
            continue
        chrom, a = p
        a0 = max(0, a - window)
        a1 = a + 1 + window
        ok = (feat_chrom == chrom) & (feat_end > a0) & (feat_start < a1)
        rows = [int(r) for r in np.nonzero(ok)[0]]
        rows.sort(key=lambda r: bedtools_bin(int(feat_start[r]), int(feat_end[r])) + (int(bed_index[r]), r))
        for r in rows:
            out.append((k, r))
    return out


def hit_rows(pairs):
    seen, rows = set(), []
    for _, r in pairs:
        if r not in seen:
            seen.add(r)
            rows.append(r)
    return rows
