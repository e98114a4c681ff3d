"""Generates the committed golden vectors under tests/golden/.

Run HERE (the build container), where /root/reference exists:
    make -C oracle && python tests/golden/make_golden.py

  getpearson_*.npz  inputs + outputs of the UNMODIFIED reference C file
                    (/root/reference/coexpression_v2.c built by oracle/Makefile `ref`)
  ranks_*.npz       scipy.stats.rankdata outputs (the reference's rank transform,
                    1-make-network.py:110-111)
  network_*.json/.npz  outputs of oracle/network_oracle.py (consistent mode) DRIVING that
                    reference C build -- the Python drivers of the reference are Python-2
                    only and cannot run in this image, so these pin the restatement, not
                    executed reference Python.
Nothing on the GPU box reads /root/reference: tests only read the files written here.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from coexpression_b200 import synth  # noqa: E402
from oracle import network_oracle as no  # noqa: E402

assert "reference" in no.available_kinds(), "build oracle/_ref/coexpression_v2.so first (make -C oracle)"


def getpearson_vectors():
    rng = np.random.default_rng(1234)
    t = synth.make_table(300, 37, seed=11)
    X = t.X.copy()
    X[5] = 3.25                 # constant non-zero row -> 0/0 NaN (coexpression_v2.c:113)
    X[6] = -X[7]                # mixed-sign rows
    X[8, :] = 0.0; X[8, 0] = 1.0; X[8, 1] = -1.0   # in-order sum exactly 0 -> NaN (:89-92)
    ia = rng.integers(0, 300, 4000).astype(np.int32)
    ib = rng.integers(0, 300, 4000).astype(np.int32)
    ia[:40] = np.arange(40); ib[:40] = np.arange(40)          # self pairs
    ia[40:60] = [5, 6, 7, 8] * 5; ib[40:60] = np.arange(20)
    R = no.rank_rows(X)
    np.savez_compressed(os.path.join(HERE, "getpearson_raw.npz"), X=X, ia=ia, ib=ib,
                        r=no.get_pearson(X, ia, ib, "reference"))
    np.savez_compressed(os.path.join(HERE, "getpearson_ranked.npz"), X=R, ia=ia, ib=ib,
                        r=no.get_pearson(R, ia, ib, "reference"))
    # edge sizes the reference call sites produce (0-prepare-coex.py:641-642,687): 0 and 1 pair
    np.savez_compressed(os.path.join(HERE, "getpearson_onepair.npz"), X=X, ia=ia[:1], ib=ib[:1],
                        r=no.get_pearson(X, ia[:1], ib[:1], "reference"))
    # ranks incl. heavy ties, all-equal row, negative zero
    Y = X[:64].copy()
    Y[0] = 0.0
    Y[1, ::2] = -0.0; Y[1, 1::2] = 0.0
    Y[2] = np.arange(37)[::-1]
    np.savez_compressed(os.path.join(HERE, "ranks_small.npz"), X=Y, ranks=no.rank_rows(Y))


NETWORK_CASES = {
    # name: (shape, n_loci, window, kwargs)
    "spearman_default": ("tiny", 40, 20000, {}),
    "spearman_anticor": ("tiny", 40, 20000, {"anticorrelation": True}),
    "spearman_noiter_j001": ("tiny", 40, 20000, {"noiter": True, "pval_to_join": 0.01}),
    "spearman_useavg_m": ("tiny", 40, 20000, {"useaverage": True, "selectionthreshold_is_j": True}),
    "spearman_global_nsp0": ("tiny", 40, 20000, {"nsp": 0.0}),
    "pearson_default": ("tiny", 40, 20000, {"correlationmeasure": "Pearson"}),
    "pearson_anticor_j05": ("tiny", 40, 20000, {"correlationmeasure": "Pearson", "anticorrelation": True,
                                               "pval_to_join": 0.5}),
}


def network_inputs(shape, n_loci, window, cm):
    N, n = synth.SHAPES[shape]
    t = synth.make_table(N, n, seed=21)
    hits, loci, m = synth.make_permutation_hits(t, n_loci, 2, window=window, seed=5)
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 30000, seed=6)
    g = no.get_pearson(R if cm == "Spearman" else t.X, ga, gb, "reference")
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    return t, hits, R, glist


def network_vectors():
    cache = {}
    for name, (shape, n_loci, window, kw) in NETWORK_CASES.items():
        cm = kw.get("correlationmeasure", "Spearman")
        key = (shape, n_loci, window, cm)
        if key not in cache:
            cache[key] = network_inputs(*key)
        t, hits, R, glist = cache[key]
        perms = []
        for k, h in enumerate(hits):
            res = no.make_network(t.names, t.X, [t.names[r] for r in h], t.addresses(), glist.tolist(),
                                  mode="consistent", kind="reference", Xrank=R, return_counts=True, **kw)
            perms.append({
                "hit_rows": [int(r) for r in h],
                "indmeasure": res["indmeasure"],
                "prom_to_join": res["prom_to_join"],
                "winners": res["winners"],
                "allpromoterscores": res["allpromoterscores"],
                "counts": res["_counts"].tolist(),
            })
        with open(os.path.join(HERE, f"network_{name}.json"), "w") as o:
            json.dump({"shape": shape, "table_seed": 21, "n_loci": n_loci, "window": window, "kwargs": kw,
                       "perms": perms}, o)
    # the shared inputs, so tests do not depend on numpy's RNG streams staying stable
    for (shape, n_loci, window, cm), (t, hits, R, glist) in cache.items():
        np.savez_compressed(os.path.join(HERE, f"network_inputs_{shape}_{cm}.npz"), X=t.X,
                            names=np.array(t.names), chrom=t.chrom, start=t.start, end=t.end, glist=glist)


if __name__ == "__main__":
    getpearson_vectors()
    network_vectors()
    print("golden vectors written to", HERE)
