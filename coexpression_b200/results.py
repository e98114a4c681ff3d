"""Turns the flat arrays of coex_network_batch back into the eight objects the reference dumps
per permutation (1-make-network.py:523-532), keyed by region labels."""
from __future__ import annotations

import numpy as np


def perm_objects(res, k, names, with_scores=True):
    """res: api.BatchResult; names: table row labels.  Key orders follow the explicit orders
    of the restatement (groups ascending by label, members ascending by name)."""
    p = res.perm(k)
    hit_names = [names[r] for r in p["hit_rows"]]
    G = len(p["label"])
    labels = [hit_names[i] for i in p["label"]]
    members = [[] for _ in range(G)]
    if G == 0:      # fewer than two hits: individualscores is empty, nothing is grouped (1-make-network.py:333-338)
        out = dict(indmeasure={}, prom_to_join={}, joining_instructions={}, winners={}, allpromoterscores={})
        if with_scores and res.scores is not None:
            out["individualscores"] = {}
            out["indjoined"] = {}
        return out
    for i, g in enumerate(p["group_of_hit"]):
        members[g].append(hit_names[i])
    prom_to_join = {labels[g]: sorted(members[g]) for g in range(G)}
    joining_instructions = {nm: labels[g] for nm, g in zip(hit_names, p["group_of_hit"])} if G else {}
    winners = {labels[g]: hit_names[p["winner"][g]] for g in range(G)}
    indmeasure = {labels[g]: float(p["density"][g]) for g in range(G)}
    allpromoterscores = {nm: float(s) for nm, s in zip(hit_names, p["hit_score"])} if G else {}
    out = dict(indmeasure=indmeasure, prom_to_join=prom_to_join, joining_instructions=joining_instructions,
               winners=winners, allpromoterscores=allpromoterscores)
    if with_scores and res.scores is not None:
        S = res.score_matrix(k)
        P = len(hit_names)
        out["individualscores"] = {hit_names[i]: {hit_names[j]: float(S[i, j]) for j in range(P) if j != i}
                                   for i in range(P)} if P > 1 else {}
        w = p["winner"]
        out["indjoined"] = {labels[g]: {labels[h]: float(S[w[g], w[h]]) for h in range(G) if h != g}
                            for g in range(G)}
    return out


def pooled_permuted_densities(res, skip_first=True):
    """perm_indm of 2-collate-results.py:68-76: every permuted density, k != 0."""
    vals = []
    for k in range(1 if skip_first else 0, len(res.n_groups)):
        vals.append(res.perm(k)["density"])
    return np.concatenate(vals) if vals else np.zeros(0)
