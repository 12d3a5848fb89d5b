"""How many global-table insertions could a CTA-local (shared-memory) pre-dedup stage remove?
CPU analysis with the oracle (no GPU needed): for one ogbn-products-shaped mini-batch, fan-out
(15,10,5), the candidates of every hop in the order the sampling kernel visits them
(target-major = virtual position i*k + j) are cut into groups of 32 (a warp), 256, 2048 (the tile a
CTA of the compaction kernel owns) and 16384 candidates; a candidate is "removable" when an earlier
candidate of the same group names the same vertex.  Every candidate still needs its atomicMax on
the winner's entry unless it is removable, and the group's survivors still go to the global table.
    python tools/dedup_potential.py [--scale 1.0]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import oracle as O
from salient_plusplus_b200 import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--shape", default="products")
a = ap.parse_args()
n, e, f, dt = S.SHAPES[a.shape]
n, e = int(n * a.scale), int(e * a.scale)
rowptr, col = S.powerlaw_graph(n, e, seed=1)
seeds = S.seeds(n, 1024, seed=7).numpy()
sizes = [15, 10, 5]
n_id, adjs = O.multilayer_sample(seeds, sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER, rng_seed=12345)
out = {"shape": a.shape, "scale": a.scale, "nodes_in_batch": int(n_id.size), "hops": []}
for h, (rp, cl, _e, (T, Sz)) in enumerate(reversed(adjs)):      # hop order
    cand = n_id[cl]                                              # global id of every kept candidate, target-major
    rec = {"hop": h, "targets": int(T), "candidates": int(cand.size), "new_nodes": int(Sz - T)}
    for g in (32, 256, 2048, 16384):
        removable = 0
        for s0 in range(0, cand.size, g):
            blk = cand[s0:s0 + g]
            removable += blk.size - np.unique(blk).size
        rec[f"removable_in_groups_of_{g}"] = round(removable / max(cand.size, 1), 4)
    out["hops"].append(rec)
print(json.dumps(out, indent=1))
