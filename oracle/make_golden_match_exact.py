"""TEST INFRASTRUCTURE ONLY — tests/golden/match_exact.npz from the REAL reference Matcher.

Run in the build container:  python -m oracle.make_golden_match_exact

The index and the 24 queries of tests/golden/match.npz, matched by the reference's `Matcher.match_hashes`
(afp/audfprint/audfprint_match.py:318-349) with
  * exact_count = True, find_time_range = True, hashesfor = 0   (`_exact_match_counts` :183-233,
    `_unique_match_hashes` :130-152, `_calculate_time_ranges` :155-181), and
  * exact_count = False, find_time_range = True                  (the time-range columns of `_approx_match_counts`).
threshcount is lowered to 3 for half of the queries so that several rows / several modes per track appear.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ns = ref_loader.load()
    g = np.load(os.path.join(GOLD, "match.npz"))
    ht = ns.hash_table.HashTable()
    ht.depth = 100
    idx = g["counts_nonzero_idx"]
    ht.table[idx] = g["table_rows"]
    ht.counts[idx] = g["counts_nonzero"]
    ht.hashesperid = g["hashesperid"].copy()
    ht.names = [f"track{t:04d}" for t in range(int(g["n_tracks"]))]
    out = {"meta": str(g["meta"]) + " | match_exact: reference Matcher, exact_count / find_time_range / hashesfor"}
    for q in range(int(g["n_queries"])):
        qh = g[f"q{q}"]
        m = ns.match.Matcher()
        m.exact_count, m.find_time_range = True, True
        if len(ht.get_hits(qh)) == 0:
            continue
        res, hf = m.match_hashes(ht, qh, hashesfor=0) if len(g[f"res{q}"]) else (m.match_hashes(ht, qh)[0], np.zeros((0, 2)))
        out[f"exact{q}"] = np.asarray(res, np.int32).reshape(-1, 7)
        out[f"hashesfor{q}"] = np.asarray(hf, np.int64).reshape(-1, 2)
        m2 = ns.match.Matcher()
        m2.find_time_range = True
        out[f"approx_tr{q}"] = np.asarray(m2.match_hashes(ht, qh)[0], np.int32).reshape(-1, 7)
    np.savez_compressed(os.path.join(GOLD, "match_exact.npz"), **out)
    print("match_exact.npz:", {k: v.shape for k, v in out.items() if k != "meta"})


if __name__ == "__main__":
    main()
