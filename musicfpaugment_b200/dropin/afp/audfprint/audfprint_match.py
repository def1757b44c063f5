"""Drop-in for afp/audfprint/audfprint_match.py: Matcher with the reference defaults
(audfprint_match.py:76-100); match_hashes runs the libmfpa matching kernels."""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np

from musicfpaugment_b200 import lib


class Matcher(object):
    def __init__(self) -> None:
        self.window = 2
        self.threshcount = 5
        self.max_returns = 1
        self.search_depth = 100
        self.sort_by_time = False
        self.verbose = 1
        self.exact_count = False
        self.find_time_range = False
        self.time_quantile = 0.05
        self.max_alignments_per_id = 100

    def _params(self) -> lib.MatchParams:
        if self.exact_count or self.find_time_range:
            raise NotImplementedError("exact_count / find_time_range are not on the B200 path "
                                      "(the reference's drivers leave both False, audfprint_match.py:93-95)")
        p = lib.MatchParams()
        p.window, p.threshcount, p.search_depth = self.window, self.threshcount, self.search_depth
        p.max_alignments_per_id = self.max_alignments_per_id
        return p

    def match_hashes(self, ht, hashes, hashesfor: Optional[int] = None) -> Tuple[Any, Any]:
        """-> (int32 [k,7] rows (id, filtered count, time skew, raw count, rank, 0, 0) sorted by
        filtered count descending, None).  audfprint_match.py:318-349."""
        res = self.match_hashes_batch(ht, [hashes])[0]
        if hashesfor is not None:
            raise NotImplementedError("hashesfor (matching-hash dump) is a reporting helper, not on the B200 path")
        return res, None

    def match_hashes_batch(self, ht, hashes_list, max_rows: int = 128):
        """The batched form: list of int32 [n_i,2] query hash arrays -> list of result arrays."""
        import torch

        ctx = ht._device()
        cap = max(1, max(len(h) for h in hashes_list))
        hb = np.zeros((len(hashes_list), cap, 2), np.int32)
        nh = np.zeros(len(hashes_list), np.int32)
        for i, h in enumerate(hashes_list):
            h = np.asarray(h, dtype=np.int32).reshape(-1, 2)
            hb[i, : len(h)], nh[i] = h, len(h)
        hb_d, nh_d = torch.from_numpy(hb).cuda(), torch.from_numpy(nh).cuda()
        res, nrows = ctx.match(hb_d, nh_d, self._params(), max_rows)
        res, nrows = res.cpu().numpy(), nrows.cpu().numpy()
        if (nrows == -6).any():
            # a track collected >= 65536 hits from one query: the packed 16-bit counters of the one-kernel matcher
            # overflowed. That is detected (never silent); such a query is also far past the 8192 candidate hits the
            # alignment step sorts, so it cannot be answered - MFPA_OPT_MATCH_UNFUSED = 2 still gives its exact raw
            # counts through mfpa_match_counts.
            raise lib.MfpaError("match: a track collected >= 65536 hits from one query (INTEGRATION.md section 5)")
        if (nrows == -5).any():
            raise lib.MfpaError("match: a query hash time lies outside [0, 16384) frames - the matching kernels pack "
                                "t_ref - t_q next to the table's 14-bit reference times (INTEGRATION.md section 5)")
        if (nrows < 0).any():
            raise lib.MfpaError("match: a per-query capacity was exceeded (too many candidate hits / result rows)")
        return [res[i, : nrows[i]].copy() for i in range(len(hashes_list))]

    def match_file(self, analyzer, ht, filename: str):
        """audfprint_match.py:351-371."""
        q_hashes = analyzer.wavfile2hashes(filename)
        durd = 0.0 if len(q_hashes) == 0 else analyzer.n_hop * q_hashes[-1][0] / analyzer.target_sr
        rslts, _ = self.match_hashes(ht, q_hashes)
        if self.sort_by_time:
            rslts = rslts[(-rslts[:, 2]).argsort(), :]
        return rslts[: self.max_returns, :], durd, len(q_hashes)

    def file_match_to_msgs(self, analyzer, ht, qry: str):
        """-> ("MATCH", name, aligned hashes) or ("NOMATCH", "", 0).  audfprint_match.py:373-435."""
        rslts, dur, nhash = self.match_file(analyzer, ht, qry)
        if len(rslts) == 0:
            return "NOMATCH", "", 0
        tophitid, nhashaligned = int(rslts[-1][0]), int(rslts[-1][1])  # the loop's last row, as in the reference
        return "MATCH", ht.names[tophitid], nhashaligned
