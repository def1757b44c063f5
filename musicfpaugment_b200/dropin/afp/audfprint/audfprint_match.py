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
        p = lib.MatchParams()
        p.window, p.threshcount, p.search_depth = self.window, self.threshcount, self.search_depth
        p.max_alignments_per_id = self.max_alignments_per_id
        return p

    def match_hashes(self, ht, hashes, hashesfor: Optional[int] = None) -> Tuple[Any, Any]:
        """-> (int32 [k,7] rows (id, filtered count, time skew, raw count, rank, 0, 0) sorted by
        filtered count descending, None).  audfprint_match.py:318-349."""
        if not (self.exact_count or self.find_time_range or hashesfor is not None):
            return self.match_hashes_batch(ht, [hashes])[0], None
        # The reporting options (off in the reference's drivers, audfprint_match.py:93-95).  The table look-ups, the
        # per-track counts and the candidate ranking stay on the GPU (mfpa_get_hits, mfpa_match_counts/_select, or
        # the whole of mfpa_match when only the time range is wanted); what is left - distinct (time, hash) pairs and
        # quantiles of a few hundred rows per candidate - is host arithmetic on those hits.
        hits = ht.get_hits(hashes)
        if self.exact_count:
            ids, raw = self._candidates(ht, hashes)
            res = self._exact_rows(hits, ids, raw)
        else:
            res = self.match_hashes_batch(ht, [hashes])[0]
            if len(res):
                by_time = hits[np.argsort(hits[:, 3], kind="stable")]
                for row in res:
                    row[5], row[6] = self._time_range(by_time, int(row[0]), int(row[2]))
        res = res[np.argsort(-res[:, 1], kind="stable")]
        if hashesfor is None:
            return res, None
        return res, self._unique_match_hashes(int(res[hashesfor, 0]), hits, int(res[hashesfor, 2]))

    def _candidates(self, ht, hashes):
        """(track ids, raw counts) ranked as _best_count_ids does (audfprint_match.py:102-129), from the GPU."""
        import torch

        h = np.ascontiguousarray(np.asarray(hashes, dtype=np.int32).reshape(1, -1, 2))
        if h.shape[1] == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        ctx = ht._device()
        h_d, n_d = torch.from_numpy(h).cuda(), torch.tensor([h.shape[1]], dtype=torch.int32, device="cuda")
        cand, ncand = ctx.match_select(ctx.match_counts(h_d, n_d), self._params())
        k = int(ncand[0])
        c = cand[0, :k].cpu().numpy().astype(np.int64)
        return c[:, 0], c[:, 1]

    def _unique_match_hashes(self, id: int, hits, mode) -> Any:
        """Distinct [time, hash] rows of track `id` whose skew is within the window of `mode`
        (audfprint_match.py:130-152), ordered by (hash, time)."""
        near = hits[(hits[:, 0] == id) & (np.abs(hits[:, 1] - mode) <= self.window)]
        pairs = np.unique(near[:, [2, 3]].astype(np.int64), axis=0)       # rows (hash, time), sorted by hash then time
        return pairs[:, ::-1].copy()

    def _time_range(self, by_time, id: int, mode) -> Tuple[int, int]:
        """The time_quantile / 1 - time_quantile points of the matching query times (audfprint_match.py:155-181);
        `by_time` = hits ordered by query time."""
        t = by_time[(by_time[:, 0] == id) & (np.abs(by_time[:, 1] - mode) <= self.window), 3]
        return int(t[int(len(t) * self.time_quantile)]), int(t[int(len(t) * (1.0 - self.time_quantile)) - 1])

    def _exact_rows(self, hits, ids, rawcounts):
        """audfprint_match.py:183-233: one row per candidate and modal skew whose distinct-hash count reaches the
        threshold.  A skew is modal when its count is a local maximum of the track's skew histogram (ties resolved
        like the reference's locmax: a plateau peaks at its last bin) and at least threshcount."""
        rows = []
        if len(hits):
            by_time = hits[np.argsort(hits[:, 3], kind="stable")]
            for rank, (id, raw) in enumerate(zip(ids, rawcounts)):
                dts = by_time[by_time[:, 0] == id, 1].astype(np.int64)
                lo = int(dts.min())
                hist = np.bincount(dts - lo)
                rising = np.r_[True, hist[1:] >= hist[:-1]]           # >= the bin before (the first bin always is)
                falling = np.r_[hist[1:] < hist[:-1], True]           # > the bin after (the last bin always is)
                for mode in np.nonzero(rising & falling & (hist >= self.threshcount))[0] + lo:
                    n = len(self._unique_match_hashes(int(id), by_time, int(mode)))
                    if n >= self.threshcount:
                        t0, t1 = self._time_range(by_time, int(id), int(mode)) if self.find_time_range else (0, 0)
                        rows.append([int(id), n, int(mode), int(raw), rank, t0, t1])
        return np.asarray(rows, dtype=np.int32).reshape(-1, 7)

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
