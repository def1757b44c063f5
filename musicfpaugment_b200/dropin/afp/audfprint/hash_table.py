"""Drop-in for afp/audfprint/hash_table.py.

Same public fields (`hashbits`, `depth`, `maxtimebits`, `table`, `counts`, `names`, `hashesperid`,
`ht_version`, `dirty`) and the same gzip-pickle file format (hash_table.py:46-68, 118-198), so an index
the reference built loads here and a file saved here loads in the reference.  `get_hits` and the Matcher
run on the GPU against a device copy of the table; every mutation bumps a version counter and the copy
is refreshed when it is behind.

The host-side bookkeeping is vectorised: `store` places all rows that find a free slot at once and only
the rows that land in a FULL bucket go through Python's `random.randint` - one draw per such row, in row
order, which is exactly the sequence of draws the reference's per-row loop makes (hash_table.py:91-112), so
a table built here under the same `random.seed` is identical to the reference's.
"""
from __future__ import annotations

import gzip
import math
import pickle
import random
from typing import Any, Callable, List, Optional, Union

import numpy as np

HT_VERSION = 20170724            # current format: ids stored as id + 1 (hash_table.py:24-30)
HT_COMPAT_VERSION = 20170724
HT_OLD_COMPAT_VERSION = 20140920


def _bitsfor(maxval: int) -> int:
    """log2 of a power of two (hash_table.py:33-43)."""
    bits = int(round(math.log(maxval) / math.log(2)))
    if maxval != (1 << bits):
        raise ValueError("maxval must be a power of 2, not %d" % maxval)
    return bits


class HashTable(object):
    """Fixed-array landmark index: table[2^hashbits, depth] uint32, entry = ((id + 1) << maxtimebits) + time."""

    def __init__(self, filename: Optional[str] = None, hashbits: int = 20, depth: int = 100, maxtime: int = 16384):
        self._version = 0
        if filename is not None:
            self.load(filename)
            return
        self.hashbits, self.depth, self.maxtimebits = hashbits, depth, _bitsfor(maxtime)
        self.table = np.zeros((1 << hashbits, depth), dtype=np.uint32)
        self.counts = np.zeros(1 << hashbits, dtype=np.int32)
        self.names: List[Any] = []
        self.hashesperid = np.zeros(0, np.uint32)
        self.ht_version = HT_VERSION
        self.dirty = True

    # ---- device residency (never pickled: the reference's class must be able to load our files) -------------
    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._version = 0

    def _touch(self) -> None:
        self._version = getattr(self, "_version", 0) + 1
        self.dirty = True

    def _device(self):
        """The libmfpa context holding a current copy of this table (uploaded when a mutation made it stale, or
        when another table took the context's index slot)."""
        from musicfpaugment_b200 import runtime

        ctx = runtime.get_context()
        if getattr(self, "_dev_version", None) != self._version or getattr(ctx, "_index_owner", None) is not self:
            if len(self.hashesperid) == 0:
                raise ValueError("hash table is empty")
            ctx.index_load(self.table, self.counts, self.hashesperid, 0, self.hashbits, self.maxtimebits)
            ctx._index_owner = self
            self._dev_version = self._version
        return ctx

    # ---- building (hash_table.py:70-116); index building is not the query hot path --------------------------
    def store(self, name: Union[int, str], timehashpairs) -> None:
        id_ = self.name_to_id(name, add_if_missing=True)
        pairs = np.asarray(timehashpairs, dtype=np.int64).reshape(-1, 2)
        n = len(pairs)
        if n:
            h = pairs[:, 1] & ((1 << self.hashbits) - 1)
            val = (((id_ + 1) << self.maxtimebits) + (pairs[:, 0] & ((1 << self.maxtimebits) - 1))).astype(np.uint32)
            # bucket fill level each row sees = level before the call + earlier rows of this call in the same bucket
            order = np.argsort(h, kind="stable")
            hs = h[order]
            first = np.r_[0, np.nonzero(hs[1:] != hs[:-1])[0] + 1]
            run_len = np.diff(np.r_[first, n])
            earlier = np.empty(n, np.int64)
            earlier[order] = np.arange(n) - np.repeat(first, run_len)
            level = self.counts[h].astype(np.int64) + earlier
            free = level < self.depth
            self.table[h[free], level[free]] = val[free]
            for i in np.nonzero(~free)[0]:                      # full bucket: overwrite a random slot, maybe (:105-110)
                slot = random.randint(0, int(level[i]))
                if slot < self.depth:
                    self.table[h[i], slot] = val[i]
            self.counts[hs[first]] += run_len.astype(np.int32)
        self.hashesperid[id_] += n
        self._touch()

    def name_to_id(self, name: Union[int, str], add_if_missing: bool = False) -> int:
        """Index of `name` in `names`; a new name takes the first freed slot, else a new one (hash_table.py:254-275)."""
        if not isinstance(name, str):
            return name
        if name in self.names:
            return self.names.index(name)
        if not add_if_missing:
            raise ValueError("name " + name + " not found")
        if None in self.names:
            id_ = self.names.index(None)
            self.names[id_] = name
            self.hashesperid[id_] = 0
        else:
            id_ = len(self.names)
            self.names.append(name)
            self.hashesperid = np.append(self.hashesperid, [0]).astype(np.uint32)
        return id_

    def remove(self, name: Union[str, int]) -> None:
        """Drop every entry of one item (hash_table.py:277-297): its buckets are compacted, `counts` forgets the
        entries those buckets had already dropped, the name slot is freed."""
        id_ = self.name_to_id(name)
        mine = (self.table >> self.maxtimebits) == id_ + 1
        removed = 0
        for h in np.nonzero(mine.any(axis=1))[0]:
            stored = min(self.depth, int(self.counts[h]))
            row = self.table[h, :stored]
            keep = row[~mine[h, :stored]]
            self.table[h] = 0
            self.table[h, : len(keep)] = keep
            self.counts[h] = len(keep)
            removed += int(mine[h].sum())
        self.names[id_] = None
        self.hashesperid[id_] = 0
        self._touch()
        print("Removed", name, "(", removed, "hashes).")

    def retrieve(self, name: Union[str, int]) -> np.ndarray:
        """(time, hash) rows of one item, buckets ascending, slot order inside a bucket (hash_table.py:299-317)."""
        id_ = self.name_to_id(name)
        stored = np.arange(self.depth)[None, :] < np.minimum(self.counts, self.depth)[:, None]
        h, slot = np.nonzero(((self.table >> self.maxtimebits) == id_ + 1) & stored)
        times = self.table[h, slot] & ((1 << self.maxtimebits) - 1)
        return np.stack([times, h], axis=1).astype(np.int32)

    def list(self, print_fn: Optional[Callable[[str], None]] = None) -> None:
        print_fn = print_fn or print
        for name, count in zip(self.names, self.hashesperid):
            if name:
                print_fn(name + " (" + str(count) + " hashes)")

    def reset(self) -> None:
        self.table[:, :] = 0
        self.counts[:] = 0
        self.names = []
        self.hashesperid = np.zeros(0, np.uint32)
        self._touch()

    # ---- persistence (hash_table.py:118-198) -----------------------------------------------------------------
    def save(self, name: str, params: Any = None, file_object: Any = None) -> None:
        self.params = params
        with (file_object or gzip.open(name, "wb")) as f:
            pickle.dump(self, f, pickle.HIGHEST_PROTOCOL)
        self.dirty = False
        total = int(self.totalhashes())
        dropped = total - int(np.minimum(self.depth, self.counts).sum())
        print("Saved fprints for", sum(n is not None for n in self.names), "files (", total, "hashes) to", name,
              "(%.2f%% dropped)" % (100.0 * dropped / max(1, total)))

    def load(self, name: str) -> None:
        self.load_pkl(name)

    def load_pkl(self, name: str, file_object: Any = None) -> None:
        """Read a table the reference (or this class) pickled; tables older than the id + 1 convention are upgraded
        in memory (hash_table.py:141-198)."""
        with (file_object or gzip.open(name, "rb")) as f:
            src = pickle.load(f, encoding="latin1")
        if src.ht_version < HT_OLD_COMPAT_VERSION:
            raise ValueError("Version of " + name + " is " + str(src.ht_version) + " which is not at least " + str(HT_OLD_COMPAT_VERSION))
        self.hashbits, self.depth = src.hashbits, src.depth
        self.maxtimebits = src.maxtimebits if hasattr(src, "maxtimebits") else _bitsfor(src.maxtime)
        table = np.asarray(src.table, dtype=np.uint32)
        if src.ht_version < HT_COMPAT_VERSION:      # ids were stored zero-based: shift every non-empty entry by one id
            print("Loading database version", src.ht_version, "in compatibility mode.")
            table = table + np.uint32(1 << self.maxtimebits) * (table != 0)
        self.table, self.ht_version = table, HT_VERSION
        self.counts = np.asarray(src.counts, dtype=np.int32)
        self.names = list(src.names)
        self.hashesperid = np.array(src.hashesperid).astype(np.uint32)
        self.params = getattr(src, "params", None)
        self._touch()
        self.dirty = False
        total = int(self.totalhashes())
        dropped = total - int(np.minimum(self.depth, self.counts).sum())
        print("Read fprints for", sum(n is not None for n in self.names), "files (", total, "hashes) from", name,
              "(%.2f%% dropped)" % (100.0 * dropped / max(1, total)))

    # ---- queries ---------------------------------------------------------------------------------------------
    def get_entry(self, hash_: int) -> np.ndarray:
        """[id, time] rows of one bucket (hash_table.py:210-218)."""
        vals = self.table[hash_, : min(self.depth, self.counts[hash_])]
        return np.c_[(vals >> self.maxtimebits) - 1, vals & ((1 << self.maxtimebits) - 1)].astype(np.int32)

    def get_hits(self, hashes) -> np.ndarray:
        """[id, delta_time, hash, time] rows for each (time, hash) query row - GPU gather (hash_table.py:220-246)."""
        import torch

        hashes = np.ascontiguousarray(np.asarray(hashes, dtype=np.int32).reshape(-1, 2))
        if len(hashes) == 0:
            return np.zeros((0, 4), np.int32)
        return self._device().get_hits(torch.from_numpy(hashes).cuda()).cpu().numpy()

    def totalhashes(self):
        return np.sum(self.counts)
