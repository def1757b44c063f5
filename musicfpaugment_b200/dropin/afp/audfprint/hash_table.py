"""Drop-in for afp/audfprint/hash_table.py.  Same fields and gzip-pickle format
(hash_table.py:46-68,118-198), so an index the reference built loads here and vice versa;
`get_hits` and the Matcher run on the GPU against a device copy of the table that is
refreshed whenever the host table is marked dirty."""
from __future__ import annotations

import gzip
import math
import pickle
import random
from typing import Any, List, Union

import numpy as np

HT_VERSION = 20170724
HT_COMPAT_VERSION = 20170724
HT_OLD_COMPAT_VERSION = 20140920


def _bitsfor(maxval: int) -> int:
    maxvalbits = int(round(math.log(maxval) / math.log(2)))
    if maxval != (1 << maxvalbits):
        raise ValueError("maxval must be a power of 2, not %d" % maxval)
    return maxvalbits


class HashTable(object):
    def __init__(self, filename=None):
        if filename is not None:
            self.load(filename)
        else:
            self.hashbits = 20
            self.depth = 100
            self.maxtimebits = _bitsfor(16384)
            size = 2 ** self.hashbits
            self.table = np.zeros((size, self.depth), dtype=np.uint32)
            self.counts = np.zeros(size, dtype=np.int32)
            self.names: List[Any] = []
            self.hashesperid = np.zeros(0, np.uint32)
            self.ht_version = HT_VERSION
            self.dirty = True

    # ---- device residency (not pickled)
    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_dev")}

    def _device(self):
        """Upload the table once per modification; returns the libmfpa context holding it."""
        from musicfpaugment_b200 import runtime

        ctx = runtime.get_context()
        stamp = (id(self.table), int(self.counts.sum()), len(self.hashesperid))
        if getattr(self, "_dev_stamp", None) != stamp or getattr(ctx, "_index_owner", None) is not self:
            if len(self.hashesperid) == 0:
                raise ValueError("hash table is empty")
            ctx.index_load(self.table, self.counts, self.hashesperid, 0, self.hashbits, self.maxtimebits)
            ctx._index_owner = self
            self._dev_stamp = stamp
        return ctx

    # ---- store (hash_table.py:70-116): host side, index building is not the query hot path
    def store(self, name, timehashpairs) -> None:
        id_ = self.name_to_id(name, add_if_missing=True)
        hashmask = (1 << self.hashbits) - 1
        timemask = (1 << self.maxtimebits) - 1
        idval = (id_ + 1) << self.maxtimebits
        for time_, hash_ in np.asarray(timehashpairs).reshape(-1, 2):
            hash_ = int(hash_) & hashmask
            count = int(self.counts[hash_])
            val = idval + (int(time_) & timemask)
            if count < self.depth:
                self.table[hash_, count] = val
            else:
                slot = random.randint(0, count)  # reservoir-style overwrite (:105-110)
                if slot < self.depth:
                    self.table[hash_, slot] = val
            self.counts[hash_] = count + 1
        self.hashesperid[id_] += len(timehashpairs)
        self.dirty = True

    # ---- persistence (hash_table.py:118-198)
    def save(self, name: str) -> None:
        with gzip.open(name, "wb") as f:
            pickle.dump(self, f, pickle.HIGHEST_PROTOCOL)
        self.dirty = False

    def load(self, name: str) -> None:
        with gzip.open(name, "rb") as f:
            temp = pickle.load(f, encoding="latin1")
        if temp.ht_version < HT_OLD_COMPAT_VERSION:
            raise ValueError("Version of %s is %s which is not at least %s" % (name, temp.ht_version, HT_OLD_COMPAT_VERSION))
        self.hashbits = temp.hashbits
        self.depth = temp.depth
        self.maxtimebits = temp.maxtimebits if hasattr(temp, "maxtimebits") else _bitsfor(temp.maxtime)
        if temp.ht_version < HT_COMPAT_VERSION:
            temp.table += np.array(1 << self.maxtimebits).astype(np.uint32) * (temp.table != 0)
            temp.ht_version = HT_VERSION
        self.table = temp.table
        self.ht_version = temp.ht_version
        self.counts = temp.counts
        self.names = temp.names
        self.hashesperid = np.array(temp.hashesperid).astype(np.uint32)
        self.dirty = False

    def reset(self) -> None:
        self.table[:, :] = 0
        self.counts[:] = 0
        self.names = []
        self.hashesperid = np.zeros(0, np.uint32)
        self.dirty = True

    # ---- queries
    def get_entry(self, hash_: int):
        vals = self.table[hash_, : min(self.depth, self.counts[hash_])]
        ids = (vals >> self.maxtimebits) - 1
        return np.c_[ids, vals & ((1 << self.maxtimebits) - 1)].astype(np.int32)

    def get_hits(self, hashes):
        """[id, delta_time, hash, time] rows for each (time, hash) query row — GPU gather
        (hash_table.py:220-246)."""
        import torch

        hashes = np.ascontiguousarray(np.asarray(hashes, dtype=np.int32).reshape(-1, 2))
        if len(hashes) == 0:
            return np.zeros((0, 4), np.int32)
        return self._device().get_hits(torch.from_numpy(hashes).cuda()).cpu().numpy()

    def totalhashes(self):
        return np.sum(self.counts)

    def name_to_id(self, name: Union[int, str], add_if_missing: bool = False) -> int:
        if isinstance(name, str):
            if name not in self.names:
                if not add_if_missing:
                    raise ValueError("name " + name + " not found")
                try:
                    id_ = self.names.index(None)
                    self.names[id_] = name
                    self.hashesperid[id_] = 0
                except ValueError:
                    self.names.append(name)
                    self.hashesperid = np.append(self.hashesperid, [0]).astype(np.uint32)
            return self.names.index(name)
        return name
