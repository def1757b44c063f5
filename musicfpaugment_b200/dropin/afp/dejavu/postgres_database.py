"""Drop-in for afp/dejavu/postgres_database.py: `PostgreSQLDatabase` with the reference's method surface,
backed by an in-memory store and a device-resident, hash-sorted fingerprints table instead of a Postgres
server.

The reference issues one SELECT per query hash (`return_matches`, postgres_database.py:180-229, batch_size 1);
here the distinct query hashes are looked up by binary search on the GPU (`mfpa_dejavu_return_matches`).  Stores
are shared per database name within the process - two `Dejavu` objects built from the same config see the same
songs, like two connections to one server (testing/dejavu_exps.py:21-26, :82-96).  `save` / `load` persist a
store as an .npz file (what the Postgres volume did).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import numpy as np

FIELD_SONG_ID, FIELD_SONGNAME, FIELD_FINGERPRINTED = "song_id", "song_name", "fingerprinted"
FIELD_FILE_SHA1, FIELD_TOTAL_HASHES = "file_sha1", "total_hashes"


def split_hash(h: str) -> Tuple[int, int]:
    """A Dejavu hash (hex digits of a SHA-1, FINGERPRINT_REDUCTION = 20 of them, variables.py:22) ->
    (leading 16 digits as an integer, remaining digits as an integer): the index's (key, tail)."""
    h = h.upper()
    if len(h) > 20:
        raise ValueError(f"hash {h!r} has more than 20 hex digits")
    h = h.ljust(20, "0")
    return int(h[:16], 16), int(h[16:], 16)


class _Store:
    def __init__(self):
        self.songs: List[dict] = []              # row i = song id i + 1 (SERIAL starts at 1)
        self.keys = np.zeros(0, np.uint64)
        self.tails = np.zeros(0, np.uint16)
        self.song = np.zeros(0, np.int32)
        self.offset = np.zeros(0, np.int32)
        self.pending: List[tuple] = []           # (keys, tails, songs, offsets) chunks not merged yet
        self.version = 0
        self.device = None                       # (version, lib.DejavuIndex)

    def merge(self):
        if not self.pending:
            return
        k = np.concatenate([self.keys] + [p[0] for p in self.pending])
        t = np.concatenate([self.tails] + [p[1] for p in self.pending])
        s = np.concatenate([self.song] + [p[2] for p in self.pending])
        o = np.concatenate([self.offset] + [p[3] for p in self.pending])
        self.pending = []
        # UNIQUE (song_id, offset, hash) ... ON CONFLICT DO NOTHING (postgres_database.py:288-296)
        order = np.lexsort((o, s, t, k))
        k, t, s, o = k[order], t[order], s[order], o[order]
        keep = np.ones(len(k), bool)
        keep[1:] = (k[1:] != k[:-1]) | (t[1:] != t[:-1]) | (s[1:] != s[:-1]) | (o[1:] != o[:-1])
        self.keys, self.tails, self.song, self.offset = k[keep], t[keep], s[keep], o[keep]
        self.version += 1

    def index(self):
        from musicfpaugment_b200 import lib, runtime

        self.merge()
        if self.device is None or self.device[0] != self.version:
            if self.device is not None:
                self.device[1].close()
            self.device = (self.version, lib.DejavuIndex(runtime.get_context(), self.keys, self.tails, self.song, self.offset,
                                                         len(self.songs) + 1))
        return self.device[1]


_STORES: Dict[str, _Store] = {}


class PostgreSQLDatabase:
    type = "memory"

    def __init__(self, **options):
        self._options = dict(options)
        self._name = str(options.get("database", "dejavu"))
        self._store = _STORES.setdefault(self._name, _Store())

    # -- schema ------------------------------------------------------------------------------
    def before_fork(self) -> None:
        pass

    def after_fork(self) -> None:
        pass

    def setup(self) -> None:
        """CREATE TABLE IF NOT EXISTS ...; DELETE unfingerprinted songs (postgres_database.py:32-39)."""
        self.delete_unfingerprinted_songs()

    def empty(self) -> None:
        old = _STORES.get(self._name)
        if old is not None and old.device is not None:
            old.device[1].close()
        self._store = _STORES[self._name] = _Store()

    def delete_unfingerprinted_songs(self) -> None:
        dead = [s[FIELD_SONG_ID] for s in self._store.songs if s is not None and not s[FIELD_FINGERPRINTED]]
        if dead:
            self.delete_songs_by_id(dead)

    # -- songs -------------------------------------------------------------------------------
    def get_num_songs(self) -> int:
        return sum(1 for s in self._store.songs if s is not None and s[FIELD_FINGERPRINTED])

    def get_num_fingerprints(self) -> int:
        self._store.merge()
        return int(len(self._store.keys))

    def set_song_fingerprinted(self, song_id) -> None:
        self._store.songs[int(song_id) - 1][FIELD_FINGERPRINTED] = 1

    def get_songs(self) -> List[Dict[str, str]]:
        return [{k: s[k] for k in (FIELD_SONG_ID, FIELD_SONGNAME, FIELD_FILE_SHA1, FIELD_TOTAL_HASHES)}
                for s in self._store.songs if s is not None and s[FIELD_FINGERPRINTED]]

    def get_song_by_id(self, song_id: int) -> Dict[str, str]:
        i = int(song_id) - 1
        if i < 0 or i >= len(self._store.songs) or self._store.songs[i] is None:
            return None
        s = self._store.songs[i]
        return {k: s[k] for k in (FIELD_SONGNAME, FIELD_FILE_SHA1, FIELD_TOTAL_HASHES)}

    def insert_song(self, song_name: str, file_hash: str, total_hashes: int) -> int:
        sid = len(self._store.songs) + 1
        self._store.songs.append({FIELD_SONG_ID: sid, FIELD_SONGNAME: song_name, FIELD_FINGERPRINTED: 0,
                                  FIELD_FILE_SHA1: str(file_hash).upper(), FIELD_TOTAL_HASHES: int(total_hashes)})
        self._store.version += 1
        return sid

    def delete_songs_by_id(self, song_ids: List[int], batch_size: int = 1000) -> None:
        st = self._store
        st.merge()
        ids = np.asarray(list(song_ids), np.int32)
        keep = ~np.isin(st.song, ids)           # ON DELETE CASCADE
        st.keys, st.tails, st.song, st.offset = st.keys[keep], st.tails[keep], st.song[keep], st.offset[keep]
        for i in ids:
            if 1 <= i <= len(st.songs):
                st.songs[int(i) - 1] = None
        st.version += 1

    # -- fingerprints ------------------------------------------------------------------------
    def insert(self, fingerprint: str, song_id: int, offset: int) -> None:
        self.insert_hashes(song_id, [(fingerprint, offset)])

    def insert_hashes(self, song_id: int, hashes: Iterable[Tuple[str, int]], batch_size: int = 100) -> None:
        hashes = list(hashes)
        if not hashes:
            return
        kt = [split_hash(h) for h, _ in hashes]
        self._store.pending.append((np.array([k for k, _ in kt], np.uint64), np.array([t for _, t in kt], np.uint16),
                                    np.full(len(hashes), int(song_id), np.int32),
                                    np.array([int(o) for _, o in hashes], np.int32)))

    def query(self, fingerprint: str = None) -> List[Tuple]:
        st = self._store
        st.merge()
        if fingerprint is None:
            return list(zip(st.song.tolist(), st.offset.tolist()))
        k, t = split_hash(fingerprint)
        lo, hi = np.searchsorted(st.keys, np.uint64(k), "left"), np.searchsorted(st.keys, np.uint64(k), "right")
        sel = st.tails[lo:hi] == t
        return list(zip(st.song[lo:hi][sel].tolist(), st.offset[lo:hi][sel].tolist()))

    def get_iterable_kv_pairs(self) -> List[Tuple]:
        return self.query(None)

    # -- the query hot path ------------------------------------------------------------------
    def match_pairs(self, hashes: Iterable[Tuple[str, int]]):
        """`return_matches` with the results left on the device: (pairs int32 [n,2] cuda tensor of
        (song id, database offset - query offset), dedup int32 [n_songs + 1] cuda tensor)."""
        mapper: Dict[str, List[int]] = {}
        for h, off in hashes:
            mapper.setdefault(h.upper(), []).append(int(off))       # postgres_database.py:198-204
        keys, tails, start, offs = [], [], [0], []
        for h, lst in mapper.items():
            k, t = split_hash(h)
            keys.append(k); tails.append(t); offs.extend(lst); start.append(len(offs))
        return self._store.index().return_matches(np.array(keys, np.uint64), np.array(tails, np.uint16),
                                                  np.array(start, np.int32), np.array(offs, np.int32))

    def return_matches(self, hashes: Iterable[Tuple[str, int]], batch_size: int = 1):
        """postgres_database.py:180-229: ([(song id, offset difference), ...], {song id: matched hashes})."""
        pairs, dedup = self.match_pairs(hashes)
        rows = pairs.cpu().numpy()
        d = dedup.cpu().numpy()
        return [(int(s), int(o)) for s, o in rows], {int(i): int(d[i]) for i in np.nonzero(d)[0]}

    # -- persistence (the Postgres volume) ---------------------------------------------------
    def save(self, path: str) -> None:
        st = self._store
        st.merge()
        songs = [s for s in st.songs]
        np.savez_compressed(path, keys=st.keys, tails=st.tails, song=st.song, offset=st.offset,
                            songs=np.array([repr(s) for s in songs], dtype=object))

    def load(self, path: str) -> None:
        import ast

        g = np.load(path, allow_pickle=True)
        self.empty()
        st = self._store
        st.keys, st.tails, st.song, st.offset = g["keys"], g["tails"], g["song"], g["offset"]
        st.songs = [ast.literal_eval(s) for s in g["songs"].tolist()]
        st.version += 1

    def __getstate__(self):
        return (self._options,)

    def __setstate__(self, state):
        (self._options,) = state
        self.__init__(**self._options)
