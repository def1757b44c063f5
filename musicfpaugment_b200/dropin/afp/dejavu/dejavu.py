"""Drop-in for afp/dejavu/dejavu.py: `Dejavu`, `read`, `unique_hash` with the reference's signatures.

What changes: the fingerprints live in an in-memory, device-resident index instead of Postgres
(`postgres_database.PostgreSQLDatabase` of this package), `find_matches` is one batched GPU lookup instead of
one SELECT per hash, and `align_matches` takes its vote on the GPU (`mfpa_dejavu_align`).  The module does not
load Demucs at import (dejavu.py:33-42 does, from a hard-coded checkpoint path): `denoising_model="demucs"` is
outside the B200 path and raises NotImplementedError; `"unet"` runs inside `fingerprint()`.
"""
from __future__ import annotations

import os
import pickle
from hashlib import sha1
from time import time
from typing import Dict, List, Tuple

import numpy as np
import torch

from dejavu.fingerprint import fingerprint
from dejavu.postgres_database import PostgreSQLDatabase
from dejavu.variables import (FINGERPRINTED_CONFIDENCE, FINGERPRINTED_HASHES, HASHES_MATCHED, INPUT_CONFIDENCE,
                              INPUT_CONFIDENCE_2, INPUT_HASHES, OFFSET, OFFSET_SECS, SONG_ID, SONG_NAME, TOPN)

try:  # the reference reads both from testing.parameters (dejavu.py:29)
    from testing.parameters import WAVEFORM_SAMPLING_RATE, afp_settings
except Exception:  # reference tree not on sys.path: its values (testing/parameters.py:15-35)
    WAVEFORM_SAMPLING_RATE = 8000
    afp_settings = {"dejavu": {"samplerate": 8000, "n_fft": 512, "n_hop": 256, "fan_value": 3, "amp_min": 50, "peak_neighborhood_size": 10}}


def unique_hash(file_path: str, block_size: int = 2 ** 20) -> str:
    """SHA-1 of the file's bytes, upper-case hex (dejavu.py:45-63)."""
    digest = sha1()
    with open(file_path, "rb") as fh:
        for block in iter(lambda: fh.read(block_size), b""):
            digest.update(block)
    return digest.hexdigest().upper()


def _resample(audio: torch.Tensor, sr_from: int, sr_to: int) -> torch.Tensor:
    if sr_from == sr_to:
        return audio    # torchaudio's Resample(sr, sr) is the identity (dejavu.py:93)
    from torchaudio.transforms import Resample

    return Resample(sr_from, sr_to)(audio)


def read(filename: str, denoising: bool = False, denoising_model: str = "unet") -> Tuple[List[torch.Tensor], int, str]:
    """(channels, sample rate, file hash) of a query pickle or an mp3 (dejavu.py:66-117): mono, scaled by 32767."""
    if denoising is True:
        assert denoising_model in ["demucs", "unet"]
    sr = afp_settings["dejavu"]["samplerate"]
    ext = filename.split(".")[-1]
    if ext == "pkl":
        with open(filename, "rb") as fh:
            audio = torch.tensor(pickle.load(fh)).squeeze(0)
        if denoising is True and denoising_model == "demucs":
            raise NotImplementedError("the Demucs waveform denoiser is outside the B200 query path (DESIGN.md section 6)")
        channels = [_resample(audio, WAVEFORM_SAMPLING_RATE, sr) * 32767]
    elif ext == "mp3":
        import torchaudio

        audio, sr_origin = torchaudio.load(filename)
        channels = [_resample(audio.mean(axis=0), sr_origin, sr) * 32767]
    else:
        raise ValueError(f"unsupported file type: {filename}")
    return channels, sr, unique_hash(filename)


class Dejavu:
    def __init__(self, config, settings, state="set", denoising=False, denoising_model=None):
        self.config = config
        self.settings = settings
        self.db = PostgreSQLDatabase(**config.get("database", {}))
        self.denoising = denoising
        self.denoising_model = denoising_model
        if self.denoising is True:
            assert self.denoising_model in ["unet", "demucs"]
        if state == "set":
            self.db.setup()
        elif state == "clear":
            self.db.empty()
        self._load_fingerprinted_audio_hashes()

    def _load_fingerprinted_audio_hashes(self) -> None:
        self.songs = self.db.get_songs()
        self.songhashes_set = {song["file_sha1"] for song in self.songs}

    # -- indexing ----------------------------------------------------------------------------
    def fingerprint_directory(self, mp3_path_list: list, nprocesses: int = None) -> None:
        """Fingerprint and store every file not indexed yet (dejavu.py:144-216)."""
        for path in mp3_path_list:
            if unique_hash(path) in self.songhashes_set:
                continue
            song_name, hashes, file_hash = self._fingerprint_worker((path, None))
            sid = self.db.insert_song(song_name, file_hash, len(hashes))
            self.db.insert_hashes(sid, hashes)
            self.db.set_song_fingerprinted(sid)
            self._load_fingerprinted_audio_hashes()

    @staticmethod
    def _fingerprint_worker(arguments):
        file_name = arguments[0] if isinstance(arguments, (tuple, list)) else arguments
        song_name, _ = os.path.splitext(os.path.basename(file_name))
        fingerprints, file_hash = Dejavu.get_file_fingerprints(file_name, print_output=False)
        return song_name, fingerprints, file_hash

    @staticmethod
    def get_file_fingerprints(file_name: str, print_output: bool = False):
        channels, fs, file_hash = read(file_name)
        fingerprints = set()
        for n, channel in enumerate(channels, start=1):
            if print_output:
                print(f"Fingerprinting channel {n}/{len(channels)} for {file_name}")
            fingerprints |= set(fingerprint(channel, Fs=fs))
        return fingerprints, file_hash

    # -- query -------------------------------------------------------------------------------
    def generate_fingerprints(self, samples, get_masks: bool = False):
        """dejavu.py:256-293: (hashes, seconds) - or (peak mask, spectrogram) of a FILE when get_masks."""
        fs = self.settings["samplerate"]
        t = time()
        if get_masks:
            channels, _, _ = read(samples, denoising=self.denoising, denoising_model=self.denoising_model)
            _, peak_mask, specgram = fingerprint(channels[0], Fs=fs, denoising=self.denoising,
                                                 denoising_model=self.denoising_model, get_masks=True)
            return peak_mask, specgram
        hashes = fingerprint(samples, Fs=fs, denoising=self.denoising, denoising_model=self.denoising_model, get_masks=False)
        return hashes, time() - t

    def find_matches(self, hashes) -> Tuple["DeviceMatches", Dict[int, int], float]:
        """dejavu.py:295-310.  `matches` behaves as the reference's list of (song id, offset difference) tuples and
        keeps the pairs on the device for align_matches."""
        t = time()
        pairs, dedup = self.db.match_pairs(hashes)
        d = dedup.cpu().numpy()
        dedup_hashes = {int(i): int(d[i]) for i in np.nonzero(d)[0]}
        return DeviceMatches(pairs), dedup_hashes, time() - t

    def align_matches(self, matches, dedup_hashes: Dict[int, int], queried_hashes: int, topn: int = TOPN) -> List[Dict[str, any]]:
        """dejavu.py:312-378: the (song, offset difference) with the most votes; ties go to the smaller song id, then
        the smaller offset (what the reference's stable sorts put first).  topn > 1 is not used by the reference
        (TOPN = 1, variables.py:38) and not built."""
        if topn != 1:
            raise NotImplementedError("align_matches returns the top match only (TOPN = 1 in the reference)")
        if isinstance(matches, DeviceMatches):
            pairs = matches.pairs
        else:
            pairs = torch.tensor(list(matches), dtype=torch.int32).reshape(-1, 2).cuda()
        song_id, offset, votes = self.db._store.index().align(pairs)
        if song_id < 0:
            return []
        song = self.db.get_song_by_id(song_id)
        hashes_matched = dedup_hashes[song_id]
        nseconds = round(float(offset) / self.settings["samplerate"] * self.settings["n_hop"], 5)
        return [{
            SONG_ID: song_id,
            SONG_NAME: song.get(SONG_NAME, None).encode("utf8"),
            INPUT_HASHES: queried_hashes,
            FINGERPRINTED_HASHES: song.get("total_hashes", None),
            HASHES_MATCHED: hashes_matched,
            INPUT_CONFIDENCE: round(hashes_matched / queried_hashes, 2),
            INPUT_CONFIDENCE_2: round(votes / queried_hashes, 2),
            "nb_matches_with_offset": votes,
            FINGERPRINTED_CONFIDENCE: round(hashes_matched / song.get("total_hashes", None), 2),
            OFFSET: offset,
            OFFSET_SECS: nseconds,
            "file_sha1": song.get("file_sha1", None).encode("utf8"),
        }]


class DeviceMatches:
    """The `matches` of find_matches: a lazily materialised list of (song id, offset difference) tuples whose
    int32 [n, 2] tensor stays on the GPU."""

    def __init__(self, pairs: torch.Tensor):
        self.pairs = pairs
        self._list = None

    def _rows(self):
        if self._list is None:
            self._list = [(int(s), int(o)) for s, o in self.pairs.cpu().numpy()]
        return self._list

    def __len__(self):
        return int(self.pairs.shape[0])

    def __iter__(self):
        return iter(self._rows())

    def __getitem__(self, i):
        return self._rows()[i]
