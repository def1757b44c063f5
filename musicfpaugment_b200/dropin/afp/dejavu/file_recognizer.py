"""Drop-in for afp/dejavu/file_recognizer.py: `BaseRecognizer`, `FileRecognizer` (same results dictionary)."""
from __future__ import annotations

import abc
from time import time
from typing import Dict, List, Tuple

import numpy as np

from dejavu.dejavu import afp_settings, read
from dejavu.variables import MIN_HASHES


class BaseRecognizer(object, metaclass=abc.ABCMeta):
    def __init__(self, dejavu):
        self.dejavu = dejavu
        self.Fs = afp_settings["dejavu"]["samplerate"]

    def _recognize(self, *data) -> Tuple[List[Dict[str, any]], int, int, int]:
        """file_recognizer.py:17-34: fingerprint every channel, look the distinct hashes up, take the vote."""
        times, hashes = [], set()
        for channel in data:
            fingerprints, seconds = self.dejavu.generate_fingerprints(channel)
            times.append(seconds)
            hashes |= set(fingerprints)
        matches, dedup_hashes, query_time = self.dejavu.find_matches(hashes)
        t = time()
        final_results = self.dejavu.align_matches(matches, dedup_hashes, len(hashes))
        return final_results, np.sum(times), query_time, time() - t

    @abc.abstractmethod
    def recognize(self) -> Dict[str, any]:
        pass


class FileRecognizer(BaseRecognizer):
    def recognize_file(self, filename: str) -> Dict[str, any]:
        channels, self.Fs, _ = read(filename, denoising=self.dejavu.denoising, denoising_model=self.dejavu.denoising_model)
        t = time()
        matches, fingerprint_time, query_time, align_time = self._recognize(*channels)
        t = time() - t
        is_match = bool(len(matches)) and matches[0]["nb_matches_with_offset"] > MIN_HASHES    # :56-62
        return {"total_time": t, "fingerprint_time": fingerprint_time, "query_time": query_time, "align_time": align_time,
                "results": matches, "match": is_match}

    def recognize(self, filename: str) -> Dict[str, any]:
        return self.recognize_file(filename)
