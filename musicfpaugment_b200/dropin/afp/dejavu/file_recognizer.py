"""Drop-in for afp/dejavu/file_recognizer.py: `BaseRecognizer` / `FileRecognizer` with the reference's call surface
and result dictionary (file_recognizer.py:12-79); the work is in the Dejavu drop-in (GPU fingerprints, lookup, vote)."""
from __future__ import annotations

import abc
import time

import numpy as np

from dejavu.dejavu import afp_settings, read
from dejavu.variables import MIN_HASHES


def _timed(fn, *args):
    start = time.time()
    out = fn(*args)
    return out, time.time() - start


class BaseRecognizer(metaclass=abc.ABCMeta):
    def __init__(self, dejavu):
        self.dejavu, self.Fs = dejavu, afp_settings["dejavu"]["samplerate"]

    def _recognize(self, *data):
        """(ranked matches, seconds fingerprinting, seconds looking up, seconds voting) over all channels; the channels'
        fingerprints are pooled as a set before the lookup (file_recognizer.py:17-34)."""
        pooled, spent = set(), []
        for samples in data:
            prints, seconds = self.dejavu.generate_fingerprints(samples)
            pooled.update(prints)
            spent.append(seconds)
        found, distinct, lookup_seconds = self.dejavu.find_matches(pooled)
        ranked, vote_seconds = _timed(self.dejavu.align_matches, found, distinct, len(pooled))
        return ranked, np.sum(spent), lookup_seconds, vote_seconds

    @abc.abstractmethod
    def recognize(self):
        ...


class FileRecognizer(BaseRecognizer):
    def recognize_file(self, filename: str):
        channels, self.Fs, _ = read(filename, denoising=self.dejavu.denoising, denoising_model=self.dejavu.denoising_model)
        (ranked, t_print, t_lookup, t_vote), total = _timed(self._recognize, *channels)
        # a match needs more than MIN_HASHES aligned hashes on the best song (file_recognizer.py:56-62)
        return dict(total_time=total, fingerprint_time=t_print, query_time=t_lookup, align_time=t_vote, results=ranked,
                    match=bool(ranked) and ranked[0]["nb_matches_with_offset"] > MIN_HASHES)

    recognize = recognize_file
