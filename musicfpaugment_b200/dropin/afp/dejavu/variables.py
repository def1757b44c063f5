"""Drop-in for afp/dejavu/variables.py.  Every name the reference module defines is defined here with the same
value: the five numeric settings the peak finder and the hash pairing read (variables.py:18-22) and the string
keys of the result dictionaries that Dejavu / FileRecognizer build (:24-40)."""

# (name, value) of the numeric settings
_SETTINGS = (("CONNECTIVITY_MASK", 2), ("PEAK_NEIGHBORHOOD_SIZE", 10), ("MIN_HASH_TIME_DELTA", 0),
             ("MAX_HASH_TIME_DELTA", 200), ("FINGERPRINT_REDUCTION", 20), ("TOPN", 1), ("MIN_HASHES", 1))

# result-dictionary keys: the constant's name is the key in upper case unless listed otherwise
_KEYS = {k.upper(): k for k in ("offset", "song_id", "song_name", "input_confidence", "input_confidence_2",
                                "fingerprinted_confidence", "results")}
_KEYS.update(OFFSET_SECS="offset_seconds", INPUT_HASHES="input_total_hashes",
             FINGERPRINTED_HASHES="fingerprinted_hashes_in_db", HASHES_MATCHED="hashes_matched_in_input")

globals().update(dict(_SETTINGS), **_KEYS)
__all__ = [name for name, _ in _SETTINGS] + sorted(_KEYS)
