"""afp/dejavu/variables.py constants used by the peak finder / hash pairing."""
CONNECTIVITY_MASK = 2
PEAK_NEIGHBORHOOD_SIZE = 10
MIN_HASH_TIME_DELTA = 0
MAX_HASH_TIME_DELTA = 200
FINGERPRINT_REDUCTION = 20
