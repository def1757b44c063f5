"""Drop-in for augmentation/constants.py (the configuration contract of AugmentFP, constants.py:1-24): the same three
names with the same values, written as a table of (parameter, range, probability) per degradation."""

WAVEFORM_SAMPLING_RATE = 8000
IMPULSE_RESPONSE_DIR = "/workspace/noise_databases/mit_ir_survey/Audio"

# degradation parameter -> (lowest value, highest value); every stage is applied with probability 0.8
_RANGES = {
    "cutoff_freq1": (0.0, 150.0),         # loudspeaker high-pass, Hz
    "snr_in_db": (-10, 10),               # background noise
    "gain_in_db": (-5.0, 5.0),
    "percentile_threshold": (None, 0.01),  # clipping: only an upper bound
    "cutoff_freq2": (3000.0, 3999.0),     # low-pass, Hz
    "cutoff_freq3": (30.0, 150.0),        # microphone high-pass, Hz
}
_STAGES = ("cutoff_freq1", "snr_in_db", "ir_response", "gain_in_db", "percentile_threshold", "cutoff_freq2", "cutoff_freq3")

DEFAULT_PARAMETERS = {f"proba_{stage}": 0.8 for stage in _STAGES}
for _name, (_lo, _hi) in _RANGES.items():
    if _lo is not None:
        DEFAULT_PARAMETERS[f"min_{_name}"] = _lo
    DEFAULT_PARAMETERS[f"max_{_name}"] = _hi
del _name, _lo, _hi
