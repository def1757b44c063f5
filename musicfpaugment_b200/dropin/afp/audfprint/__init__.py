"""Drop-in `afp.audfprint`: every module of the reference package (audfprint_match, hash_table,
peak_extractor, stft) has a B200 replacement here; `extend_path` keeps anything else a reference
checkout later on sys.path may add to the package importable."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
