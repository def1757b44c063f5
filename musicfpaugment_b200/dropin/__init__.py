"""Drop-in modules with the reference's import paths.

Put this directory ahead of the reference on ``sys.path`` and
``from afp.audfprint.peak_extractor import Audfprint_peaks``,
``from afp.audfprint.hash_table import HashTable``,
``from afp.audfprint.audfprint_match import Matcher``,
``from afp.dejavu.fingerprint import get_2D_peaks`` and ``from augmentation import AugmentFP``
resolve to the B200 path; ``testing/generate_queries.py``, ``audfprint_exps.py`` and
``dejavu_exps.py`` keep their call sites.
"""
import os


def path() -> str:
    return os.path.dirname(os.path.abspath(__file__))
