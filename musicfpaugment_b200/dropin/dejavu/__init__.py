"""The reference imports the Dejavu modules as `dejavu.*` (afp/dejavu/dejavu.py:11-13)."""
from afp.dejavu import fingerprint, variables  # noqa: F401
