"""B200-native query-side hot path of deezer/musicFPaugment.

AugmentFP degradation chain → magnitude STFT → audfprint peak / landmark /
20-bit-hash extraction (+ Dejavu 2-D peak finder) → landmark-hash matching,
as hand-written sm_100a CUDA kernels behind a C-ABI shared library
(``include/mfpa.h``).  This package is the host-side mirror of the
reference's Python call surface; it holds no CPU fallback — importing
``musicfpaugment_b200.lib`` raises if the CUDA library has not been built.
"""

__version__ = "0.1.0"
