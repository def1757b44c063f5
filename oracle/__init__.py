"""TEST INFRASTRUCTURE ONLY — CPU oracle for the query-side hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the *checker* (or as the CPU
baseline being timed), never as the thing shipped.  The product path
(``musicfpaugment_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Contents
--------
``audfprint_np``   numpy restatement of afp/audfprint (stft, find_peaks,
                   peaks2landmarks, landmarks2hashes, HashTable.get_hits,
                   Matcher.match_hashes).
``augment_np``     numpy/torch-free restatement of the AugmentFP arithmetic
                   (julius low-pass FIR, IR FFT convolution, noise mix, gain,
                   clipping quantiles, peak normalisation).
``dejavu_np``      numpy restatement of Dejavu's get_2D_peaks and of fingerprint()'s
                   mlab.specgram front end (pinned against scipy.signal.spectrogram).
``unet_torch``     fp32 torch restatement of the UNet denoiser (training/unet.py), pinned
                   bit-for-bit against the real class (tests/golden/unet.npz).
``synth``          seeded synthetic inputs.
``ref_loader``     imports the *real* reference from /root/reference with
                   stub modules (build container only; it cannot travel).
``make_golden``    generates tests/golden/*.npz by running the real reference.

Parity status: pinned.  The reference ships no tests or golden vectors
(SURVEY.md §4), so the restatements are pinned against outputs of the
reference's own functions executed in the build container
(``oracle/make_golden.py`` → ``tests/golden``), with one exception stated in
``augment_np``: ``julius.lowpass_filter`` (julius 0.2.7, un-vendored
third-party dependency, not installed) is restated from its published
algorithm — that one function is "parity unpinned".
"""
