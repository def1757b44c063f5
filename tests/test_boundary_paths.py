"""CPU: the drop-in packages EXTEND the reference's packages instead of shadowing them.

With INTEGRATION.md's path order (repo, drop-ins, reference checkout) every import the three named drivers make
must resolve: the replaced modules to the drop-ins, everything else to the reference
(testing/generate_queries.py:14-16, audfprint_exps.py:10-14, dejavu_exps.py:10-13).  No kernel is launched.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")
REF = os.environ.get("MFPA_REFERENCE", "/root/reference")


def _run(code, with_ref):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, DROPIN] + ([REF] if with_ref else []))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)


def test_dropins_alone_import():
    """Without a reference checkout the replaced modules still import (their own copies of the constants)."""
    r = _run("import afp.audfprint.peak_extractor, afp.audfprint.hash_table, afp.audfprint.audfprint_match, afp.audfprint.stft\n"
             "import afp.dejavu.fingerprint, afp.dejavu.variables, afp.dejavu.dejavu, afp.dejavu.file_recognizer\n"
             "import dejavu.dejavu, dejavu.postgres_database, dejavu.variables, afp.dejavu.postgres_database\n"
             "import augmentation, testing.metrics, augmentation.transformations.colored_noise, augmentation.transformations.band_filters\n"
             "assert dejavu.dejavu is afp.dejavu.dejavu and dejavu.postgres_database is afp.dejavu.postgres_database\n"
             "assert dejavu.variables.SONG_ID == 'song_id' and dejavu.variables.TOPN == 1\n"
             "print('ok')", with_ref=False)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "testing")), reason="reference checkout not mounted")
def test_reference_drivers_import_through_the_dropins():
    code = (
        "import os, testing.generate_queries, testing.audfprint_exps, testing.dejavu_exps\n"
        "import afp.dejavu.dejavu, dejavu.dejavu, testing.parameters, testing.metrics, testing.fma_preprocessing\n"
        f"D, R = {DROPIN!r}, {REF!r}\n"
        "inside = lambda m, root: os.path.abspath(m.__file__).startswith(root + os.sep)\n"
        "import augmentation, afp.audfprint.peak_extractor as pe, afp.dejavu.fingerprint as fp\n"
        "assert all(inside(m, D) for m in (augmentation, pe, fp, afp.dejavu.dejavu, testing.metrics)), 'a replaced module came from the reference'\n"
        "assert all(inside(m, R) for m in (testing.generate_queries, testing.audfprint_exps, testing.dejavu_exps, testing.parameters,\n"
        "                                  testing.fma_preprocessing)), 'a driver module did not come from the reference'\n"
        "assert testing.audfprint_exps.Audfprint_peaks is pe.Audfprint_peaks\n"
        "assert testing.dejavu_exps.Dejavu is afp.dejavu.dejavu.Dejavu and testing.generate_queries.AugmentFP is augmentation.AugmentFP\n"
        "import importlib.util\n"
        "spec = importlib.util.find_spec('augmentation.transformations.pass_filters')\n"
        "assert spec is not None and spec.origin.startswith(R), 'augmentation.* of the reference is shadowed'\n"
        "spec = importlib.util.find_spec('augmentation.transformations.colored_noise')\n"
        "assert spec is not None and spec.origin.startswith(D), 'AddColoredNoise did not come from the drop-ins'\n"
        "spec = importlib.util.find_spec('augmentation.transformations.band_filters')\n"
        "assert spec is not None and spec.origin.startswith(D), 'the band filters did not come from the drop-ins'\n"
        "spec = importlib.util.find_spec('augmentation.transformations.gain')\n"
        "assert spec is not None and spec.origin.startswith(R), 'the unreplaced transforms must stay with the reference'\n"
        "print('ok')")
    r = _run(code, with_ref=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "testing")), reason="reference checkout not mounted")
def test_dropin_constants_equal_the_reference():
    """The drop-ins' copies of the configuration constants (augmentation/constants.py, afp/dejavu/variables.py) are
    written in their own form; every name of the reference modules must carry the same value and type."""
    import importlib.util

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    for rel in ("augmentation/constants.py", "afp/dejavu/variables.py"):
        ours, ref = load(os.path.join(DROPIN, rel), "ours_" + rel[-6:-3]), load(os.path.join(REF, rel), "ref_" + rel[-6:-3])
        names = [n for n in dir(ref) if n.isupper()]
        assert len(names) >= 3
        for n in names:
            a, b = getattr(ours, n), getattr(ref, n)
            assert a == b and type(a) is type(b), (rel, n)
            if isinstance(b, dict):
                assert list(a) == list(b) or set(a) == set(b)
                assert all(type(a[k]) is type(b[k]) for k in b), (rel, n)
