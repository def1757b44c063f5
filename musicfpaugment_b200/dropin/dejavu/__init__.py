"""`dejavu` = `afp.dejavu` under the name the reference's own Dejavu modules import it by
(afp/dejavu/dejavu.py:11-13, file_recognizer.py:6-7, postgres_database.py:6).

`afp.dejavu` is a namespace package: this directory tree supplies `fingerprint`, `variables`,
`postgres_database` (in-memory / GPU index), a reference checkout later on sys.path supplies the
rest (`dejavu`, `file_recognizer`, `database`).  `dejavu.X` must be the SAME module object as
`afp.dejavu.X` (one libmfpa context, one in-memory index), so a meta-path finder aliases the names
instead of importing the files a second time.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import afp.dejavu as _real

__path__ = list(_real.__path__)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, module):
        self._module = module

    def create_module(self, spec):
        return self._module

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("dejavu."):
            return None
        try:
            module = importlib.import_module("afp." + fullname)
        except ModuleNotFoundError as e:
            if e.name == "afp." + fullname:
                return None
            raise
        return importlib.util.spec_from_loader(fullname, _AliasLoader(module))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
