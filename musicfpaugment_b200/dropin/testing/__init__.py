"""Drop-in `testing` package: only `testing.metrics` is replaced (GPU sums instead of the per-peak Python
loops and the torchmetrics dependency); `extend_path` keeps every other module of the reference's package
(`parameters`, `generate_queries`, `audfprint_exps`, `dejavu_exps`, `fma_preprocessing`) importable from a
reference checkout later on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
