"""Run the reference file byte-unmodified in this container.  TEST INFRASTRUCTURE ONLY.

``/root/reference/src/MaCroDNA/macrodna.py`` needs two things this image lacks:

1. ``gurobipy`` (``macrodna.py:2-3``) -- supplied by ``oracle/ref_shim/gurobipy.py``
   (exact LAP via scipy);
2. a pandas that accepts a ``set`` as a ``.loc`` indexer (``macrodna.py:90-91``;
   pandas >= 2 raises ``TypeError``) -- patched here by turning set keys into
   sorted lists inside ``_LocIndexer.__getitem__``.

The reference module is loaded from where it lies (never copied).  This only
works in the build container: ``/root/reference`` does not exist on the GPU box,
so nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.
``oracle/make_golden.py`` uses it to generate the fixtures in ``tests/golden``.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("MACRODNA_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")

VARIANTS = {
    "src": "src/MaCroDNA/macrodna.py",
    "crc": "CRC_data_analysis/macrodna.py",
    "loo": "Resampling_stability_analyses/BE_data_analyses/run_loo_experiment.py",
    "random": "Resampling_stability_analyses/BE_data_analyses/random_assignment_test.py",
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, VARIANTS["src"]))


@contextlib.contextmanager
def _pandas_set_indexer_patch():
    from pandas.core import indexing

    orig = indexing._LocIndexer.__getitem__

    def patched(self, key):
        if isinstance(key, tuple):
            key = tuple(sorted(k) if isinstance(k, (set, frozenset)) else k for k in key)
        elif isinstance(key, (set, frozenset)):
            key = sorted(key)
        return orig(self, key)

    indexing._LocIndexer.__getitem__ = patched
    try:
        yield
    finally:
        indexing._LocIndexer.__getitem__ = orig


def load_reference_module(variant: str = "src"):
    """Import a reference file as a module (its ``__main__`` block does not run), with the shim on the path."""
    path = os.path.join(REFERENCE_ROOT, VARIANTS[variant])
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    sys.path.insert(0, _SHIM_DIR)
    try:
        spec = importlib.util.spec_from_file_location("_reference_module_" + variant, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(_SHIM_DIR)
    return mod


def load_reference_class(variant: str = "src"):
    """Import the reference's ``MaCroDNA`` class from its own file, with the shim on the path."""
    path = os.path.join(REFERENCE_ROOT, VARIANTS[variant])
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    sys.path.insert(0, _SHIM_DIR)
    try:
        spec = importlib.util.spec_from_file_location("_reference_macrodna_" + variant, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(_SHIM_DIR)
    return mod.MaCroDNA


def run_reference(rna_df, dna_df, dna_label=None, variant="src", method="cell2cell_assignment", quiet=True):
    """Call the reference's own method on copies of the frames; returns (result, captured stdout)."""
    cls = load_reference_class(variant)
    obj = cls(rna_df.copy(), dna_df.copy(), None if dna_label is None else dna_label.copy())
    buf = io.StringIO()
    with _pandas_set_indexer_patch():
        if quiet:
            with contextlib.redirect_stdout(buf):
                out = getattr(obj, method)()
        else:
            out = getattr(obj, method)()
    return out, buf.getvalue()


if __name__ == "__main__":
    cls = load_reference_class("src")
    with _pandas_set_indexer_patch():
        cls().tiny_test()
