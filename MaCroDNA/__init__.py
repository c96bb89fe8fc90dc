"""Import-name shim: the reference is used as ``from MaCroDNA import MaCroDNA`` (README.md:73)."""
from macrodna_b200 import MaCroDNA  # noqa: F401
