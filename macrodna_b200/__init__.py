"""macrodna_b200 -- B200-native cell-matching hot path of MaCroDNA behind the reference's API.

``from macrodna_b200 import MaCroDNA`` (or ``from MaCroDNA import MaCroDNA``, the reference's
import line, README.md:73) gives the drop-in class; the numeric work lives in
``libmacrodna_b200.so`` (``include/macrodna_b200.h``).
"""
from .api import MaCroDNA, get_handle, random_test  # noqa: F401

__all__ = ["MaCroDNA", "get_handle", "random_test"]
__version__ = "0.1.0"
