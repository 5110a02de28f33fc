"""cppyml: Python bindings of the ML++ clustering algorithms (EM, K-means) on the B200 CUDA backend.

Mirrors the package layout of the reference (cppyml/cppyml/__init__.py:18-21): the extension module
`cppyml.cppyml` holds the submodules.  Only `clustering` is provided here.
"""
from .cppyml import clustering  # noqa: F401
from .cppyml import distributed  # noqa: F401
from .cppyml import __version__, backend  # noqa: F401
from . import utils  # noqa: F401,E402
