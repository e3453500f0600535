"""Importable alias of the ``seed-x_b200/`` package directory (a hyphen is not a legal module name).

``import seedx_b200`` resolves sub-modules from ``<repo>/seed-x_b200``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "seed-x_b200")
__path__.insert(0, _real)  # noqa: F821  (package attribute)

from ._version import __version__  # noqa: E402,F401
