"""Importable alias of the package directory ``r7020e-visual-odometry_b200/`` (whose name is not a
valid Python identifier).  ``import vo_b200.api`` resolves modules from that directory."""
import os as _os

_PKG_DIR = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                                           "r7020e-visual-odometry_b200"))
__path__.insert(0, _PKG_DIR)
__doc__ = open(_os.path.join(_PKG_DIR, "__init__.py")).read().split('"""')[1]

from .api import (Context, SIFTPoints, VoError, detectSIFTFeatures, estworldpose,  # noqa: E402,F401
                  extractFeatures, matchFeatures, match_top2, rigidtform3d, sift_batch,
                  triangulate, default_context)
