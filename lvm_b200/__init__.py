"""Importable alias of the ``live-video-magnification_b200/`` package directory (a hyphenated
directory name cannot be imported directly)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "live-video-magnification_b200"))

from .processor import *  # noqa: F401,F403,E402
