"""Import the UNMODIFIED reference (rezoo/movis) from /root/reference with stand-ins for the
packages this image lacks.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the oracle restatement and
(b) generate the golden vectors under tests/golden/ (see oracle/gen_golden.py).  /root/reference
does not exist on the GPU box, so nothing that runs there may import this module.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("MOVIS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "PySide6", "PySide6.QtGui", "PySide6.QtCore", "PySide6.QtWidgets",
    "imageio", "imageio.core", "imageio.core.format", "soundfile", "librosa",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "movis"))


def import_reference():
    """Returns the reference ``movis`` module (and makes ``movis.contrib.segmentation`` importable)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock(name=name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import movis  # noqa
    import movis.contrib.segmentation  # noqa
    return movis
