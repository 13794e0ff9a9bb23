"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by tests/golden/make_golden.py to generate fixtures and by the optional oracle-vs-reference
check; never at run time on the GPU box (the reference tree does not travel). The three shims are
the ones SURVEY 8c lists: inspect.getargspec, stub modules for open3d / pytorch3d, stub thop.
"""
import inspect
import os
import sys
from unittest import mock

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load():
    """Returns (models.dgcnn, models.dgcnn_opensrc, utils.general_utils) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = lambda f: inspect.getfullargspec(f)[:4]   # models/modelio.py:27
    for name in ("open3d", "pytorch3d", "pytorch3d.structures", "pytorch3d.transforms", "thop"):
        sys.modules.setdefault(name, mock.MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.dgcnn as ref_dgcnn
    import models.dgcnn_opensrc as ref_opensrc
    import utils.general_utils as ref_utils
    return ref_dgcnn, ref_opensrc, ref_utils
