"""ORACLE side - test infrastructure, not product code.

Import the UNMODIFIED reference modules of the hot path: from /root/reference in the build container, or from the
byte-for-byte staged copies under the git-ignored baseline/_ref/ (tools/stage_reference.py) on the GPU box, where
/root/reference does not exist. Used by the golden generators, by `bench.py --impl reference` / `cpu_baseline` /
`gpu_eager_baseline` and by the tests that drive the reference's own `pointops.py` through this package's
`pointops_cuda`. The shims are the ones SURVEY 8c lists: inspect.getargspec, stub modules for open3d / pytorch3d /
thop (and thesis.utils, whose only use on this path is a parameter counter).
"""
import inspect
import os
import sys
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(_HERE, "..", "baseline", "_ref"))
STUBS = ("open3d", "pytorch3d", "pytorch3d.structures", "pytorch3d.transforms", "pytorch3d.loss", "thop", "thesis",
         "thesis.utils")


def root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "dgcnn.py")):
            return os.path.abspath(c)
    return None


def available():
    return root() is not None


def _prepare():
    r = root()
    if r is None:
        raise RuntimeError("reference tree not present (neither /root/reference nor baseline/_ref)")
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = lambda f: inspect.getfullargspec(f)[:4]   # models/modelio.py:27
    for name in STUBS:
        sys.modules.setdefault(name, mock.MagicMock())
    if r not in sys.path:
        sys.path.insert(0, r)
    return r


def load():
    """Returns (models.dgcnn, models.dgcnn_opensrc, utils.general_utils) of the reference."""
    _prepare()
    import models.dgcnn as ref_dgcnn
    import models.dgcnn_opensrc as ref_opensrc
    import utils.general_utils as ref_utils
    return ref_dgcnn, ref_opensrc, ref_utils


def load_folding_net():
    """models.folding_net of the reference (DGCNN_Cls_Encoder)."""
    _prepare()
    import models.folding_net as ref_folding
    return ref_folding


def load_pointtransformer():
    """(models.pointtransformer.pointops, models.pointtransformer.seg_model) of the reference. `pointops.py:13` does
    `import pointops_cuda`: the repo root (which holds this package's drop-in module of that name) must be on
    sys.path - it is for everything started from the repo root."""
    _prepare()
    repo = os.path.abspath(os.path.join(_HERE, ".."))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    import models.pointtransformer.pointops as ref_pointops
    import models.pointtransformer.seg_model as ref_seg
    return ref_pointops, ref_seg
