"""Top-level alias so that `import pointops_cuda` (models/pointtransformer/pointops.py:13) resolves to
the B200 implementation when this repository is on sys.path."""
from fissure_segmentation_b200.pointops_cuda import *  # noqa: F401,F403
