"""Stage the UNMODIFIED reference files of the hot path under the git-ignored baseline/_ref/ (SURVEY 8c), so that
the GPU box - which only receives /root/repo - can run the real reference modules: `bench.py --impl reference`
(`cpu_baseline.kind = "reference"`), the GPU-eager baseline arm and the tests that drive the reference's own
`pointops.py` through this package's `pointops_cuda`.

The reference is a plain Python tree without setup.py / pyproject.toml (pip install is not possible, DESIGN.md
section 8); the files are copied byte for byte, never edited, and never committed (.gitignore: baseline/_ref/).
Run in the build container:  python tools/stage_reference.py      (also called by __graft_entry__.build())
"""
import os
import shutil
import sys

REFERENCE_ROOT = "/root/reference"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
DEST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "models/__init__.py", "models/dgcnn.py", "models/dgcnn_opensrc.py", "models/modelio.py", "models/point_seg_net.py",
    "models/folding_net.py", "models/pointtransformer/__init__.py", "models/pointtransformer/pointops.py",
    "models/pointtransformer/seg_model.py", "utils/__init__.py", "utils/general_utils.py", "utils/model_utils.py",
    "shapes/__init__.py", "shapes/shape_constructor.py",
]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return False
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print("staged %d reference files under %s" % (len(FILES), DEST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
