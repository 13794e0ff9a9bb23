"""fissure_segmentation_b200 — B200-native (sm_100a) DGCNN EdgeConv hot path of
kaftanski/fissure-segmentation: kNN graph build, fused EdgeConv forward/backward, Chamfer loss and
the pointops kNN/FPS operators, behind the reference's own Python interfaces.
"""
from . import _lib, ops, pointops_cuda  # noqa: F401
from .chamfer_loss import ChamferLoss, chamfer_distance  # noqa: F401
from .dgcnn import (ConvBlock, DGCNNBase, DGCNNReg, DGCNNSeg, EdgeConv, ImageFeatures,  # noqa: F401
                    SharedFullyConnected, SpatialTransformer, init_weights)
from .knn import create_neighbor_features, knn, pairwise_dist  # noqa: F401
from .modelio import LoadableModel, PointSegmentationModelBase, store_config_args  # noqa: F401
from .ops import KnnGraph  # noqa: F401

__version__ = "0.1.0"
