"""pcseg_b200 — B200-native (sm_100a) implementation of the PointNet-style segmentation network of
seokjuchung/point-cloud-cnn-segmentation.  Importing this package loads libpcseg_b200.so; there is no
CPU / PyTorch fallback."""
from ._lib import lib as _lib, PcsegError  # noqa: F401  (raises ImportError if the CUDA library is missing)
from .model import PointNetSegmentation, PredictStream, load_checkpoint, f1_scores, lengths_from_masks  # noqa: F401
from .trainer import FusedTrainer  # noqa: F401
from .engine import launch_count  # noqa: F401

__all__ = ["PointNetSegmentation", "load_checkpoint", "f1_scores", "lengths_from_masks", "PredictStream", "FusedTrainer", "PcsegError", "launch_count"]
