"""Public model API of the package: the builder and the segmentation wrapper (the engine-backed mirrors in models.py)."""
from . import models as _models

ModelBuilder = _models.ModelBuilder
SegmentationModule = _models.SegmentationModule

__all__ = ["ModelBuilder", "SegmentationModule"]
