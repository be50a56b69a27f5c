from .models import ModelBuilder, SegmentationModule  # noqa: F401
