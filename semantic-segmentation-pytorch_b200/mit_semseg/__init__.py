"""b200-semseg: a Blackwell-native engine behind the public model API of MIT CSAIL's semantic-segmentation-pytorch.

`mit_semseg.models.ModelBuilder`, `SegmentationModule`, `mit_semseg.lib.nn.SynchronizedBatchNorm2d`,
`UserScatteredDataParallel` keep the reference's names, signatures and state-dict layout; the compute underneath is
hand-written sm_100a CUDA (libsseg_b200.so, C ABI in include/sseg_b200.h) driven by `mit_semseg.engine`.
"""
__version__ = '1.0.0-b200'
from .models import ModelBuilder, SegmentationModule  # noqa: F401
