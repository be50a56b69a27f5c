from .batchnorm import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d  # noqa: F401
from .parallel import (DataParallelWithCallback, UserScatteredDataParallel, async_copy_to,  # noqa: F401
                       patch_replication_callback, user_scattered_collate)
