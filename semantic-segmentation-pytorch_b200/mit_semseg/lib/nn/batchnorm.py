"""SynchronizedBatchNorm{1,2,3}d — parameter/buffer container with the reference's state-dict layout.

Reference: mit_semseg/lib/nn/modules/batchnorm.py:37-139 (vacancy/Synchronized-BatchNorm-PyTorch).
The reference synchronises replicas of ONE process through Python queues (comm.py) and torch.cuda.comm; here every
GPU is its own process. Inside a SegmentationModule step program (mit_semseg/engine/program.py) the [sum, sum^2, count]
message of a layer is pooled over NVLink peer memory by the finalize kernel itself (csrc/peer.cu; NCCL all-reduces with
SSEG_PEER_SYNC=0). Which formula is used follows the reference's switch (batchnorm.py:58):

  * not parallel, or eval  -> F.batch_norm semantics: biased var + eps, running stats with `momentum`
  * parallel and training  -> pooled statistics, clamp(var, eps)^-0.5, accumulator-style running stats

"parallel" = torch.distributed is initialised with world_size > 1 (or the module was marked by
`patch_replication_callback`, see parallel.py).

When the module is called on its own (outside a SegmentationModule program) it runs the same CUDA kernels through
`mit_semseg.engine.functional.batch_norm` (bf16 storage, fp32 statistics; the message is all-reduced with
torch.distributed when synchronised), forward and backward.
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

__all__ = ['SynchronizedBatchNorm1d', 'SynchronizedBatchNorm2d', 'SynchronizedBatchNorm3d']


class _SynchronizedBatchNorm(_BatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.001, affine=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self._is_parallel = False
        self._parallel_id = None
        # the reference's accumulator-style running statistics (batchnorm.py:48-54)
        self._moving_average_fraction = 1. - momentum
        self.register_buffer('_tmp_running_mean', torch.zeros(self.num_features))
        self.register_buffer('_tmp_running_var', torch.ones(self.num_features))
        self.register_buffer('_running_iter', torch.ones(1))

    def is_synchronized(self):
        """True when training statistics must be pooled across devices (reference batchnorm.py:58)."""
        if not self.training:
            return False
        if self._is_parallel:
            return True
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def forward(self, input):
        from ...engine import functional as EF
        return EF.batch_norm(self, input)

    def __data_parallel_replicate__(self, ctx, copy_id):
        # kept for API compatibility with patch_replication_callback (reference batchnorm.py:88-96)
        self._is_parallel = True
        self._parallel_id = copy_id


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError('expected 5D input (got {}D input)'.format(input.dim()))
