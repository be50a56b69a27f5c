"""-m gpu: SynchronizedBatchNorm{1,2}d called as a module, in the structure of the reference's own tests
(lib/nn/modules/tests/test_sync_batchnorm.py:37-108): the module next to nn.BatchNorm{1,2}d with the same parameters, same
input; outputs, input gradients and running statistics must agree - in training and in evaluation mode, unsynchronised
(F.batch_norm formula) and with the synchronised formula switched on by the replication callback (one device: the pooled
statistics are the local ones; the two-process / two-GPU run is tests/test_gpu_north_star.py).

The reference's tolerance is 1e-3 between two fp32 implementations. The engine's BN kernels store activations in bf16
(fp32 statistics), so inputs here are bf16-representable and outputs / gradients are compared at one bf16 rounding of their
scale (2^-8 relative to the largest magnitude); running statistics, which never pass through bf16, at the reference's 1e-3."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _find_bn(module):
    from mit_semseg.lib.nn import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, SynchronizedBatchNorm1d, SynchronizedBatchNorm2d)):
            return m


def _sync_parameters(bn1, bn2):
    bn1.reset_parameters()
    bn2.reset_parameters()
    if bn1.affine and bn2.affine:
        g = torch.Generator().manual_seed(5)
        bn1.weight.data.copy_(0.5 + torch.rand(bn1.num_features, generator=g))   # non-trivial affine (the reference resets to 1 / 0)
        bn1.bias.data.copy_(torch.randn(bn1.num_features, generator=g) * 0.1)
        bn2.weight.data.copy_(bn1.weight.data)
        bn2.bias.data.copy_(bn1.bias.data)


def _close(a, b, what, rel=2 ** -8, atol=None):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    tol = atol if atol is not None else rel * max(b.abs().max().item(), 1e-3)
    err = (a - b).abs().max().item()
    assert err <= tol, "%s: max |diff| %.3e > %.3e" % (what, err, tol)


def _check(bn1, bn2, inp, is_train, sync_formula=False):
    """test_sync_batchnorm.py:44-65 (_checkBatchNormResult)."""
    bn1.train(mode=is_train)
    bn2.train(mode=is_train)
    inp = inp.bfloat16().float().cuda()
    bn1.cuda()
    bn2.cuda()
    _sync_parameters(_find_bn(bn1), _find_bn(bn2))
    g = torch.Generator().manual_seed(11)
    gout = torch.randn(inp.shape, generator=g).bfloat16().float().cuda()
    in1 = inp.clone().requires_grad_(True)
    out1 = bn1(in1)
    (out1 * gout).sum().backward()
    in2 = inp.clone().requires_grad_(True)
    out2 = bn2(in2)
    (out2 * gout).sum().backward()
    _close(out2, out1, "output")
    _close(in2.grad, in1.grad, "input gradient")
    if is_train and not sync_formula:
        # F.batch_norm's running-statistics update (momentum form); the synchronised formula keeps accumulators instead
        _close(_find_bn(bn2).running_mean, _find_bn(bn1).running_mean, "running_mean", atol=1e-3)
        _close(_find_bn(bn2).running_var, _find_bn(bn1).running_var, "running_var", atol=1e-3)
    if _find_bn(bn1).affine:
        _close(_find_bn(bn2).weight.grad, _find_bn(bn1).weight.grad, "weight gradient", rel=2 ** -7)
        _close(_find_bn(bn2).bias.grad, _find_bn(bn1).bias.grad, "bias gradient", rel=2 ** -7)


def test_sync_batchnorm_normal_train():
    from mit_semseg.lib.nn import SynchronizedBatchNorm1d
    bn, sync_bn = nn.BatchNorm1d(10, momentum=0.1), SynchronizedBatchNorm1d(10, momentum=0.1)
    _check(bn, sync_bn, torch.rand(16, 10, generator=torch.Generator().manual_seed(0)), True)


def test_sync_batchnorm_normal_eval():
    from mit_semseg.lib.nn import SynchronizedBatchNorm1d
    bn, sync_bn = nn.BatchNorm1d(10), SynchronizedBatchNorm1d(10)
    _check(bn, sync_bn, torch.rand(16, 10, generator=torch.Generator().manual_seed(1)), False)


def test_sync_batchnorm_2d_train_and_eval():
    from mit_semseg.lib.nn import SynchronizedBatchNorm2d
    for is_train in (True, False):
        bn, sync_bn = nn.BatchNorm2d(10, momentum=0.1), SynchronizedBatchNorm2d(10, momentum=0.1)
        _check(bn, sync_bn, torch.rand(16, 10, 16, 16, generator=torch.Generator().manual_seed(2)), is_train)


def test_sync_batchnorm_2d_sync_formula_through_the_replication_callback():
    """test_sync_batchnorm.py:98-108 (testSyncBatchNorm2DSyncTrain) on one device: DataParallelWithCallback switches the
    module to the pooled-statistics formula (clamp(var, eps)^-0.5, batchnorm.py:123-139); outputs and input gradients equal
    nn.BatchNorm2d's, the running statistics follow the accumulator recurrence (checked against its closed form)."""
    from mit_semseg.lib.nn import DataParallelWithCallback, SynchronizedBatchNorm2d
    bn = nn.BatchNorm2d(16)
    sync_bn = SynchronizedBatchNorm2d(16)
    wrapped = DataParallelWithCallback(sync_bn.cuda(), device_ids=[0])
    assert sync_bn.is_synchronized()
    x = torch.rand(16, 16, 16, 16, generator=torch.Generator().manual_seed(3))
    _check(bn, wrapped.module, x, True, sync_formula=True)
    xb = x.bfloat16().float()
    mean = xb.transpose(0, 1).reshape(16, -1).mean(1)
    unb = xb.transpose(0, 1).reshape(16, -1).var(1, unbiased=True)
    m = sync_bn.momentum
    it = 1.0 * (1 - m) + 1.0
    _close(sync_bn.running_mean, (0.0 * (1 - m) + mean) / it, "running_mean (accumulator formula)", atol=1e-3)
    _close(sync_bn.running_var, (1.0 * (1 - m) + unb) / it, "running_var (accumulator formula)", atol=1e-3)
