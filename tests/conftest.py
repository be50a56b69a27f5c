import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semantic-segmentation-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the torch references of the kernel tests are fp32: no TF32 in cuDNN / cuBLAS (their 10-bit mantissa is coarser than
    # several of the tolerances, e.g. 1e-4 on the stem weight gradient)
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


# "1": the Python restatement of the C ABI (tests/abi_emulator.py); "sim": the REAL kernel sources compiled for the CPU
# simulator (tests/cusim: OS thread per CUDA thread, functional TMA / mbarrier / tcgen05 models)
EMULATE_MODE = os.environ.get("SSEG_GPU_TESTS_ON_EMULATOR", "0")
EMULATE = EMULATE_MODE in ("1", "sim")
_SIM_LIB = None


def sim_lib():
    """libsseg_sim.so (built on first use by tests/cusim/build_sim.sh) behind the same ctypes declarations as the product."""
    global _SIM_LIB
    if _SIM_LIB is None:
        import ctypes
        import subprocess
        from mit_semseg.engine import _C
        env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}   # a preloaded sanitizer runtime is for us, not for g++
        out = subprocess.run([os.path.join(ROOT, "tests", "cusim", "build_sim.sh")], capture_output=True, text=True, env=env)
        assert out.returncode == 0, out.stderr[-3000:]
        L = ctypes.CDLL(out.stdout.strip().splitlines()[-1])
        _C._declare(L)
        _SIM_LIB = L
    return _SIM_LIB


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available() or EMULATE:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def install_cuda_stand_in(setattr_, mode):
    """Map every device="cuda" request to the CPU and the library to the emulated ABI (mode "1") or to the simulated kernel
    sources (mode "sim"). `setattr_(obj, name, value)`: monkeypatch.setattr inside tests, plain setattr for a whole process
    (tests/run_reference_script.py runs the reference's unmodified train.py that way)."""
    import functools
    import types
    import torch
    import torch.nn as nn
    from abi_emulator import EmuLib
    from mit_semseg.engine import _C, ops
    lib = sim_lib() if mode == "sim" else EmuLib()
    setattr_(_C, "lib", lambda: lib)
    setattr_(ops, "_stream", lambda: None)

    def remap(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            if str(k.get("device")) .startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped
    for name in ("randn", "rand", "zeros", "ones", "empty", "full", "tensor", "arange", "randint"):
        setattr_(torch, name, remap(getattr(torch, name)))
    real_gen = torch.Generator

    class _CpuGenerator(real_gen):
        def __new__(cls, device="cpu"):
            return real_gen(device="cpu")
    setattr_(torch, "Generator", _CpuGenerator)
    setattr_(torch.cuda, "synchronize", lambda *a, **k: None)
    setattr_(torch.cuda, "is_available", lambda: True)
    setattr_(torch.cuda, "set_device", lambda *a, **k: None)
    import contextlib

    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_stream(self, other):
            pass

        def wait_event(self, ev):
            pass

        def synchronize(self):
            pass
    setattr_(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    setattr_(torch.cuda, "Stream", _Stream)
    setattr_(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    setattr_(torch.cuda, "device", lambda d: contextlib.nullcontext())
    setattr_(torch.cuda, "current_device", lambda: 0)
    setattr_(torch.cuda, "get_device_properties", lambda d=None: types.SimpleNamespace(
        name="stand-in", total_memory=180 << 30, multi_processor_count=148, major=10, minor=0))
    setattr_(torch.Tensor, "record_stream", lambda self, s: None)
    setattr_(torch.Tensor, "cuda", lambda self, *a, **k: self)
    setattr_(torch.Tensor, "is_cuda", property(lambda self: True))
    setattr_(nn.Module, "cuda", lambda self, *a, **k: self)
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, (str, torch.device)) and str(x).startswith("cuda")) else x for x in a)
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)
    setattr_(torch.Tensor, "to", to)
    from mit_semseg.engine import prefetch as PF
    setattr_(PF.DevicePrefetcher, "_use_streams", False)
    # peer arenas (simulator: POSIX shared memory behind the library's IPC entry points) as host tensors
    import ctypes
    import numpy as np
    from mit_semseg.engine import peer as PEER

    def host_view(ptr, count, typestr, device):
        ct = ctypes.c_float if typestr == "<f4" else ctypes.c_int32
        return torch.from_numpy(np.ctypeslib.as_array((ct * count).from_address(ptr)))
    setattr_(PEER.PeerArena, "_view", staticmethod(host_view))
    # step programs: built as schedules on CPU tensors, executed against the emulator, no CUDA graphs / streams
    from mit_semseg.engine import accurate as ACC
    from mit_semseg.engine import program as PR
    for cls in (PR.SegProgram, ACC.AccurateInference):
        real_init = cls.__init__

        def init(self, *a, _real=real_init, **k):
            k["dry_run"] = True
            _real(self, *a, **k)
            self.dry_run, self.serial = False, True
        setattr_(cls, "__init__", init)
        setattr_(cls, "capture", lambda self, warm=True: None)


@pytest.fixture(autouse=True)
def _gpu_tests_on_the_emulator(request, monkeypatch):
    """SSEG_GPU_TESTS_ON_EMULATOR=1 (development aid, no GPU needed): run the `gpu`-marked KERNEL tests against
    tests/abi_emulator.py with every device="cuda" request mapped to the CPU. A test that passes on the B200 and here
    pins the emulator to the kernels' validated behaviour; a gated test that passes here has sound test code."""
    if not (EMULATE and "gpu" in request.keywords):
        yield
        return
    install_cuda_stand_in(monkeypatch.setattr, EMULATE_MODE)
    yield
