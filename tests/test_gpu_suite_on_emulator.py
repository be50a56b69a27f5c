"""CPU: the `-m gpu` suite itself, executed against tests/abi_emulator.py (SSEG_GPU_TESTS_ON_EMULATOR=1, see conftest.py).

Two purposes. (1) Tests that pass on the B200 AND here pin the emulator to the kernels' validated behaviour - the same 57
kernel tests and 10 whole-step tests constrain both implementations of the C ABI. (2) The gated tests of code that has not
run on hardware yet (SSEG_TEST_EXPERIMENTAL=1) are debugged here first: their first GPU run then measures the kernels,
not mistakes in the test code (this found a calibration batch that would have crashed on the GPU)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(files, extra=()):
    env = dict(os.environ, SSEG_GPU_TESTS_ON_EMULATOR="1", SSEG_TEST_EXPERIMENTAL="1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-n", "4", *extra,
           *[os.path.join(ROOT, "tests", f) for f in files]]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    tail = "\n".join(out.stdout.splitlines()[-15:])
    assert out.returncode == 0, tail
    return tail


def test_validated_gpu_tests_also_pass_on_the_emulator():
    tail = _run(["test_gpu_igemm.py", "test_gpu_elementwise.py", "test_gpu_e2e.py"],
                ["-k", "not full_config"])          # (the 2x3x512x512 step is too slow for a CPU emulation)
    assert " passed" in tail and "failed" not in tail


def test_gated_and_widening_gpu_tests_pass_on_the_emulator():
    tail = _run(["test_gpu_widen_hrnet.py"])
    assert " passed" in tail and "failed" not in tail and "skipped" not in tail
