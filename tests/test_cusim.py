"""The REAL kernel sources on the CPU simulator (tests/cusim): csrc/*.cu compiled by g++ against a functional model of the
CUDA execution model and of the sm_100a instructions the kernels use (TMA with 128B swizzle + out-of-bounds fill,
mbarrier, tensor memory, tcgen05.mma / .commit / .ld, named barriers, warp shuffles), one OS thread per CUDA thread.

What this pins (and what it cannot):
  * control flow, indexing, pipeline phases, tile loops, the grid barrier and the peer handshake of every kernel -- a
    deadlock is reported by the simulator's watchdog instead of hanging a B200 box;
  * arithmetic against the same references the `-m gpu` tests use (the gpu tests themselves are re-run on the simulator);
  * NOT performance, NOT the hardware memory model, NOT descriptor bits the model does not decode. The model itself is
    pinned by the kernels that were validated on B200 in round 1 (forward / dgrad / wgrad GEMMs incl. MN-major operands).
"""
import os
import subprocess
import sys
import threading
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Every kernel-level `-m gpu` test (the whole-step tests pass too - profiles/r1_gpu_suite_on_simulator.log - but take
# 1-15 minutes each:  SSEG_GPU_TESTS_ON_EMULATOR=sim SSEG_TEST_EXPERIMENTAL=1 SSEG_DRY_RUN_SMS=16 CUSIM_SMS=16 python -m
# pytest tests/test_gpu_e2e.py tests/test_gpu_widen_hrnet.py -m gpu -n 7)
_KERNEL_LEVEL = ("test_gpu_igemm or test_gpu_elementwise or test_fused_conv_bn_train_kernel or test_fused_conv_bn_dgrad_kernel or "
                 "test_conv_with_folded_affine_epilogue or test_pair_kernels_against_torch or test_mobilenet_kernels")


def _start_gpu_tests_on_sim(kexpr, sms, extra_env=None, workers=4):
    env = dict(os.environ, SSEG_GPU_TESTS_ON_EMULATOR="sim", SSEG_TEST_EXPERIMENTAL="1", CUSIM_SMS=str(sms),
               SSEG_DRY_RUN_SMS=str(sms), CUSIM_TIMEOUT="120")
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-n", str(workers), "-p", "no:cacheprovider", "-k", kexpr,
           os.path.join(ROOT, "tests", "test_gpu_igemm.py"), os.path.join(ROOT, "tests", "test_gpu_elementwise.py"),
           os.path.join(ROOT, "tests", "test_gpu_widen_hrnet.py"), os.path.join(ROOT, "tests", "test_gpu_e2e.py")]
    cmd = env.get("CUSIM_CMD_PREFIX", "").split() + cmd
    return subprocess.Popen(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def _finish(proc, timeout=1500):
    out, _ = proc.communicate(timeout=timeout)
    assert proc.returncode == 0, out[-4000:]
    return int(out.strip().splitlines()[-1].split(" passed")[0].split()[-1])


def _run_gpu_tests_on_sim(kexpr, sms, extra_env=None, timeout=1500, workers=4):
    return _finish(_start_gpu_tests_on_sim(kexpr, sms, extra_env, workers), timeout)


def test_gpu_tests_pass_on_the_simulator():
    """Two runs side by side. (1) All of test_gpu_igemm.py and test_gpu_elementwise.py (the tests that pass on B200: they pin
    the simulator's model of TMA / UMMA descriptors / tensor memory) plus the kernel-level tests of the code that has not run
    on hardware yet - both cooperative conv+BN kernels at every size (148 simulated SMs, up to 7 resident accumulators per
    CTA), the folded-affine epilogue, the fp32-pair kernels, the depthwise / stem kernels with the MobileNetV2 inference
    schedule, the input transforms and the prefetcher. (2) Two whole training steps through the real kernel sources: the
    default step captured / replayed vs the autograd path, and the SSEG_COOP_BN=1 step (fused conv+BN kernels in both
    directions) against the default schedule."""
    import conftest
    conftest.sim_lib()          # build once, before the two pools ask for it
    whole = _start_gpu_tests_on_sim("test_fused_conv_bn_train_schedule_matches_the_default_step or "
                                    "test_graph_replay_matches_eager_and_autograd_path", sms=16, workers=2)
    kernels = _start_gpu_tests_on_sim(_KERNEL_LEVEL + " or input_transforms or device_prefetcher", sms=148, workers=5)
    assert _finish(kernels) >= 85
    assert _finish(whole, timeout=2400) == 2


def test_cooperative_kernels_with_uneven_tile_distribution_and_the_persistent_gemm():
    """5 'SMs' -> CTAs hold different numbers of resident accumulators; the persistent GEMM variant walks 5 CTAs through
    many tiles and both tensor-memory buffers (passes on B200 with the same switches: profiles/r1_summary.md)."""
    _run_gpu_tests_on_sim("(test_fused_conv_bn_train_kernel and (1-64-128-32 or 3-64-64-32)) or "
                          "(test_fused_conv_bn_dgrad_kernel and (3-32-128-64 or 1-16-256-64))", sms=5)
    _run_gpu_tests_on_sim("test_pointwise_wide_k_and_stats or test_3x3_cout256_cin512 or test_ragged_spatial or test_virtual_concat_3x3 "
                          "or test_dgrad_with_fused_bn_backward_reduce", sms=8,
                          extra_env={"SSEG_IGEMM_PERSISTENT": "2", "SSEG_IGEMM_PERSISTENT_CTAS": "5"})
    # the 128 x 256 tile variant (long-K layers on B200) forced onto the small test shapes: 8 accumulator chunks, the batched
    # y-tile copy of the fused BN-backward epilogue, a ragged last N tile
    _run_gpu_tests_on_sim("test_pointwise_wide_k_and_stats or test_3x3_cout256_cin512 or test_ragged or addend or "
                          "test_dgrad_with_fused_bn_backward_reduce or test_virtual_concat_3x3 or test_wgrad_pointwise or "
                          "test_wgrad_concat or test_wgrad_classifier_padded", sms=8,
                          extra_env={"SSEG_IGEMM_N256_KSTEPS": "1", "SSEG_IGEMM_N256_TILES": "1", "SSEG_WGRAD_N256_TILES": "1"})


# ------------------------------------------------------------------------------------------------ two ranks, one process
@pytest.fixture()
def sim(monkeypatch):
    import conftest
    from mit_semseg.engine import _C, ops
    os.environ.setdefault("CUSIM_SMS", "8")
    lib = conftest.sim_lib()
    monkeypatch.setattr(_C, "lib", lambda: lib)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    return lib


def _arena_ns(arenas, rank):
    from mit_semseg.engine import _C
    bases = (_C.c_void_p * 8)()
    for r, a in enumerate(arenas):
        bases[r] = a.data_ptr()
    return types.SimpleNamespace(bases=bases, world=len(arenas), rank=rank)


def _both_ranks(fn):
    """fn(rank) on two host threads at once: the kernels of the two 'GPUs' must overlap in time, they poll each other's
    flags (ctypes releases the GIL for the duration of a call)."""
    errs = []

    def run(r):
        try:
            fn(r)
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))
    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs


@pytest.mark.parametrize("k,cin,cout,hw", [(1, 64, 128, 32), (3, 64, 64, 16)])
def test_synchronised_conv_bn_kernels_two_ranks_over_shared_arenas(sim, k, cin, cout, hw):
    """world = 2: sseg_conv_bn_train and sseg_conv_dgrad_bn pool their partial sums through the peers' arenas inside the
    kernel (flag handshake after the grid barrier). Two host threads stand in for the two GPUs; the reference is the
    Python restatement of the ABI (tests/abi_emulator.py) run rank by rank over the same arenas."""
    from abi_emulator import EmuLib
    from mit_semseg.engine import _C, ops
    g = torch.Generator().manual_seed(11)
    n, world = 1, 2
    cp = (cout + 7) // 8 * 8
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).bfloat16()
    w2 = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    xs = [torch.randn(n, hw, hw, cin, generator=g).bfloat16() for _ in range(world)]
    count = float(n * hw * hw)
    DATA_F, DATA_B, FLAGS = 0, 1024, 3000

    def fresh():
        st = types.SimpleNamespace()
        st.arenas = [torch.zeros(4096) for _ in range(world)]
        for a in st.arenas:
            a[DATA_F + 2 * cout] = count
        st.y = [torch.full((n, hw, hw, cp), float("nan"), dtype=torch.bfloat16) for _ in range(world)]
        st.a = [torch.full((n, hw, hw, cp), float("nan"), dtype=torch.bfloat16) for _ in range(world)]
        st.vec = [torch.zeros(4, cout) for _ in range(world)]
        st.tmp = [torch.zeros(3, cout) for _ in range(world)]       # tmp mean | tmp var | running_iter[0]
        st.cnt = [torch.zeros(1) for _ in range(world)]
        st.counter = [torch.zeros(4, dtype=torch.int32) for _ in range(world)]
        st.step = [torch.ones(1, dtype=torch.int32) for _ in range(world)]
        st.keep = []
        return st

    def forward(st, r):
        ar = st.arenas[r]
        peer = ops.make_coop_peer(_arena_ns(st.arenas, r), DATA_F, cout, FLAGS, st.step[r])
        bn = ops.make_bn_fused(gamma, beta, 1e-5, 0.1, count, ar[DATA_F:DATA_F + cout], ar[DATA_F + cout:DATA_F + 2 * cout],
                               st.counter[r][:1], st.vec[r][0], st.vec[r][1], st.vec[r][2], st.vec[r][3], peer=peer,
                               tmp_running_mean=st.tmp[r][0], tmp_running_var=st.tmp[r][1], running_iter=st.tmp[r][2][:1],
                               count_out=st.cnt[r])
        st.keep.append((peer, bn))
        geom = ops.make_geom([xs[r]], ops.conv_taps(k, 1))
        ops.conv_bn_train(geom, w2, cout, st.y[r], st.a[r], bn)

    # ---- simulator: both ranks concurrently
    S = fresh()
    _both_ranks(lambda r: forward(S, r))
    # ---- reference: emulator, rank after rank; a second pass after clearing each rank's own sums sees complete arenas
    E = fresh()
    emu = EmuLib()
    import pytest as _pt
    mp = _pt.MonkeyPatch()
    mp.setattr(_C, "lib", lambda: emu)
    try:
        for r in range(world):
            forward(E, r)
        for r in range(world):
            E.arenas[r][DATA_F:DATA_F + 2 * cout] = 0
            E.tmp[r].zero_()
            forward(E, r)
    finally:
        mp.undo()
    for r in range(world):
        # fp32 accumulation order differs between the two restatements: single bf16 roundings may flip
        ys, ye = S.y[r][..., :cout].float(), E.y[r][..., :cout].float()
        assert (ys - ye).abs().max().item() <= 2 ** -7 * ye.abs().max().item() and (ys != ye).float().mean().item() < 0.02
        assert torch.allclose(S.arenas[r][:2 * cout], E.arenas[r][:2 * cout], rtol=2e-3, atol=5e-2)
        assert torch.allclose(S.vec[r], E.vec[r], rtol=2e-3, atol=2e-4)
        assert torch.allclose(S.tmp[r], E.tmp[r], rtol=2e-3, atol=2e-4) and S.cnt[r].item() == world * count
        assert (S.a[r][..., :cout].float() - E.a[r][..., :cout].float()).abs().max().item() <= 2 ** -6 * E.a[r].float()[..., :cout].abs().max().item()
    assert torch.equal(S.vec[0], S.vec[1])                 # pooled in rank order: bit-identical coefficients on both ranks

    # ---- backward twin: dgrad of a consumer conv + this layer's BN backward, sums pooled over the ranks
    cnext = 64
    wn = (torch.randn(cnext, cout, k, k, generator=g) * 0.05).bfloat16()
    wd = torch.zeros(cout, k * k * cnext, dtype=torch.bfloat16)
    dys = [(torch.randn(n, hw, hw, cnext, generator=g) * 0.1).bfloat16() for _ in range(world)]
    dh, dw = ops.conv_taps(k, 1)

    def backward(st, r, lib_is_emu):
        ar = st.arenas[r]
        if not hasattr(st, "wd_ready"):
            ops.prep_conv_weight(wn.float().contiguous(), None, wd, o_pad=cnext)
            st.wd_ready = True
        st.step[r].fill_(2)
        peer = ops.make_coop_peer(_arena_ns(st.arenas, r), DATA_B, cp, FLAGS, st.step[r])
        gd = ops.make_geom([dys[r]], ([-v for v in dh], [-v for v in dw]), tap_koff=[t * cnext for t in range(k * k)])
        st.dx = getattr(st, "dx", [None] * world)
        st.dgb = getattr(st, "dgb", [None] * world)
        st.dx[r] = torch.full((n, hw, hw, cp), float("nan"), dtype=torch.bfloat16)
        st.dgb[r] = torch.zeros(2, cout)
        st.counter[r].zero_()
        st.keep.append((peer, gd))
        ops.conv_dgrad_bn(gd, wd, cout, st.y[r], st.dx[r], st.vec[r][2], st.vec[r][3], st.vec[r][0], st.vec[r][1], count,
                          ar[DATA_B:DATA_B + cout], ar[DATA_B + cp:DATA_B + cp + cout], st.dgb[r][0], st.counter[r][:1],
                          peer=peer, count_dev=st.cnt[r], dbeta_out=st.dgb[r][1])

    emu_prep = EmuLib()
    mp = _pt.MonkeyPatch()
    mp.setattr(_C, "lib", lambda: emu_prep)
    try:
        ops.prep_conv_weight(wn.float().contiguous(), None, wd, o_pad=cnext)
        S.wd_ready = E.wd_ready = True
        # the emulated reference consumes the SIMULATOR's forward results, so only the backward kernels are compared
        for r in range(world):
            E.y[r], E.vec[r], E.cnt[r] = S.y[r].clone(), S.vec[r].clone(), S.cnt[r].clone()
        for r in range(world):
            backward(E, r, True)
        for r in range(world):
            E.arenas[r][DATA_B:DATA_B + 2 * cp] = 0
            backward(E, r, True)
    finally:
        mp.undo()
    _both_ranks(lambda r: backward(S, r, False))
    for r in range(world):
        assert torch.allclose(S.arenas[r][DATA_B:DATA_B + 2 * cp], E.arenas[r][DATA_B:DATA_B + 2 * cp], rtol=1e-4, atol=1e-3)
        assert torch.allclose(S.dgb[r], E.dgb[r], rtol=1e-3, atol=1e-3)
        ref = E.dx[r][..., :cout].float()
        assert torch.isfinite(S.dx[r][..., :cout].float()).all()
        assert (S.dx[r][..., :cout].float() - ref).abs().max().item() <= 2 ** -6 * ref.abs().max().item()
    assert torch.equal(S.dgb[0], S.dgb[1])


def test_watchdog_reports_a_rank_that_waits_for_a_peer_that_never_arrives(sim, monkeypatch):
    """The failure mode that would hang a GPU box: one rank enters the synchronised kernel alone. The simulator aborts the
    launch after CUSIM_TIMEOUT seconds and the C ABI returns an error naming the wait."""
    from mit_semseg.engine import _C, ops
    monkeypatch.setenv("CUSIM_TIMEOUT", "2")
    g = torch.Generator().manual_seed(3)
    cin, cout, hw = 64, 64, 16
    x = torch.randn(1, hw, hw, cin, generator=g).bfloat16()
    w2 = (torch.randn(cout, cin, generator=g) * 0.1).bfloat16()
    arenas = [torch.zeros(1024) for _ in range(2)]
    arenas[0][2 * cout] = hw * hw
    step = torch.ones(1, dtype=torch.int32)
    peer = ops.make_coop_peer(_arena_ns(arenas, 0), 0, cout, 512, step)
    vec, tmp, cnt = torch.zeros(4, cout), torch.zeros(3, cout), torch.zeros(1)
    counter = torch.zeros(4, dtype=torch.int32)
    bn = ops.make_bn_fused(torch.ones(cout), torch.zeros(cout), 1e-5, 0.1, hw * hw, arenas[0][:cout], arenas[0][cout:2 * cout],
                           counter[:1], vec[0], vec[1], vec[2], vec[3], peer=peer, tmp_running_mean=tmp[0],
                           tmp_running_var=tmp[1], running_iter=tmp[2][:1], count_out=cnt)
    y, a = torch.zeros(1, hw, hw, cout, dtype=torch.bfloat16), torch.zeros(1, hw, hw, cout, dtype=torch.bfloat16)
    with pytest.raises(_C.SsegError, match="deadlock"):
        ops.conv_bn_train(ops.make_geom([x], ops.conv_taps(1, 1)), w2, cout, y, a, bn)


# ------------------------------------------------------------------------------------------------ detectors
def _san_env(kind):
    rt = subprocess.run(["gcc", "-print-file-name=lib%s.so" % {"address": "asan", "thread": "tsan"}[kind]], capture_output=True,
                        text=True).stdout.strip()
    if not os.path.isabs(rt) or not os.path.exists(rt):
        pytest.skip("no %s sanitizer runtime" % kind)
    # the sanitizers do not follow ucontext switches: OS thread per CUDA thread instead of the default fiber executor
    env = dict(os.environ, CUSIM_SANITIZE=kind, CUSIM_EXECUTOR="threads", LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0",
               TSAN_OPTIONS="exitcode=66 report_signal_unsafe=0")
    build = subprocess.run([os.path.join(ROOT, "tests", "cusim", "build_sim.sh")], capture_output=True, text=True,
                           env={k: v for k, v in env.items() if k != "LD_PRELOAD"})
    assert build.returncode == 0, build.stderr[-2000:]
    return env, build.stdout.strip().splitlines()[-1]


def _selftest(lib, which, env):
    return subprocess.run(env.get("CUSIM_CMD_PREFIX", "").split() +
                          [sys.executable, os.path.join(ROOT, "tests", "cusim", "selftest_run.py"), lib, which], env=env,
                          capture_output=True, text=True, timeout=300)


def test_detectors_catch_planted_defects():
    """The three failure classes the simulator exists for, each planted in a tiny kernel (tests/cusim/selftest.cu): a
    missing __syncthreads (ThreadSanitizer names both source lines), a wait on an mbarrier phase that never completes
    (watchdog), a store one element past a buffer (AddressSanitizer)."""
    env, lib = _san_env("thread")
    # a dynamic detector: whether a particular execution exposes the race depends on how the host schedules the 64 threads
    # (observed here: silent in roughly one run out of three) - the planted race must show up within a few attempts,
    # the race-free variant below must never be flagged
    import time
    for attempt in range(8):
        out = _selftest(lib, "race", env)
        if "ThreadSanitizer: data race" in out.stderr:
            break
        time.sleep(1.0)
    assert "ThreadSanitizer: data race" in out.stderr and "race_kernel" in out.stderr and "selftest.cu:" in out.stderr, out.stderr[:3000]
    out = _selftest(lib, "norace", env)
    assert "ThreadSanitizer" not in out.stderr and "rc 0 [63.0, 62.0, 61.0]" in out.stdout
    import conftest
    conftest.sim_lib()
    plain = os.path.join(ROOT, "tests", "cusim", "_build", "libsseg_sim.so")
    out = _selftest(plain, "stuck", dict(os.environ, CUSIM_TIMEOUT="2", CUSIM_EXECUTOR="threads"))      # time-out watchdog
    text = out.stdout + out.stderr
    assert "deadlock" in text and "thread 32" in text and "mbarrier" in text and "rc 719" in out.stdout, (out.stdout, out.stderr)
    out = _selftest(plain, "stuck", dict(os.environ, CUSIM_EXECUTOR="fibers"))     # fiber scheduler: found at once
    text = out.stdout + out.stderr
    assert "every live thread is blocked" in text and "thread 32 waits on mbarrier" in text, (out.stdout, out.stderr)
    for ex in ("threads", "fibers"):
        assert "rc 0" in _selftest(plain, "notstuck", dict(os.environ, CUSIM_TIMEOUT="2", CUSIM_EXECUTOR=ex)).stdout
    env, lib = _san_env("address")
    out = _selftest(lib, "oob", env)
    assert "heap-buffer-overflow" in out.stderr and "oob_kernel" in out.stderr


_SANITIZED_QUICK = ("test_pointwise_basic or (test_fused_conv_bn_train_kernel and 3-96-48-16) or "
                    "(test_fused_conv_bn_dgrad_kernel and 3-14-48-96) or test_pair_kernels_against_torch")
_SANITIZED = ("test_pointwise_basic or test_ragged_spatial or test_channel_counts_not_multiple_of_64 or test_wgrad_ragged or "
              "(test_dgrad_with_fused_bn_backward_reduce and 3-38) or "
              "(test_fused_conv_bn_train_kernel and (3-64-64-32 or 3-96-48-16)) or "
              "(test_fused_conv_bn_dgrad_kernel and (3-14-48-96 or 1-16-256-64)) or test_pair_kernels_against_torch or "
              "(test_bn_forward_backward and 64-40) or (test_bilinear and 3-64)")


@pytest.mark.parametrize("kind", ["address", "thread"])
def test_kernels_are_clean_under_the_sanitizers(kind):
    """compute-sanitizer's memcheck / racecheck, CPU edition: every global and shared memory access of the kernels is
    bounds-checked (torch's CPU allocations get red zones), resp. checked for unordered conflicting accesses between CUDA
    threads (barriers, mbarriers and flags are real synchronisation in the simulator, so ordered accesses are silent).
    Default: the kernels that have not run on hardware yet; SSEG_TEST_SANITIZERS_FULL=1: the wider list (2 min each)."""
    env, _ = _san_env(kind)
    log = "/tmp/cusim_%s_%d" % (kind, os.getpid())
    if kind == "thread":
        env["TSAN_OPTIONS"] += " log_path=" + log
    else:
        env["ASAN_OPTIONS"] += " log_path=" + log
    full = os.environ.get("SSEG_TEST_SANITIZERS_FULL", "0") == "1"
    _run_gpu_tests_on_sim(_SANITIZED if full else _SANITIZED_QUICK, sms=5, extra_env=env, timeout=3000)
    import glob
    # one file per process, several reports per file; torch's own OpenMP pool (libgomp is not TSan-aware) shows up too
    reports = [r for f in glob.glob(log + ".*") for r in open(f).read().split("==================")]
    def between_cuda_threads(r):
        """both conflicting accesses come from simulated CUDA threads (a kernel's store followed by torch's own OpenMP
        workers reading the result is reported too: libgomp's hand-over is invisible to the sanitizer)"""
        if kind == "address":
            return "libsseg_sim" in r
        head, _, rest = r.partition("Previous ")
        prev = rest.split("\n\n")[0]
        return "libsseg_sim" in head and "libsseg_sim" in prev
    mine = [r for r in reports if between_cuda_threads(r)]
    for f in glob.glob(log + ".*"):
        os.remove(f)
    assert not mine, mine[0][:3000]


def test_every_kernel_waits_for_its_predecessor_grid():
    """Every kernel is launched with programmatic stream serialization (common.h::launch_k), i.e. it may start while its
    predecessor is still running: each __global__ body must execute griddepcontrol.wait (pdl_sync / pdl_wait) before it
    touches global memory. The simulator ignores PDL, so the presence of the wait is checked on the source."""
    import glob
    import re
    csrc = os.path.join(ROOT, "semantic-segmentation-pytorch_b200", "csrc")
    kernels = 0
    for path in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
        src = open(path).read()
        parts = re.split(r"__global__\s+void", src)
        for body in parts[1:]:
            body = body[:body.find("\n}\n")]
            name = re.search(r"(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)", "", body, count=1)).group(1)
            kernels += 1
            assert "pdl_sync()" in body or "pdl_wait()" in body, "%s::%s has no griddepcontrol.wait" % (os.path.basename(path), name)
    assert kernels >= 40


def test_peer_memory_syncbn_kernels_two_ranks(sim):
    """csrc/peer.cu (the default multi-GPU schedule's SyncBN exchange; ran on 2 and 4 B200s): arenas allocated / opened
    through the library's own IPC entry points, two host threads as the two GPUs, three consecutive 'steps' (the flags carry
    step numbers and are never reset). Reference: the Python restatement of the ABI on the complete arenas."""
    import ctypes
    from abi_emulator import EmuLib
    from mit_semseg.engine import _C, ops
    lib = _C.lib()
    C, world = 96, 2
    g = torch.Generator().manual_seed(4)
    ptrs, handles = [], []
    for r in range(world):        # every rank allocates its arena and exports a handle; the peers open it
        p, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _C.check(lib.sseg_peer_alloc(4 * 8192, ctypes.byref(p), h))
        ptrs.append(p), handles.append(h)
    opened = []
    for r in range(world):
        q = ctypes.c_void_p()
        _C.check(lib.sseg_peer_open(handles[r], ctypes.byref(q)))
        assert q.value == ptrs[r].value        # one process here: the mapping is the allocation itself
        opened.append(q)
    arenas = [torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * 8192).from_address(p.value))) for p in ptrs]
    assert all(float(a.abs().sum()) == 0.0 for a in arenas)        # zero-initialised
    STATS, PART, FLAGS = 0, 1024, 3000
    INBOX, INBOX_B = 4100, 4100 + 2 * world * (2 * C + 2)   # push protocol: 8-byte {value, step} slots per sender
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    steps = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
    state = [dict(rm=torch.zeros(C), rv=torch.ones(C), tm=torch.zeros(C), tv=torch.zeros(C), it=torch.zeros(1)) for _ in range(world)]
    ref_state = [dict((k, v.clone()) for k, v in s.items()) for s in state]
    emu = EmuLib()
    for step_no in range(1, 4):
        for r in range(world):     # this step's partial sums [sum | sqsum | count] and backward partials [s1 | s2raw]
            x = torch.randn(500 + 100 * r, C, generator=g) + 0.3 * r
            arenas[r][STATS:STATS + C], arenas[r][STATS + C:STATS + 2 * C] = x.sum(0), (x * x).sum(0)
            arenas[r][STATS + 2 * C] = x.shape[0]
            arenas[r][PART:PART + 2 * C] = torch.randn(2 * C, generator=g)
        out = [dict(v=torch.zeros(4, C), cnt=torch.zeros(1), tot=torch.zeros(4, C)) for _ in range(world)]
        ll = [dict(v=torch.zeros(4, C), cnt=torch.zeros(1), tot=torch.zeros(4, C)) for _ in range(world)]

        def rank_step(r):
            ns = _arena_ns(arenas, r)
            ops.peer_step(steps[r])
            s = state[r]
            ops.bn_finalize_peer(ns, STATS, FLAGS, steps[r], gamma, beta, 1e-5, 0.1, out[r]["v"][0], out[r]["v"][1], out[r]["v"][2],
                                 out[r]["v"][3], out[r]["cnt"], running=(s["rm"], s["rv"], s["tm"], s["tv"], s["it"]), update_running=True)
            ops.bn_bwd_peer_sum(ns, PART, FLAGS + 8, steps[r], out[r]["tot"][0], out[r]["tot"][1], out[r]["tot"][2], out[r]["tot"][3],
                                mean=out[r]["v"][0], invstd=out[r]["v"][1], s2_raw=True)
            # the push protocol (messages into the peers' inboxes, no flags): bit-identical totals
            ops.bn_finalize_peer(ns, STATS, INBOX, steps[r], gamma, beta, 1e-5, 0.1, ll[r]["v"][0], ll[r]["v"][1], ll[r]["v"][2],
                                 ll[r]["v"][3], ll[r]["cnt"], ll=True)
            ops.bn_bwd_peer_sum(ns, PART, INBOX_B, steps[r], ll[r]["tot"][0], ll[r]["tot"][1], ll[r]["tot"][2], ll[r]["tot"][3],
                                mean=ll[r]["v"][0], invstd=ll[r]["v"][1], s2_raw=True, ll=True)
            # the same exchange folded into the BN-backward apply pass (its own flag slots), against peer_sum + apply
            ops.bn_bwd_apply_peer(ns, PART, FLAGS + 16, steps[r], bw[r]["g"], None, bw[r]["y"], out[r]["v"][0], out[r]["v"][1],
                                  out[r]["v"][2], out[r]["cnt"], bw[r]["dy_fused"], bw[r]["db"], bw[r]["dg"], fshift=out[r]["v"][3],
                                  s2_raw=True)
            ops.bn_bwd_apply(bw[r]["g"], None, bw[r]["y"], out[r]["v"][0], out[r]["v"][1], out[r]["v"][2], out[r]["tot"][0],
                             out[r]["tot"][1], 1.0, bw[r]["dy_ref"], count_dev=out[r]["cnt"], fshift=out[r]["v"][3])
        bw = [dict(g=torch.randn(1, 6, 10, C, generator=g).bfloat16(), y=torch.randn(1, 6, 10, C, generator=g).bfloat16(),
                   dy_fused=torch.zeros(1, 6, 10, C, dtype=torch.bfloat16), dy_ref=torch.zeros(1, 6, 10, C, dtype=torch.bfloat16),
                   db=torch.zeros(C), dg=torch.zeros(C)) for _ in range(world)]
        _both_ranks(rank_step)
        for r in range(world):
            assert torch.equal(ll[r]["v"], out[r]["v"]) and torch.equal(ll[r]["tot"], out[r]["tot"])
            assert ll[r]["cnt"].item() == out[r]["cnt"].item()
            assert torch.equal(bw[r]["dy_fused"], bw[r]["dy_ref"])
            assert torch.allclose(bw[r]["db"], out[r]["tot"][2], rtol=1e-6, atol=1e-7)
            assert torch.allclose(bw[r]["dg"], out[r]["tot"][3], rtol=1e-6, atol=1e-7)
        assert all(int(s) == step_no for s in steps)
        for r in range(world):
            ns = _arena_ns(arenas, r)
            v, cnt, tot, s = torch.zeros(4, C), torch.zeros(1), torch.zeros(4, C), ref_state[r]
            emu.sseg_bn_finalize_peer(ns.bases, world, r, STATS, FLAGS, _C.ptr(steps[r]), _C.ptr(gamma), _C.ptr(beta), 1e-5, 0.1, 1,
                                      _C.ptr(s["rm"]), _C.ptr(s["rv"]), _C.ptr(s["tm"]), _C.ptr(s["tv"]), _C.ptr(s["it"]), _C.ptr(v[0]),
                                      _C.ptr(v[1]), _C.ptr(v[2]), _C.ptr(v[3]), _C.ptr(cnt), C, None)
            emu.sseg_bn_bwd_peer_sum(ns.bases, world, r, PART, FLAGS + 8, _C.ptr(steps[r]), _C.ptr(tot[0]), _C.ptr(tot[1]),
                                     _C.ptr(tot[2]), _C.ptr(tot[3]), _C.ptr(v[0]), _C.ptr(v[1]), 1, C, None)
            assert torch.allclose(out[r]["v"], v, rtol=1e-5, atol=1e-6) and out[r]["cnt"].item() == cnt.item() == 1100
            assert torch.allclose(out[r]["tot"], tot, rtol=1e-4, atol=1e-5)
            for k in ("rm", "rv", "tm", "tv", "it"):
                assert torch.allclose(state[r][k], s[k], rtol=1e-5, atol=1e-6), k
        assert torch.equal(out[0]["v"], out[1]["v"]) and torch.equal(out[0]["tot"], out[1]["tot"])   # pooled in rank order
    for r in range(world):
        _C.check(lib.sseg_peer_close(opened[r]))
    del arenas
    for r in range(world):
        _C.check(lib.sseg_peer_free(ptrs[r]))
