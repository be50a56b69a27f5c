"""Driver of the simulator self-test kernels (tests/cusim/selftest.cu); run by tests/test_cusim.py in a subprocess."""
import ctypes, sys, os
lib = ctypes.CDLL(sys.argv[1]); which = sys.argv[2]
buf = (ctypes.c_float * 64)()
if which == "race": print("rc", lib.cusim_selftest_race(buf, 0))
if which == "norace": print("rc", lib.cusim_selftest_race(buf, 1), list(buf)[:3])
if which == "stuck": print("rc", lib.cusim_selftest_stuck(1)); lib.cusim_abort_message.restype = ctypes.c_char_p; print(lib.cusim_abort_message())
if which == "notstuck": print("rc", lib.cusim_selftest_stuck(0))
if which == "oob":
    import ctypes.util
    libc = ctypes.CDLL(None); libc.malloc.restype = ctypes.c_void_p
    p = libc.malloc(32 * 4)
    print("rc", lib.cusim_selftest_oob(ctypes.c_void_p(p), 32))
