"""The shared-memory bank model behind the swizzled FFT work array (tools/proto/fft_bank_model.py): the plain
layout must show the replay excess the profiler measured, the swizzle none, and the swizzle must stay inside
aligned groups of eight elements (what the in-place boundary loops of csrc/ringfft.cu idft_r rely on)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("fft_bank_model", os.path.join(ROOT, "tools", "proto", "fft_bank_model.py"))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


def test_plain_layout_reproduces_measured_excess():
    t, i = model.fft_cost(8192, False)
    assert abs((t - i) / t - 0.364) < 0.005  # ncu: 251 M conflict replays of 689 M wavefronts = 36 %
    t, i = model.fft_cost(4096, False)
    assert abs((t - i) / t - 0.40) < 0.005


def test_swizzle_is_conflict_free_and_local():
    for M in (16, 64, 512, 4096, 8192, 16384):
        t, i = model.fft_cost(M, True)
        assert t == i, M
    idx = np.arange(1 << 15)
    s = idx ^ ((idx >> 2) & 7)
    assert np.array_equal(np.sort(s), idx)
    assert np.all((s >> 3) == (idx >> 3))  # a permutation inside every aligned group of 8


def test_round2_layouts_are_conflict_free():
    """swz2 (half-ring kernels of the power-of-two rings: contiguous, descending, bit-reversed and radix-4 accesses;
    the plain layout is 8-way conflicted on the bit-reversed one) and swz3 (radix-16 passes of the Bluestein array)."""
    assert model.check_round2_layouts()
    for M, lg in ((1024, 10), (4096, 12)):
        idx = np.arange(M)
        assert np.all((model.swz2(idx, lg) >> 3) == (idx >> 3)) and np.all((model.swz3(idx) >> 3) == (idx >> 3))
