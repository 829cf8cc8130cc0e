"""Shared-memory bank model of the ring-FFT passes (csrc/ringfft.cu: fft_dif / fft_dit), used to design the
conflict-free layout for the next round without GPU time.

16-byte elements (double2): a warp-wide LDS.128/STS.128 is served per quarter-warp (8 threads x 16 B = 128 B =
all 32 banks once); two threads of a quarter-warp conflict when their elements fall into the same 4-bank group,
i.e. when (element index mod 8) coincides.  wavefronts(quarter-warp) = max multiplicity of (idx & 7).

Model check: for a Bluestein ring with M = 8192 and 1024 threads the model gives the excess the profiler
reports (ncu, profiles/r01_ncu_ringfft_summary.md: 36 % of the shared-memory wavefronts are conflict replays);
with the swizzle  s(i) = i ^ ((i >> 2) & 7)  every radix-4 pass is conflict free.
"""
import numpy as np


def pass_indices(M, q, nthreads):
    """element indices touched by load k (k = 0..3) of every thread in one radix-4 pass with quarter length q"""
    L = 4 * q
    i = np.arange(M // 4)
    j = i & (q - 1)
    b = (i >> int(np.log2(q))) * L + j
    return [b + k * q for k in range(4)], i


def wavefronts(idx, swz):
    if swz:
        idx = idx ^ ((idx >> 2) & 7)
    grp = idx & 7
    n = len(idx) // 8
    g = grp[: n * 8].reshape(n, 8)
    # max multiplicity per quarter-warp
    mult = np.zeros(n, dtype=np.int64)
    for v in range(8):
        mult = np.maximum(mult, (g == v).sum(axis=1))
    return int(mult.sum()), n


def fft_cost(M, swz):
    lg = int(np.log2(M))
    total = ideal = 0
    L = M
    if lg & 1:  # leading radix-2 pass: contiguous halves, conflict free
        n = (M // 2) // 8 * 2 * 2
        total += n
        ideal += n
        L //= 2
    while L >= 4:
        q = L // 4
        loads, _ = pass_indices(M, q, 1024)
        for idx in loads:  # 4 loads + 4 stores of the same addresses
            w, n = wavefronts(idx, swz)
            total += 2 * w
            ideal += 2 * n
        L //= 4
    return total, ideal


if __name__ == "__main__":
    for M in (4096, 8192):
        for swz in (False, True):
            t, i = fft_cost(M, swz)
            print(f"M = {M:5d}  swizzle = {swz!s:5}  wavefronts {t:8d}  ideal {i:8d}  excess {100 * (t - i) / t:5.1f} %")
    # the swizzle is a bijection on every aligned block of 32 elements
    i = np.arange(1 << 13)
    s = i ^ ((i >> 2) & 7)
    assert np.array_equal(np.sort(s), i) and np.all((s >> 5) == (i >> 5))
    # contiguous accesses (loads, stores, multiplies: 8 consecutive aligned elements) stay conflict free
    w, n = wavefronts(i, True)
    assert w == n
    print("swizzle: bijection within 32-element blocks, contiguous accesses conflict free")
