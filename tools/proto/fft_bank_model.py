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


def swz2(i, lg):
    """half-ring kernels of the power-of-two rings: low three bits ^ bits 2..4 ^ the top three bits"""
    return i ^ (((i >> 2) ^ (i >> (lg - 3))) & 7)


def swz3(i):
    """Bluestein work array with the radix-16 passes: low three bits ^ bits 4..6"""
    return i ^ ((i >> 4) & 7)


def plain_wavefronts(idx):
    w, n = wavefronts(idx, False)
    return w, n


def check_round2_layouts():
    """(1) swz2: contiguous, descending, bit-reversed and every radix-4 pass conflict free for M = 1024 .. 4096;
    (2) swz3: contiguous accesses and every element column of every radix-16 pass conflict free."""
    for M in (1024, 2048, 4096):
        lg = int(np.log2(M))
        i = np.arange(M)
        assert np.array_equal(np.sort(swz2(i, lg)), i)
        br = np.array([int(format(x, f"0{lg}b")[::-1], 2) for x in i])
        for pat in (i, i[::-1].copy(), br):
            w, n = plain_wavefronts(swz2(pat, lg))
            assert w == n, (M, w, n)
        w, n = plain_wavefronts(br)
        assert w == 8 * n  # the plain layout: 8-way conflicts on the bit-reversed access
        L = M // 2 if lg & 1 else M
        while L >= 4:
            q = L // 4
            loads, _ = pass_indices(M, q, 256)
            for idx in loads:
                w, n = plain_wavefronts(swz2(idx, lg))
                assert w == n, (M, q, w, n)
            L //= 4
    for M in (256, 4096, 8192):
        i = np.arange(M)
        assert np.array_equal(np.sort(swz3(i)), i)
        w, n = plain_wavefronts(swz3(i))
        assert w == n
        L = M
        while (int(np.log2(L)) & 3) != 0:
            L //= 2
        while L >= 16:
            q, q1 = L // 4, L // 16
            t = np.arange(M // 16)
            base = (t >> int(np.log2(q1))) * L + (t & (q1 - 1))
            for a in range(4):
                for c in range(4):
                    w, n = plain_wavefronts(swz3(base + a * q + c * q1))
                    assert w == n, (M, L, a, c, w, n)
            L //= 16
    return True


if __name__ == "__main__":
    check_round2_layouts()
    print("round-2 layouts: swz2 (half-ring kernels) and swz3 (radix-16 Bluestein) conflict free")
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
