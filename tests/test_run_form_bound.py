"""Host-side check of a design bound of the run-recursive y-sweep (csrc/kernels.cu, ysweep_run_kernel, long groups).

Long groups walk the backward sum B against its stable direction,  B_{j+1} = B_j / a + (a^N x_{j+1+N} - x_j / a),  so a rounding
error grows by 1/a = exp(2 pi / N) per row.  csrc/device.cuh bounds the rows of a group by yr_group_cap(N) so that the growth stays
below 8.  This test restates the kernel's recurrences in numpy (same operation order, fp64; numpy has no fused multiply-add, which
only makes the emulation slightly less accurate than the kernel) and checks the deviation from the direct sum evaluated in extended
precision at the bound, for the smallest, typical and largest half-widths.  The cap is read from the header, not copied."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "digital-filtering_b200", "csrc", "device.cuh")


def _cap_from_header():
    src = open(HDR).read()
    jl = int(re.search(r"constexpr int YR_JL = (\d+);", src).group(1))
    nmin = int(re.search(r"constexpr int YR_LOCK_MIN_N = (\d+);", src).group(1))
    num, den = map(int, re.search(r"N \* (\d+) / (\d+) < YR_JL", src).groups())
    yj = int(re.search(r"constexpr int YJ = (\d+);", src).group(1))
    return lambda N: (min(jl, N * num // den) if N >= nmin else yj), nmin, jl


def _lock_step(x, c, N, R):
    a = np.exp(-2.0 * np.pi / N)
    ia = 1.0 / a
    naN1 = -float(np.longdouble(a) ** (N + 1))
    aN = -naN1 * ia
    F = np.zeros(x.shape[1]); B = np.zeros(x.shape[1])
    for i in range(N, -1, -1):                       # the two Horner starts at the group's first row
        F = a * F + x[c - i]
        B = a * B + x[c + i]
    out = [F + B - x[c]]
    for t in range(1, R):
        j = c + t - 1
        F = a * F + (x[j + 1] + naN1 * x[j - N])
        B = ia * B + (aN * x[j + N + 1] - ia * x[j])
        out.append(F + B - x[j + 1])
    return np.stack(out)


def test_group_cap_keeps_the_growth_below_eight():
    cap, nmin, jl = _cap_from_header()
    for N in range(nmin, 513):
        R = cap(N)
        assert 1 <= R <= jl
        assert np.exp(2.0 * np.pi * R / N) <= 8.0 + 1e-9, (N, R)
    assert cap(nmin) >= 8 and cap(128) == jl         # long groups are never shorter than the short ones they replace


def test_lock_step_walk_stays_within_rounding_at_the_bound():
    cap, nmin, _ = _cap_from_header()
    rng = np.random.default_rng(5)
    worst = 0.0
    for N in (nmin, 32, 48, 64, 97, 128, 212, 300):
        R = cap(N)
        x = rng.standard_normal((R + 2 * N + 2, 512))
        c = N
        w = np.longdouble(np.exp(-2.0 * np.pi / N)) ** np.abs(np.arange(-N, N + 1))
        ref = np.stack([(w[:, None] * x[c + t - N:c + t + N + 1].astype(np.longdouble)).sum(0) for t in range(R)])
        got = _lock_step(x, c, N, R)
        dev = float(np.max(np.sqrt(((got - ref) ** 2).mean(1) / (ref ** 2).mean(1))))
        worst = max(worst, dev)
        assert dev < 1e-14, (N, R, dev)              # gate of the product: 1e-12 of the rms
    print("worst rms deviation of the lock-step walk at the cap: %.2e" % worst)
