"""The noise contract (include/dfb_rng_spec.h) as restated on the host: layout, slab-independence,
accuracy of the correctly-rounded-ops-only transform against libm, distribution -- CPU only."""
import math

import numpy as np


def test_pair_transform_matches_libm_to_a_few_ulp(O):
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(2000):
        o = rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32)
        z = O.normal_pair(o)
        U1 = (((int(o[1]) << 32) | int(o[0])) >> 11) | 1
        U2 = ((int(o[3]) << 32) | int(o[2])) >> 11
        r = math.sqrt(-2.0 * math.log(U1 / 2.0 ** 53))
        th = 2.0 * math.pi * (U2 / 2.0 ** 53)
        ref = (r * math.cos(th), r * math.sin(th))
        for a, b in zip(z, ref):
            worst = max(worst, abs(a - b) / max(r, 1e-300))     # error in units of the radius
    assert worst < 8 * 2.3e-16, worst


def test_extreme_uniforms_are_finite(O):
    for o in ([0, 0, 0, 0], [0xFFFFFFFF] * 4, [0xFFFFF800, 0xFFFFFFFF, 0, 0], [0, 0, 0xFFFFFFFF, 0xFFFFFFFF]):
        z = O.normal_pair(np.array(o, dtype=np.uint32))
        assert np.all(np.isfinite(z)) and np.all(np.abs(z) < 8.6)


def test_elements_are_position_addressed(O):
    full = O.noise_elements(5, 3, 2, 1001, 0, 1001)
    for e0, n in ((0, 1), (1, 1), (7, 13), (500, 501), (1000, 1)):
        assert np.array_equal(O.noise_elements(5, 3, 2, 1001, e0, n), full[e0:e0 + n])
    # steps are consecutive blocks of the same stream
    step0 = O.noise_elements(5, 3, 0, 1000, 0, 1000)
    step1 = O.noise_elements(5, 3, 1, 1000, 0, 1000)
    both = O.noise_elements(5, 3, 0, 2000, 0, 2000)
    assert np.array_equal(np.concatenate([step0, step1]), both)


def test_slab_regenerates_the_same_noise(O):
    Ny, Ny_max, NzG = 12, 4, 37            # odd width: pairs straddle rows
    whole = O.noise_rys(11, 2, 1, 3, Ny, Ny_max, NzG)
    for k0, k1 in ((0, 10), (9, 30), (30, 37), (5, 6)):
        assert np.array_equal(O.noise_rys(11, 2, 1, 3, Ny, Ny_max, NzG, k0, k1), whole[:, k0:k1])


def test_distribution(O):
    x = O.noise_elements(2026, 0, 0, 400000, 0, 400000)
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0) < 6e-3
    assert abs(np.mean(x ** 3)) < 2e-2 and abs(np.mean(x ** 4) - 3.0) < 5e-2
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 5e-3 and abs(np.corrcoef(x[::2], x[1::2])[0, 1]) < 6e-3
    # independent streams
    y = O.noise_elements(2026, 1, 0, 400000, 0, 400000)
    assert abs(np.corrcoef(x, y)[0, 1]) < 5e-3
