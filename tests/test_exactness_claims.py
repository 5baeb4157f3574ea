"""The arithmetic identities the fast march rests on (wx_device.cuh, notes above march_fast), checked on the CPU in
IEEE binary32 with numpy -- independent of any GPU.

(1) floor(p) from a round-down add of 1.5*2^23: the low mantissa bits are the integer.
(2) floor(RN(p / size)) == floor((floor(p) + 1/2) / size) for every finite, non-denormal p in the marching range and
    every integer size: the IEEE quotient cannot round across a lattice plane.
(3) the mask (tx <= ty && tx <= tz, ...) equals (t == min) when no tMax is NaN.
"""
import numpy as np

F = np.float32


def _near_lattice(sizes, rng, per_size=4000):
    """p values within a few ulps of lattice planes k*size (both sides), plus uniform ones, |p| <= 4096 + size."""
    out_p, out_s = [], []
    for s in sizes:
        kmax = int(4096 // s) + 2
        k = rng.integers(-kmax, kmax + 1, per_size).astype(np.float64)
        base = (k * s).astype(F)
        for steps in (-3, -2, -1, 0, 1, 2, 3):
            p = base.copy()
            for _ in range(abs(steps)):
                p = np.nextafter(p, F(np.inf) if steps > 0 else F(-np.inf))
            out_p.append(p), out_s.append(np.full(per_size, s, F))
        out_p.append(rng.uniform(-4096 - s, 4096 + s, per_size).astype(F)), out_s.append(np.full(per_size, s, F))
    return np.concatenate(out_p), np.concatenate(out_s)


def test_quotient_floor_never_crosses_a_lattice_plane():
    rng = np.random.default_rng(1)
    sizes = sorted({d * c for d in list(range(1, 40)) + [63, 64, 100, 127, 255] for c in (1, 8, 128, 4096) if d * c < 2 ** 20})
    p, s = _near_lattice(sizes, rng)
    p = p[np.abs(p) > 1e-30], s[np.abs(p) > 1e-30]
    p, s = p
    ieee = np.floor((p / s).astype(F))                                     # what the reference computes (binary32 division)
    exact = np.floor(p.astype(np.float64) / s.astype(np.float64))          # the true quotient's floor (exact in binary64 here)
    assert np.array_equal(ieee.astype(np.float64), exact)
    via_floor = np.floor((np.floor(p.astype(np.float64)) + 0.5) / s.astype(np.float64))
    assert np.array_equal(via_floor, exact)
    # and the kernel's evaluation of it: fma(x, r, r/2) with r one ulp off 1/size either way, then floor
    x = np.floor(p).astype(F)
    for bump in (-1, 0, 1):
        r = (F(1) / s).astype(F)
        for _ in range(abs(bump)):
            r = np.nextafter(r, F(np.inf) if bump > 0 else F(0))
        q = np.floor((x.astype(np.float64) * r.astype(np.float64) + (r * F(0.5)).astype(np.float64)).astype(F))  # one rounding, like an FMA
        assert np.array_equal(q.astype(np.float64), exact), bump


def test_round_down_add_is_floor():
    rng = np.random.default_rng(2)
    p = np.concatenate([rng.uniform(-2 ** 21, 2 ** 21, 200000), rng.integers(-5000, 5000, 20000).astype(np.float64),
                        np.nextafter(rng.integers(-5000, 5000, 20000).astype(F), F(-np.inf)).astype(np.float64)]).astype(F)
    magic = np.float64(12582912.0)
    # round-toward-minus-infinity of p + magic to binary32 (ulp is 1 in [2^23, 2^24)): floor of the exact sum
    t = np.floor(p.astype(np.float64) + magic)
    assert ((t >= 2 ** 23) & (t < 2 ** 24)).all()
    bits = t.astype(F).view(np.uint32).astype(np.int64)
    assert np.array_equal(bits - 0x4B400000, np.floor(p.astype(np.float64)).astype(np.int64))


def test_mask_equals_min_compare():
    rng = np.random.default_rng(3)
    t = rng.choice(np.array([0.0, -0.0, 1.0, 2.0, 2.0, 3.5, np.inf], F), size=(100000, 3)).astype(F)
    tx, ty, tz = t[:, 0], t[:, 1], t[:, 2]
    ref = np.stack([(tx <= ty) & (tx <= tz), (ty <= tz) & (ty <= tx), (tz <= tx) & (tz <= ty)], 1)
    m = np.minimum(np.minimum(tx, ty), tz)
    mine = np.stack([tx == m, ty == m, tz == m], 1)
    assert np.array_equal(ref, mine)
