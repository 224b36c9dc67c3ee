"""The mixed-arithmetic kernel's way out of the FP64 domain (csrc/coreg_lag_roll.cu, roll_segment_mixed), restated in
numpy: the coordinate FMA adds kFracMagic = 1.5 * 2^29, whose ulp is 2^-23, so the LOW word of the double is
round((x - floor_x0) * 2^23) modulo 2^32 -- the row index inside the segment in the bits above bit 22 and the
float32 mantissa of 1 + fraction below. This pins the encoding the CUDA code relies on (CPU test, no GPU)."""
import numpy as np

MAGIC = 805306368.0   # 1.5 * 2^29


def _low_word(offset):
    v = np.asarray(offset, dtype=np.float64) + MAGIC          # one rounding of the exact sum, like the FMA's
    return (v.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _fraction(lo, p):
    """What the kernel computes: validity word and float32 fraction for a pixel expected in row p of the segment."""
    u = lo ^ np.uint32(p << 23)
    frac = ((u | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0))
    return u, frac


def test_magic_has_ulp_2_pow_minus_23():
    assert MAGIC == 1.5 * 2.0 ** 29 and np.spacing(np.float64(MAGIC)) == 2.0 ** -23
    assert np.spacing(np.float64(MAGIC + 511.999)) == 2.0 ** -23     # holds for offsets below 2^9 pixels


def test_fraction_and_row_index_come_out_of_the_low_word():
    rng = np.random.default_rng(7)
    for p in (0, 1, 5, 15, 31):
        s = p + rng.random(200000)                       # offset from the shared floor: row p, fraction in [0, 1)
        u, frac = _fraction(_low_word(s), p)
        true = s - p
        up = np.round(true * 2.0 ** 23) == 2.0 ** 23    # a fraction that rounds up to 1 leaves its cell: flagged
        ok = u < np.uint32(0x00800000)
        assert np.array_equal(ok, ~up)
        assert np.max(np.abs(frac[ok].astype(np.float64) - true[ok])) <= 2.0 ** -24
        assert frac[ok].min() >= 0.0 and frac[ok].max() < 1.0


def test_offsets_outside_the_cell_are_flagged():
    # a floor changed inside the segment (rotated lags): wrong row, negative offsets, far-away coordinates
    for p, s in ((0, 1.25), (3, 2.999), (3, 4.0), (0, -1e-7), (0, -0.75), (7, -3.5), (2, 300.0), (0, 511.5)):
        u, _ = _fraction(_low_word(np.array([s])), p)
        assert u[0] >= 0x00800000, (p, s)
    # an offset less than 2^-24 pixel below the cell rounds to fraction 0 of the cell: the spline is C1 across cell
    # boundaries, so that is the same rounding as anywhere else, not a misclassification
    u, frac = _fraction(_low_word(np.array([-1e-9])), 0)
    assert u[0] < 0x00800000 and frac[0] == 0.0
    # exactly on a cell boundary from above is row p with fraction 0
    u, frac = _fraction(_low_word(np.array([4.0])), 4)
    assert u[0] < 0x00800000 and frac[0] == 0.0


def test_residual_folding_keeps_the_lag_offset_exact():
    # xoff + MAGIC rounds to 2^-23 pixel; the kernel moves the residual into the numerator (roll_lag): check that
    # hi + residual reproduces xoff exactly
    rng = np.random.default_rng(11)
    xoff = -rng.random(1000) * 2048.0
    xm = xoff + MAGIC
    resid = xoff - (xm - MAGIC)
    assert np.all(np.abs(resid) <= 2.0 ** -24) and np.array_equal((xm - MAGIC) + resid, xoff)


def test_quadratic_coordinates_stay_within_a_nanopixel():
    """roll_segment_mixed, MODE 0 (|e| <= 2^-18 over the grid, |he1| < 2.2e-8): along a 16-row segment the coordinate
    numerator(p) / (1 - e(p)) is evaluated as the quadratic q0 + p (q1 + p q2) with
    q1 = h1 * inv0 + n0 * dinv, q2 = h1 * dinv, dinv = he1 (1 + 2 e0), inv0 = 1 + e0 + e0^2.
    Restated in numpy against exact rational arithmetic for the worst cases the mode admits."""
    from fractions import Fraction as F
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(200):
        gny = int(rng.choice([64, 512, 2048, 4096]))
        emax = 2.0 ** -18
        he1 = float(rng.uniform(-1, 1)) * 2 * emax / gny          # e is linear over the grid and bounded by emax
        if abs(he1) >= 2.2e-8:
            continue                                              # the kernel takes the per-pixel series there
        e0 = float(rng.uniform(-1, 1)) * (emax - abs(he1) * 16)
        n0 = float(rng.uniform(-4096, 4096))                      # numerator of the first pixel (pixels)
        h1 = float(rng.uniform(-2, 2))                            # its slope per row
        inv0 = 1.0 + e0 + e0 * e0
        dinv = he1 * (1.0 + 2.0 * e0)
        q0, q1, q2 = n0 * inv0, h1 * inv0 + n0 * dinv, h1 * dinv
        for p in range(16):
            exact = (F(n0) + p * F(h1)) / (1 - (F(e0) + p * F(he1)))
            quad = F(q0) + p * (F(q1) + p * F(q2))
            worst = max(worst, abs(float(quad - exact)))
    # e^3 of the series (5.5e-17 relative = 2e-13 px) + the neglected curvature of the reciprocal along the segment
    assert worst < 1e-9, worst


def test_linear_x_coordinate_stays_within_its_tolerance():
    """roll_segment / roll_segment_mixed with LINX: where |hx1 he1| (P - 1)^2 < kLinXTol = 0.99e-11 (csrc/coreg_lag_roll.cu)
    the x coordinate of the quadratic form drops its p^2 term q2 = hx1 he1 (1 + 2 e0). What that neglects stays below
    1e-11 pixel over a 16-row segment, and the line stays within the quadratic form's own nanopixel of the exact
    numerator / (1 - e). A lag grid of pure CRVAL shifts (config 1: 0.5 arcsec pixels, 30 arcsec lags) qualifies with
    two orders of magnitude to spare; a 0.4 degree rotation does not."""
    from fractions import Fraction as F
    P, tol = 16, 0.99e-11
    rng = np.random.default_rng(5)
    worst_drop, worst_exact, taken = 0.0, 0.0, 0
    for _ in range(400):
        he1 = float(rng.uniform(-1, 1)) * 2.2e-8
        hx1 = float(10.0 ** rng.uniform(-9, -2)) * float(rng.choice([-1, 1]))
        if not abs(hx1 * he1) * (P - 1) ** 2 < tol:
            continue
        taken += 1
        e0 = float(rng.uniform(-1, 1)) * (2.0 ** -18 - abs(he1) * P)
        n0 = float(rng.uniform(-4096, 4096))
        inv0 = 1.0 + e0 + e0 * e0
        dinv = he1 * (1.0 + 2.0 * e0)
        q0, q1, q2 = n0 * inv0, hx1 * inv0 + n0 * dinv, hx1 * dinv
        for p in range(P):
            line = F(q0) + p * F(q1)
            worst_drop = max(worst_drop, abs(float(p * p * F(q2))))
            exact = (F(n0) + p * F(hx1)) / (1 - (F(e0) + p * F(he1)))
            worst_exact = max(worst_exact, abs(float(line - exact)))
    assert taken > 100 and worst_drop < 1e-11 and worst_exact < 1e-9, (taken, worst_drop, worst_exact)
    # config 1: x' - x = shift (1 + O(theta^2)): hx1 ~ shift[px] * theta_x * theta_y-per-row, he1 ~ shift[rad] * pixel[rad]
    arcsec = np.pi / 180 / 3600
    hx1 = (30 / 0.5) * (512 * arcsec) * (0.5 * arcsec)
    he1 = (30 * arcsec) * (0.5 * arcsec)
    assert abs(hx1 * he1) * (P - 1) ** 2 < 1e-2 * tol
    assert not abs(np.sin(np.radians(0.4)) * he1) * (P - 1) ** 2 < tol
