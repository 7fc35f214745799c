"""The invariant behind the seed loop's "accept without re-deriving the angle" (cube_slam_b200/csrc/cs_lsd.cu, lsd_region_grow; DESIGN.md
section 4), checked on the CPU with the kernel's own arithmetic restated in numpy float32.

The reference tests every neighbour against the region angle as it stands at that neighbour's turn (fastAtan2 of the running float sums,
lsd.cpp:637-688).  The kernel classifies the neighbours of a round against the angle at the START of the round: difference below prec - D
-> accepted whenever its turn comes, above prec + D -> rejected, in between -> the reference's own test at its turn, where
    D = m * (prec + 0.201) / |sums| * 1.001 + 0.0022      (m = neighbours not rejected outright; used only while D <= 0.2)
bounds how far the angle can drift within the round.  Here random rounds are played both ways -- the reference's sequential loop and the
classified loop -- and must accept exactly the same neighbours, in particular for short regions where the bound is loose."""
import numpy as np
import pytest

F = np.float32
DEG = F(0.017453292)


def _wrapped_diff(theta, a):
    """lsd.cpp:1138-1154 on radians (float64, like the reference)."""
    n = abs(theta - a)
    if n > 4.71238898038469:
        n = abs(n - 6.283185307179586)
    return n


def _round_reference(oracle, sx, sy, reg_deg, cand_deg, prec):
    """one round the reference's way; returns accepted mask and the new state"""
    acc = []
    for d in cand_deg:
        theta = float(reg_deg) * (np.pi / 180.0)
        ok = _wrapped_diff(theta, float(d) * (np.pi / 180.0)) <= prec
        acc.append(ok)
        if ok:
            a = float(F(float(d) * (np.pi / 180.0)))        # cos(float(angle)), sin(float(angle)): the per-pixel floats
            sx = F(sx + F(np.cos(a)))
            sy = F(sy + F(np.sin(a)))
            reg_deg = F(oracle.lsd_fast_atan2(float(sy), float(sx)))
    return np.array(acc), sx, sy, reg_deg


def _round_classified(oracle, sx, sy, reg_deg, cand_deg, prec):
    """the kernel's way (float32 screen, drift bound, exact test only for the undecided)"""
    precf = F(prec)
    nf = np.abs(F(reg_deg) * DEG - cand_deg.astype(F) * DEG).astype(F)
    nf = np.where(nf > F(4.712389), np.abs(nf - F(6.2831855)), nf).astype(F)
    inv_l = F(1.0) / np.sqrt(F(sx * sx + sy * sy)) * F(1.001)
    per_add = F((precf + F(0.201)) * inv_l)
    n = len(cand_deg)
    sure = np.zeros(n, bool)
    maybe = np.ones(n, bool)
    D = F(n * per_add + F(0.0022))
    if D <= F(0.2) and precf + D < F(1.5):
        m1 = nf < precf + D
        D = F(m1.sum() * per_add + F(0.0022))
        sure = nf < precf - D
        maybe = nf < precf + D
    acc = np.zeros(n, bool)
    n_exact = 0
    for i in range(n):
        if not maybe[i]:
            continue
        if not sure[i]:
            n_exact += 1
            theta = float(reg_deg) * (np.pi / 180.0)
            if not _wrapped_diff(theta, float(cand_deg[i]) * (np.pi / 180.0)) <= prec:
                continue
        acc[i] = True
        a = float(F(float(cand_deg[i]) * (np.pi / 180.0)))
        sx = F(sx + F(np.cos(a)))
        sy = F(sy + F(np.sin(a)))
        reg_deg = F(oracle.lsd_fast_atan2(float(sy), float(sx)))     # the kernel defers this to the next use; the value is the same
    return acc, sx, sy, reg_deg, n_exact


@pytest.mark.parametrize("spread_deg,prec_deg", [(4.0, 22.5), (12.0, 22.5), (25.0, 22.5), (8.0, 11.25), (20.0, 40.0)])
def test_classified_rounds_accept_what_the_sequential_loop_accepts(oracle, spread_deg, prec_deg):
    rng = np.random.default_rng(int(spread_deg * 10 + prec_deg))
    prec = prec_deg * np.pi / 180.0
    rounds = exact = total = 0
    for trial in range(120):
        base = F(rng.uniform(0, 360))
        a0 = float(F(float(base) * (np.pi / 180.0)))
        sx, sy, reg_deg = F(np.cos(a0)), F(np.sin(a0)), base      # a fresh seed: |sums| = 1, the loosest bound
        sx2, sy2, reg2 = sx, sy, reg_deg
        for _ in range(int(rng.integers(1, 40))):                  # regions of 1 .. ~300 pixels
            n = int(rng.integers(1, 25))
            # neighbours: mostly near the region direction, some near the tolerance on either side, some far away (wrapping through 0 / 360 included)
            kind = rng.random(n)
            off = np.where(kind < 0.6, rng.normal(0, spread_deg, n),
                           np.where(kind < 0.85, rng.choice([-1, 1], n) * (prec_deg + rng.normal(0, 1.5, n)), rng.uniform(-180, 180, n)))
            cand = np.mod(float(reg_deg) + off, 360.0).astype(F)
            want, sx, sy, reg_deg = _round_reference(oracle, sx, sy, reg_deg, cand, prec)
            got, sx2, sy2, reg2, ne = _round_classified(oracle, sx2, sy2, reg2, cand, prec)
            np.testing.assert_array_equal(got, want)
            assert (sx2, sy2, reg2) == (sx, sy, reg_deg)
            rounds += 1
            exact += ne
            total += int(want.sum())
    assert rounds > 1000 and total > 3000
    assert exact < 0.9 * (total + exact)          # the shortcut is actually taken (most accepts of the longer regions skip the exact test)
