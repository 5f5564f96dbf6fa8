"""CPU tests: the oracle against the golden fixtures (outputs of the unmodified reference on a B200,
tests/golden/README.md) and known-answer tests of its stages."""
import glob
import os

import numpy as np
import pytest

import ame_logs
import oracle_binding as ob
import synth_frames as sf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
FLD = ("LTx", "LTy", "RTx", "RTy", "LBx", "LBy")


def _mismatches(d, k, costs, cpmvs):
    bad = 0
    for p in range(4):
        m = d["cost_%d_%d" % (k, p)] != costs[p]
        for f in FLD:
            m |= d["cpmv_%d_%d" % (k, p)][f] != cpmvs[p][f]
        bad += int(m.sum())
    return bad


def test_golden_present():
    assert len(GOLDEN) >= 12


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_logs(path):
    d = np.load(path)
    orig, recon, qp, extra = d["orig"], d["recon"], int(d["qp"]), int(d["extra_iter"])
    n = orig.shape[0]
    lists = ob.ref_lists(n)
    for k, (poc, r) in enumerate(ame_logs.pass_list(n)):
        costs, cp = ob.ref_pass(recon[lists[poc - 1][r]], orig[poc - 1], ob.lambda_for(qp, poc), ob.default_opts(extra_iter=extra))
        assert _mismatches(d, k, costs, cp) == 0, (path, poc, r)


def test_fp_switches_do_not_change_stress_decisions():
    """tests/golden/REFERENCE_RUN_SUMMARY.txt: on these inputs neither switch matters."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "stress_ramp_416x240.npz"))
    for fused in (0, 1):
        for cvt in (0, 1):
            costs, cp = ob.ref_pass(d["recon"][0], d["orig"][0], ob.lambda_for(32, 1), ob.default_opts(fused_backsub=fused, cvt_rule=cvt))
            assert _mismatches(d, 0, costs, cp) == 0


# ---------------------------------------------------------------- stage KATs
def test_lambda_schedule():
    # -q 32 -> POC%8==0: QP 33 -> 35.167810, odd: QP 40 -> 78.949063, even: QP 39 -> 70.335619
    assert [ob.delta_qp(32, p) for p in range(1, 10)] == [40, 39, 40, 39, 40, 39, 40, 33, 40]
    assert ob.lambda_for(32, 1) == pytest.approx(78.949063, rel=1e-7)
    assert ob.lambda_for(32, 2) == pytest.approx(70.335619, rel=1e-7)
    assert ob.lambda_for(32, 8) == pytest.approx(35.167810, rel=1e-7)
    exp = {22: (11.077166, 17.583905, 15.665478), 27: (19.737266, 39.474532, 35.16781), 37: (62.661913, 140.671239, 125.323826)}
    for qp, (key, odd, even) in exp.items():
        assert ob.lambda_for(qp, 8) == pytest.approx(key, rel=1e-7)
        assert ob.lambda_for(qp, 3) == pytest.approx(odd, rel=1e-7)
        assert ob.lambda_for(qp, 4) == pytest.approx(even, rel=1e-7)


def test_reference_lists():
    L = ob.ref_lists(64)
    assert L[0] == [0] and L[1] == [1, 0] and L[2] == [2, 1, 0] and L[3] == [3, 2, 1, 0]
    assert L[4] == [4, 3, 2, 0] and L[8] == [8, 7, 6, 0] and L[10] == [10, 9, 8, 0] and L[11] == [11, 10, 8, 0]
    assert L[17] == [17, 16, 8, 0] and L[18] == [18, 16, 8, 0]
    assert sum(len(x) for x in L) == 250


def test_filter_integer_mv_is_copy():
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 1024, size=(64, 64)).astype(np.uint16)
    for (px, py, mx, my) in ((8, 8, 0, 0), (20, 12, 16 * 3, -16 * 2), (0, 0, -16 * 5, -16 * 7), (60, 60, 16 * 9, 16 * 9)):
        pred = ob.predict_4x4(ref, px, py, mx, my)
        ys = np.clip(np.arange(py + my // 16, py + my // 16 + 4), 0, 63)
        xs = np.clip(np.arange(px + mx // 16, px + mx // 16 + 4), 0, 63)
        assert (pred == ref[np.ix_(ys, xs)]).all()


def test_filter_all_phases_against_direct_formula():
    """Appendix A4 evaluated directly in numpy, incl. windows that leave the picture."""
    F = np.array([[0, 0, 0, 64, 0, 0, 0, 0], [0, 1, -3, 63, 4, -2, 1, 0], [0, 1, -5, 62, 8, -3, 1, 0], [0, 2, -8, 60, 13, -4, 1, 0],
                  [0, 3, -10, 58, 17, -5, 1, 0], [0, 3, -11, 52, 26, -8, 2, 0], [0, 2, -9, 47, 31, -10, 3, 0],
                  [0, 3, -11, 45, 34, -10, 3, 0], [0, 3, -11, 40, 40, -11, 3, 0], [0, 3, -10, 34, 45, -11, 3, 0],
                  [0, 3, -10, 31, 47, -9, 2, 0], [0, 2, -8, 26, 52, -11, 3, 0], [0, 1, -5, 17, 58, -10, 3, 0],
                  [0, 1, -4, 13, 60, -8, 2, 0], [0, 1, -3, 8, 62, -5, 1, 0], [0, 1, -2, 4, 63, -3, 1, 0]], dtype=np.int64)
    rng = np.random.default_rng(2)
    Hh, Ww = 40, 48
    ref = rng.integers(0, 1024, size=(Hh, Ww)).astype(np.uint16)
    for trial in range(200):
        px, py = int(rng.integers(0, Ww // 4)) * 4, int(rng.integers(0, Hh // 4)) * 4
        mvx, mvy = int(rng.integers(-400, 400)), int(rng.integers(-400, 400))
        ix, fx, iy, fy = mvx >> 4, mvx & 15, mvy >> 4, mvy & 15
        ys = np.clip(py + iy - 3 + np.arange(11), 0, Hh - 1)
        xs = np.clip(px + ix - 3 + np.arange(11), 0, Ww - 1)
        R = ref[np.ix_(ys, xs)].astype(np.int64)
        T = np.zeros((11, 4), np.int64)
        for c in range(4):
            T[:, c] = ((R[:, c:c + 8] * F[fx]).sum(1) - 32768) >> 2
        P = np.zeros((4, 4), np.int64)
        for r in range(4):
            P[r] = np.clip(((T[r:r + 8] * F[fy][:, None]).sum(0) + 524800) >> 10, 0, 1023)
        assert (ob.predict_4x4(ref, px, py, mvx, mvy) == P).all(), (px, py, mvx, mvy)


def test_satd_known_answers():
    z = np.zeros(16, int)
    assert ob.satd_4x4(z, z) == 0
    # constant difference d: only the DC coefficient 16d, counted at 1/4: ((16|d|>>2)+1)>>1
    assert ob.satd_4x4(np.full(16, 10), z) == ((160 >> 2) + 1) >> 1
    # single impulse: all 16 coefficients have magnitude v
    e = z.copy()
    e[5] = 8
    assert ob.satd_4x4(e, z) == ((15 * 8 + (8 >> 2)) + 1) >> 1
    rng = np.random.default_rng(3)
    Hm = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]])
    for _ in range(50):
        a, b = rng.integers(0, 1024, 16), rng.integers(0, 1024, 16)
        C = Hm @ (a - b).reshape(4, 4) @ Hm.T
        s = int(np.abs(C).sum() - abs(C[0, 0]) + (abs(C[0, 0]) >> 2))
        assert ob.satd_4x4(a, b) == (s + 1) >> 1


def test_affine_bits():
    assert ob.affine_bits(2, (0, 0, 0, 0, 0, 0)) == 4
    assert ob.affine_bits(3, (0, 0, 0, 0, 0, 0)) == 6
    # LT = (16, 0) 1/16-pel = 4 quarter-pel: t = 8 -> 1 + 2*3 = 7 bits; RT predicted by LT -> MVD 0
    assert ob.affine_bits(2, (16, 0, 16, 0, 0, 0)) == 7 + 1 + 1 + 1
    # negative: v = -4 -> t = 9 -> 1 + 2*3
    assert ob.affine_bits(2, (-16, 0, -16, 0, 0, 0)) == 7 + 3
    # large MVD: q(131071) = 2^15 -> t = 2^16 -> 33 bits; q(131068) = 32767 -> t = 65534 -> 31 bits
    assert ob.affine_bits(2, (131071, 0, 131071, 0, 0, 0)) == 33 + 1 + 1 + 1
    assert ob.affine_bits(2, (131068, 0, 131068, 0, 0, 0)) == 31 + 1 + 1 + 1


def test_rate_cost_is_single_precision():
    lam = np.float32(78.949063)
    for bits in range(2, 120):
        assert ob.lib().oracle_rate_cost(bits, lam) == int(np.floor(np.float32(lam * np.float32(bits))))


def test_scale_delta_rules():
    L = ob.lib()
    assert L.oracle_scale_delta(0.0, 1) == 0
    assert L.oracle_scale_delta(0.124, 1) == 0 and L.oracle_scale_delta(0.125, 1) == 4
    assert L.oracle_scale_delta(-0.125, 1) == -4 and L.oracle_scale_delta(-0.124, 1) == 0
    nan, inf = float("nan"), float("inf")
    assert L.oracle_scale_delta(nan, 1) == 0 and L.oracle_scale_delta(nan, 0) == 0
    assert L.oracle_scale_delta(inf, 1) == -4 and L.oracle_scale_delta(inf, 0) == 0
    assert L.oracle_scale_delta(-inf, 1) == 0 and L.oracle_scale_delta(-inf, 0) == 0


def test_solver_known_systems():
    rng = np.random.default_rng(5)
    for n in (4, 6):
        A = rng.integers(-50, 50, size=(n, n)).astype(float)
        A = A @ A.T + np.eye(n) * 10
        x = rng.integers(-5, 5, size=n).astype(float)
        M = np.zeros((7, 7))
        M[1:n + 1, :n] = A
        M[1:n + 1, n] = A @ x
        assert np.allclose(ob.solve(M, n), x, atol=1e-9)
        M0 = np.zeros((7, 7))  # all-zero system: NaNs appear below row 1 but the zero test on M[1][0] resets the result
        assert (ob.solve(M0, n) == 0).all()
        M1 = np.zeros((7, 7))  # rank 1 (e.g. a horizontal ramp): the first zero-pivot row stays finite, so the same test fires
        M1[1, 0] = 100.0
        M1[1, n] = 50.0
        assert (ob.solve(M1, n) == 0).all()


def test_geometry_tables():
    total = 0
    for ha, n in ((0, 201), (1, 284)):
        seen = set()
        for k in range(n):
            g, (x, y, w, h) = ob.cu_geometry(ha, k)
            assert g >= 0 and x % 8 == 0 and y % 8 == 0 and x + w <= 128 and y + h <= 128
            seen.add((g, x, y, w, h))
            total += w * h
        assert len(seen) == n
    assert total == 21 * 128 * 128  # 12 aligned + 9 half-aligned CTU areas


def test_synthetic_generator_is_deterministic():
    a = sf.frame(3, 64, 48)
    b = sf.frame(3, 64, 48)
    assert (a == b).all() and a.dtype == np.uint16 and a.max() <= 1023
    assert int(a.astype(np.int64).sum()) == int(sf.frame(3, 64, 48).astype(np.int64).sum())
