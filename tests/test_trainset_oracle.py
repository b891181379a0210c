"""CPU: the numpy oracle of the training-set preparation (oracle/trainset_numpy.py) against outputs of the reference's own
traindata_fixed / traindata_aligned (tests/golden/trainset.npz, oracle/gen_golden.py --only-trainset), and the host-side
Savitzky-Golay tables against scipy."""
import os

import numpy as np
import pytest

from oracle import trainset_numpy as tn

CASES = {"video1_fixed": dict(fixed=True, n=1, sav=True, L=5, O=2), "synth_fixed": dict(fixed=True, n=2, sav=True, L=7, O=3),
         "synth_fixed_nosav": dict(fixed=True, n=2, sav=False, L=7, O=3), "synth_aligned": dict(fixed=False, n=2, sav=True, L=5, O=2)}


def rel(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_reference_outputs(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "trainset.npz"))
    c = CASES[tag]
    raws = [g["%s/raw%d" % (tag, i)] for i in range(c["n"])]
    tr, te, cl = tn.traindata(raws, c["fixed"], True, 4, c["sav"], c["L"], c["O"], 0.1)
    assert rel(tr, g[tag + "/train"]) <= 1e-13 and rel(te, g[tag + "/test"]) <= 1e-13
    for i in range(c["n"]):
        assert rel(cl[i], g["%s/clean%d" % (tag, i)]) <= 1e-13
    assert not np.isnan(tr).any()


def test_outliers_are_present_in_the_fixtures(golden_dir):
    """the fixtures must exercise the IQR cut-off and both interpolation variants"""
    g = np.load(os.path.join(golden_dir, "trainset.npz"))
    for tag in ("synth_fixed", "synth_aligned"):
        raw = g[tag + "/raw0"]
        xz = (raw.T - raw.mean()) / raw.std()
        assert (np.abs(xz) > 4 * tn.iqr(xz)).sum() > 20


def test_interp_aligned_quirk_matches_numpy_interp():
    """the closed form used by the oracle / the kernels against the literal reference expression (np.interp on repeated xp)"""
    rng = np.random.RandomState(3)
    X = rng.randn(40, 7)
    X[rng.rand(40, 7) < 0.2] = np.nan
    X[:, 4] = np.nan                                  # a marker without any valid sample
    y = X.T.copy()
    nans = np.isnan(y)
    y[nans] = np.interp(nans.nonzero()[0], (~nans).nonzero()[0], y[~nans])
    assert np.array_equal(tn.interp_aligned(X), y.T)


def test_savgol_tables_match_scipy():
    scipy_signal = pytest.importorskip("scipy.signal")
    from vame_b200.create_training import savgol_tables
    rng = np.random.RandomState(0)
    X = rng.randn(3, 200)
    for L, O in ((5, 2), (7, 3), (11, 4)):
        ref = scipy_signal.savgol_filter(X, L, O)
        # (scipy fits the edge polynomials with np.polyfit on positions 0..L-1; its own conditioning error grows with the order)
        assert rel(tn.savgol_rows(X, L, O), ref) <= (1e-13 if L <= 7 else 1e-11)
        c0, h0, t0 = tn.savgol_tables(L, O)
        c1, h1, t1 = savgol_tables(L, O)
        assert np.array_equal(c0, c1) and np.array_equal(h0, h1) and np.array_equal(t0, t1)
