"""GPU parity of the training-set preparation (vame_b200/create_training.py -> vame_trainset_* kernels) with the reference's
own outputs (tests/golden/trainset.npz) and with the numpy oracle on larger random inputs.  Everything is float64; the stated
tolerance is 1e-12 max-norm relative (summation order is the only difference)."""
import os

import numpy as np
import pytest
import torch

from oracle import trainset_numpy as tn
from tests.test_trainset_oracle import CASES, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("tag", sorted(CASES))
def test_trainset_matches_reference_outputs(golden_dir, tag):
    from vame_b200 import create_training as ct
    g = np.load(os.path.join(golden_dir, "trainset.npz"))
    c = CASES[tag]
    raws = [g["%s/raw%d" % (tag, i)] for i in range(c["n"])]
    tr, te, cl = ct.trainset_arrays(raws, c["fixed"], True, 4, c["sav"], c["L"], c["O"], 0.1)
    assert tr.dtype == np.float64
    assert rel(tr, g[tag + "/train"]) <= TOL and rel(te, g[tag + "/test"]) <= TOL
    for i in range(c["n"]):
        assert rel(cl[i], g["%s/clean%d" % (tag, i)]) <= TOL


@pytest.mark.parametrize("fixed", [True, False])
def test_trainset_large_random_vs_oracle(fixed):
    from vame_b200 import create_training as ct
    rng = np.random.RandomState(5)
    F, Ns = 24, (20_000, 10_011)
    raws = []
    for N in Ns:
        x = np.cumsum(rng.randn(F, N) * 0.05, axis=1) + rng.randn(F, 1)
        idx = rng.choice(F * N, size=F * N // 200, replace=False)
        x.reshape(-1)[idx] += rng.choice([-1.0, 1.0], size=idx.size) * rng.uniform(20, 60, size=idx.size)
        if not fixed:
            x[5] = 0.5
            x[17] = 0.5
        raws.append(x)
    tr, te, cl = ct.trainset_arrays(raws, fixed, True, 4, True, 9, 3, 0.1)
    tr0, te0, cl0 = tn.traindata(raws, fixed, True, 4, True, 9, 3, 0.1)
    assert rel(tr, tr0) <= TOL and rel(te, te0) <= TOL
    for a, b in zip(cl, cl0):
        assert rel(a, b) <= TOL
    # not robust / no savgol: plain z-score
    tr, te, _ = ct.trainset_arrays(raws[:1], fixed=True, robust=False, savgol=False)
    tr0, te0, _ = tn.traindata(raws[:1], True, False, 4, False)
    assert rel(tr, tr0) <= TOL and rel(te, te0) <= TOL


def test_create_trainset_writes_the_reference_files(tmp_path, golden_dir):
    """file-level drop-in: same names, shapes and dtypes as create_training.py:180-191 writes"""
    import yaml
    from vame_b200 import create_training as ct
    g = np.load(os.path.join(golden_dir, "trainset.npz"))
    proj = tmp_path / "proj"
    for sub in ("data/train", "data/a", "data/b"):
        os.makedirs(proj / sub)
    for i, f in enumerate(("a", "b")):
        np.save(proj / "data" / f / (f + "-PE-seq.npy"), g["synth_fixed/raw%d" % i])
    cfg = dict(project_path=str(proj), video_sets=["a", "b"], all_data="yes", robust=True, iqr_factor=4, egocentric_data=True,
               test_fraction=0.1, num_features=10, savgol_filter=True, savgol_length=7, savgol_order=3)
    with open(proj / "config.yaml", "w") as fh:
        yaml.safe_dump(cfg, fh)
    ct.create_trainset(str(proj / "config.yaml"))
    tr = np.load(proj / "data" / "train" / "train_seq.npy")
    te = np.load(proj / "data" / "train" / "test_seq.npy")
    assert tr.dtype == np.float64 and rel(tr, g["synth_fixed/train"]) <= TOL and rel(te, g["synth_fixed/test"]) <= TOL
    for i, f in enumerate(("a", "b")):
        assert rel(np.load(proj / "data" / f / (f + "-PE-seq-clean.npy")), g["synth_fixed/clean%d" % i]) <= TOL
    with pytest.raises(NotImplementedError):
        ct.create_trainset(str(proj / "config.yaml"), check_parameter=True)
