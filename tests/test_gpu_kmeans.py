"""k-means on the latent vectors through the C-ABI against the golden outputs of the reference's own
same_parameterization / individual_parameterization (sklearn KMeans) and against the numpy oracle.

Tolerances: labels are integer output - EXACT agreement with the reference on the committed fixtures;
centres 1e-4 absolute (fp64 accumulation here vs sklearn's fp32 chunk sums), inertia 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_blobs.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def _agree(a, b):
    return float((np.asarray(a) == np.asarray(b)).mean())


def test_lloyd_from_given_centres_matches_sklearn(g):
    from vame_b200.kmeans import DeviceKMeans
    k = int(g["k"][0])
    km = DeviceKMeans(k).fit(g["X"], init=g["pp_init"])
    assert np.array_equal(np.asarray(km.labels_), g["single_labels"])
    assert np.abs(km.cluster_centers_ - g["single_centers"]).max() < 1e-4
    assert abs(km.inertia_ - float(g["single_inertia"][0])) / float(g["single_inertia"][0]) < 1e-5
    assert km.n_iter_ == int(g["single_n_iter"][0])


def test_kmeanspp_seeding_picks_the_same_rows(g):
    from vame_b200.kmeans import DeviceKMeans
    k = int(g["k"][0])
    X = torch.as_tensor(g["X"]).cuda()
    km = DeviceKMeans(k)
    Xd, ws = km._prep(X)
    idx = km._seed(Xd, ws, np.random.RandomState(int(g["pp_seed"][0])))
    assert idx == [int(i) for i in g["pp_idx"]]


def test_same_parameterization_matches_reference(g):
    from vame_b200.pose_segmentation import same_parameterization
    X, sp, k = g["X"], int(g["split"][0]), int(g["k"][0])
    labels, centers, usages = same_parameterization({}, ["a", "b"], [X[:sp], X[sp:]], k, "kmeans")
    assert np.array_equal(np.asarray(labels[0]), g["same_labels0"]) and np.array_equal(np.asarray(labels[1]), g["same_labels1"])
    assert np.abs(centers[0] - g["same_centers"]).max() < 1e-4
    assert np.array_equal(usages[0], g["same_usage0"]) and np.array_equal(usages[1], g["same_usage1"])


def test_individual_parameterization_matches_reference(g):
    from vame_b200.pose_segmentation import individual_parameterization
    X, sp, k = g["X"], int(g["split"][0]), int(g["k"][0])
    cfg = {"random_state_kmeans: ": 42, "n_init_kmeans": 3}
    labels, centers, _ = individual_parameterization(cfg, ["a", "b"], [X[:sp], X[sp:]], k)
    for i in range(2):
        assert np.array_equal(np.asarray(labels[i]), g["ind_labels%d" % i])
        assert np.abs(centers[i] - g["ind_centers%d" % i]).max() < 1e-4


def test_large_n_against_oracle_properties():
    """10^6 x 30 (the C4 embedding size): Lloyd from fixed centres; size-independent checks - every point is assigned to its
    nearest returned centre, centres are the means of their members, inertia equals the summed squared distances."""
    from vame_b200.kmeans import DeviceKMeans
    gen = torch.Generator(device="cuda").manual_seed(3)
    n, d, k = 1_000_000, 30, 15
    cent = torch.randn(k, d, device="cuda", generator=gen)
    X = cent[torch.randint(0, k, (n,), device="cuda", generator=gen)] + 0.8 * torch.randn(n, d, device="cuda", generator=gen)
    km = DeviceKMeans(k, max_iter=25).fit(X, init=X[:k].cpu().numpy())
    lab = km.result.labels_.long()
    C = km.result.cluster_centers_
    d2 = torch.cdist(X, C).pow(2)
    near = d2.argmin(1)
    assert (near == lab).float().mean().item() > 0.9999
    assert abs(d2.gather(1, lab[:, None]).sum().item() - km.inertia_) / km.inertia_ < 1e-4
    if km.n_iter_ < 25:                      # converged: centres are the member means
        means = torch.zeros_like(C).index_add_(0, lab, X) / torch.bincount(lab, minlength=k).clamp(min=1)[:, None]
        assert (means - C).abs().max().item() < 1e-2
