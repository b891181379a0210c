"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the k-means step of vame/analysis/pose_segmentation.py:141-143,183-185.

The arithmetic lives in scikit-learn (third-party dependency of the reference, not vendored; the container has 1.9.0).
This file restates the published algorithm that `KMeans(init='k-means++', n_clusters=k, random_state=s, n_init=m)` runs on
dense float32 data, in plain numpy:

* n_init restarts that share ONE RandomState(random_state) stream (restart i continues where restart i-1 stopped); a restart
  replaces the best so far only if its inertia is lower AND its clustering differs (up to a label permutation);
* greedy k-means++ seeding (Arthur & Vassilvitskii 2007): first centre uniform, then 2 + int(ln k) candidates per
  round sampled proportionally to the squared distance to the closest centre (searchsorted on the cumulative sum), the
  candidate that minimises the resulting potential is kept;
* Lloyd iterations on the mean-centred data: E-step argmin (ties to the lowest index), M-step means (empty clusters keep
  their centre here; sklearn relocates them to far points - not hit by the fixtures), stop when the labels repeat or when
  the summed squared centre shift <= tol * mean(var(X, axis=0)); final E-step with the final centres; inertia.

Pinned against sklearn itself by tests/test_oracle_pinned.py::test_kmeans_* on tests/golden/kmeans_blobs.npz
(written by oracle/gen_golden.py with sklearn.cluster.KMeans / kmeans_plusplus)."""
import numpy as np


def _sqdist(X, C):
    """[n, k] squared euclidean distances in float64."""
    X = X.astype(np.float64)
    C = C.astype(np.float64)
    return ((X[:, None, :] - C[None, :, :]) ** 2).sum(-1)


def kmeans_plusplus(X, k, rs):
    """Greedy k-means++ seeding; rs = numpy RandomState.  Returns (centers [k, d], indices [k])."""
    n = X.shape[0]
    trials = 2 + int(np.log(k))
    idx = np.full(k, -1, dtype=np.int64)
    first = rs.choice(n, p=np.full(n, 1.0 / n))
    idx[0] = first
    closest = _sqdist(X, X[first:first + 1])[:, 0].astype(X.dtype)
    pot = float(closest.astype(np.float64).sum())
    for c in range(1, k):
        vals = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(closest.astype(np.float64)), vals)
        np.clip(cand, None, n - 1, out=cand)
        d = _sqdist(X, X[cand]).astype(X.dtype).T                       # [trials, n]
        np.minimum(closest[None, :], d, out=d)
        pots = d.astype(np.float64).sum(1)
        best = int(np.argmin(pots))
        pot = float(pots[best])
        closest = d[best]
        idx[c] = cand[best]
    return X[idx].copy(), idx


def lloyd(X, init, max_iter=300, tol=1e-4):
    """Returns labels int32 [n], centers [k, d] (original space), inertia (float), n_iter."""
    X = np.asarray(X)
    mean = X.mean(axis=0)
    Xc = X - mean
    C = (np.asarray(init) - mean).astype(X.dtype)
    tol_abs = float(np.mean(np.var(X, axis=0)) * tol)
    k = C.shape[0]
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    it = 0
    for it in range(max_iter):
        labels = np.argmin(_sqdist(Xc, C), axis=1).astype(np.int32)
        Cn = C.copy()
        for c in range(k):
            m = labels == c
            if m.any():
                Cn[c] = Xc[m].astype(np.float64).mean(axis=0).astype(X.dtype)
        shift = float(((Cn.astype(np.float64) - C.astype(np.float64)) ** 2).sum())
        C = Cn
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = np.argmin(_sqdist(Xc, C), axis=1).astype(np.int32)
    d = _sqdist(Xc, C)
    inertia = float(d[np.arange(X.shape[0]), labels].sum())
    return labels, (C + mean).astype(X.dtype), inertia, it + 1


def same_clustering(a, b, k):
    """True if the two label vectors describe the same partition up to a permutation of the labels."""
    mapping = np.full(k, -1, dtype=np.int64)
    for x, y in zip(a.tolist(), b.tolist()):
        if mapping[x] == -1:
            mapping[x] = y
        elif mapping[x] != y:
            return False
    return True


def kmeans(X, k, random_state=42, n_init=15, max_iter=300, tol=1e-4):
    """Full fit: (labels, centers, inertia, n_iter) of the best of n_init k-means++ restarts."""
    rs = np.random.RandomState(random_state)
    Xc = X - X.mean(axis=0)                 # seeding runs on the centred data (same distances up to rounding)
    best = None
    for _ in range(n_init):
        _, idx = kmeans_plusplus(Xc, k, rs)
        res = lloyd(X, X[idx], max_iter, tol)
        if best is None or (res[2] < best[2] and not same_clustering(res[0], best[0], k)):
            best = res
    return best
