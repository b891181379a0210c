"""k-means on the latent vectors: the device-side replacement of the scikit-learn calls at
vame/analysis/pose_segmentation.py:141-143 (``KMeans(init='k-means++', n_clusters=states, random_state=42, n_init=20)``)
and :183-185 (``KMeans(init='k-means++', n_clusters=cluster, random_state=cfg[...], n_init=cfg['n_init_kmeans'])``).

Same algorithm and the same numpy ``RandomState`` stream as scikit-learn (1.9): the restarts share one stream, every restart
is a greedy k-means++ seeding (2 + int(ln k) candidates per round) followed by Lloyd iterations; all passes over the data
(candidate distances, fp64 cumulative sum + searchsorted, E/M steps) are CUDA kernels behind the C-ABI (``vame_kmeans_*``).
Only the random draws (a handful of doubles per seeding round) are made on the host.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from ._lib import VameB200Error


class KMeansResult:
    def __init__(self, labels, centers, inertia, n_iter):
        self.labels_ = labels
        self.cluster_centers_ = centers
        self.inertia_ = inertia
        self.n_iter_ = n_iter


def _same_clustering(a, b, k):
    """Equal partitions up to a label permutation (device): the k x k contingency table has one non-zero per row."""
    table = torch.bincount(a.long() * k + b.long(), minlength=k * k).view(k, k)
    return bool(((table > 0).sum(1) <= 1).all().item())


class DeviceKMeans:
    """``fit`` / ``predict`` on a CUDA tensor of shape (n, dim), dim <= 64, n_clusters <= 128."""

    def __init__(self, n_clusters, random_state=42, n_init=15, max_iter=300, tol=1e-4):
        self.k = int(n_clusters)
        self.random_state = random_state
        self.n_init = int(n_init)
        self.max_iter = int(max_iter)
        self.tol = float(tol)
        self.result = None

    # ---- plumbing ----------------------------------------------------------------------------------------------------
    def _prep(self, X):
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32))
        if not X.is_cuda:
            if not torch.cuda.is_available():
                raise VameB200Error("vame_b200.kmeans needs a CUDA device (no CPU fallback)")
            X = X.cuda()
        X = X.float().contiguous()
        n, dim = X.shape
        if self.k > n:
            raise ValueError("n_samples=%d should be >= n_clusters=%d" % (n, self.k))
        lib = L.lib()
        nbytes = lib.vame_kmeans_workspace_bytes(n, dim, self.k)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
        return X, ws

    def _seed(self, Xc, ws, rs):
        """Greedy k-means++ on the (centred) data; returns the chosen row indices (host list)."""
        lib = L.lib()
        n, dim = Xc.shape
        k = self.k
        trials = 2 + int(np.log(k))
        dev = Xc.device
        first = int(rs.choice(n, p=np.full(n, 1.0 / n)))
        idx = [first]
        cand = torch.tensor([first], dtype=torch.int64, device=dev)
        newmin = torch.empty((max(trials, 1), n), dtype=torch.float32, device=dev)
        pot = torch.zeros(8, dtype=torch.float64, device=dev)
        L.check(lib.vame_kmeans_candidates(L.ptr(Xc), n, dim, L.ptr(cand), 1, None, L.ptr(newmin), L.ptr(pot), L.cur_stream()),
                "vame_kmeans_candidates")
        closest = newmin[0].clone()
        cur_pot = float(pot[0].item())
        cand = torch.empty(trials, dtype=torch.int64, device=dev)
        for _ in range(1, k):
            vals = torch.as_tensor(rs.uniform(size=trials) * cur_pot, dtype=torch.float64).to(dev)
            L.check(lib.vame_kmeans_sample(L.ptr(closest), n, L.ptr(vals), trials, L.ptr(cand), L.ptr(ws), ws.numel(), L.cur_stream()),
                    "vame_kmeans_sample")
            L.check(lib.vame_kmeans_candidates(L.ptr(Xc), n, dim, L.ptr(cand), trials, L.ptr(closest), L.ptr(newmin), L.ptr(pot),
                                               L.cur_stream()), "vame_kmeans_candidates")
            pots = pot[:trials].cpu().numpy()
            best = int(np.argmin(pots))
            cur_pot = float(pots[best])
            closest.copy_(newmin[best])
            idx.append(int(cand[best].item()))
        return idx

    def _lloyd(self, X, ws, init):
        lib = L.lib()
        n, dim = X.shape
        labels = torch.empty(n, dtype=torch.int32, device=X.device)
        centers = torch.empty((self.k, dim), dtype=torch.float32, device=X.device)
        inertia = torch.zeros(1, dtype=torch.float64, device=X.device)
        n_iter = ctypes.c_int(0)
        L.check(lib.vame_kmeans_lloyd(L.ptr(X), n, dim, self.k, L.ptr(init), self.max_iter, self.tol, L.ptr(labels), L.ptr(centers),
                                      L.ptr(inertia), ctypes.byref(n_iter), L.ptr(ws), ws.numel(), L.cur_stream()), "vame_kmeans_lloyd")
        return labels, centers, float(inertia.item()), int(n_iter.value)

    # ---- sklearn-like surface ------------------------------------------------------------------------------------------
    def fit(self, X, init=None):
        """``init``: optional (k, dim) array -> a single Lloyd run from these centres (KMeans(init=array, n_init=1))."""
        X, ws = self._prep(X)
        if init is not None:
            init_t = torch.as_tensor(np.ascontiguousarray(init, dtype=np.float32)).to(X.device) if not isinstance(init, torch.Tensor) \
                else init.to(X.device).float().contiguous()
            self.result = KMeansResult(*self._lloyd(X, ws, init_t))
            return self
        rs = self.random_state if isinstance(self.random_state, np.random.RandomState) else np.random.RandomState(self.random_state)
        Xc = (X - X.mean(dim=0)).contiguous()          # sklearn seeds on the centred copy (same distances up to rounding)
        best = None
        for _ in range(self.n_init):
            idx = self._seed(Xc, ws, rs)
            init_t = X[torch.tensor(idx, device=X.device)].contiguous()
            res = self._lloyd(X, ws, init_t)
            if best is None or (res[2] < best[2] and not _same_clustering(res[0], best[0], self.k)):
                best = res
        self.result = KMeansResult(*best)
        return self

    def predict(self, X):
        if self.result is None:
            raise VameB200Error("DeviceKMeans.predict before fit")
        X, ws = self._prep(X)
        n, dim = X.shape
        labels = torch.empty(n, dtype=torch.int32, device=X.device)
        L.check(L.lib().vame_kmeans_assign(L.ptr(X), n, dim, self.k, L.ptr(self.result.cluster_centers_), L.ptr(labels), None, L.ptr(ws),
                                           ws.numel(), L.cur_stream()), "vame_kmeans_assign")
        return labels

    @property
    def cluster_centers_(self):
        return self.result.cluster_centers_.cpu().numpy()

    @property
    def labels_(self):
        return self.result.labels_.cpu().numpy()

    @property
    def inertia_(self):
        return self.result.inertia_

    @property
    def n_iter_(self):
        return self.result.n_iter_
