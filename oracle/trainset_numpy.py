"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's training-set preparation
(vame/model/create_training.py:94-264).  Pinned to outputs of the reference's own
traindata_fixed / traindata_aligned (tests/golden/trainset.npz, written by oracle/gen_golden.py --only-trainset).

What the reference computes per file  (create_training.py:106-137 aligned, :204-232 fixed):
    X_z = (data.T - mean(data)) / std(data)                    global mean / std over ALL entries, float64
    robust: iqr = scipy.stats.iqr(X_z); entries with |X_z| > iqr_factor * iqr become NaN, then
        fixed   (:231)  every FRAME (row of X_z) is interpolated on its own with np.interp over the marker index
        aligned (:137)  interpol(X_z) on the whole 2-D array: nan_helper's `z.nonzero()[0]` returns the FIRST-axis index of
                        y = X_z.T, i.e. the marker index, for both the NaN entries and the sample points, so np.interp sees
                        xp = marker index of every valid entry (non-decreasing, with repeats) and fp = the valid values in
                        row-major order: a NaN of marker f becomes the LAST valid time sample of marker f (np.interp returns
                        fp[j] for the last j with xp[j] <= x when x == xp[j]); a marker with no valid entry at all is
                        interpolated between the last valid sample of the previous marker and the first of the next one.
then (all files concatenated along time):
    aligned (:148-171) the two markers with the smallest standard deviation (the alignment anchors) are deleted
    savgol (:175-178)  scipy.signal.savgol_filter(X, savgol_length, savgol_order) along time, mode='interp'
    split  (:180-184)  the first int(num_frames * test_fraction) frames are the test set, the rest the train set
"""
import numpy as np


def percentile_linear(sorted_vals, q):
    """numpy's default ('linear') percentile on a sorted 1-D array, including its lerp form (numpy/lib/_function_base_impl.py:
    a + (b - a) * t, and b - (b - a) * (1 - t) where t >= 0.5)."""
    n = sorted_vals.shape[0]
    v = (n - 1) * (q / 100.0)
    lo = int(np.floor(v))
    hi = min(lo + 1, n - 1)
    t = v - lo
    a, b = sorted_vals[lo], sorted_vals[hi]
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def iqr(x):
    s = np.sort(np.asarray(x, dtype=np.float64).ravel())
    return percentile_linear(s, 75.0) - percentile_linear(s, 25.0)


def interp_rows(X):
    """fixed path: every row (frame) of X (N, F) interpolated over the marker index (np.interp: linear, clamped at the ends)."""
    X = X.copy()
    idx = np.arange(X.shape[1])
    for i in range(X.shape[0]):
        nans = np.isnan(X[i])
        if nans.any():
            X[i, nans] = np.interp(idx[nans], idx[~nans], X[i, ~nans])
    return X


def interp_aligned(X):
    """aligned path quirk (see the module docstring); X (N, F)."""
    y = X.T.copy()                                   # (F, N)
    F = y.shape[0]
    valid = ~np.isnan(y)
    has = valid.any(axis=1)
    first = np.full(F, np.nan)
    last = np.full(F, np.nan)
    for f in range(F):
        if has[f]:
            v = y[f, valid[f]]
            first[f], last[f] = v[0], v[-1]
    feats = np.nonzero(has)[0]
    for f in range(F):
        if valid[f].all():
            continue
        if has[f]:
            fill = last[f]
        else:
            lo = feats[feats < f]
            hi = feats[feats > f]
            if lo.size == 0:
                fill = first[hi[0]]                  # left of all sample points: fp[0]
            elif hi.size == 0:
                fill = last[lo[-1]]                  # right of all sample points: fp[-1]
            else:
                a, b = lo[-1], hi[0]
                slope = (first[b] - last[a]) / (b - a)
                fill = slope * (f - a) + last[a]
        y[f, ~valid[f]] = fill
    return y.T


def savgol_tables(window_length, polyorder):
    """Interior FIR coefficients and the edge matrices of scipy.signal.savgol_filter(mode='interp'): the first / last
    window_length // 2 outputs are the least-squares polynomial through the first / last window_length samples, evaluated there."""
    half = window_length // 2
    pos = np.arange(-half, window_length - half, dtype=np.float64)
    A = pos[:, None] ** np.arange(polyorder + 1)[None, :]
    coeffs = np.linalg.pinv(A)[0]                    # value of the fitted polynomial at the window centre; symmetric
    t = (np.arange(window_length, dtype=np.float64) - half) / max(half, 1)      # centred / scaled: same hat matrix, better conditioned
    V = t[:, None] ** np.arange(polyorder + 1)[None, :]
    hat = V @ np.linalg.pinv(V)                      # fitted values at the window's own positions
    return coeffs, hat[:half], hat[window_length - half:]


def savgol_rows(X, window_length, polyorder):
    """X (F, N) filtered along time."""
    coeffs, head, tail = savgol_tables(window_length, polyorder)
    half = window_length // 2
    N = X.shape[1]
    out = np.empty_like(X)
    for f in range(X.shape[0]):
        out[f] = np.convolve(X[f], coeffs[::-1], mode="same")
    out[:, :half] = X[:, :window_length] @ head.T
    out[:, N - half:] = X[:, N - window_length:] @ tail.T
    return out


def zscore_clean(data, robust, iqr_factor, fixed):
    """data (F, N) float64 -> X_z (N, F) after the optional outlier removal."""
    Xz = (data.T - np.mean(data)) / np.std(data)
    if robust:
        cut = iqr_factor * iqr(Xz)
        Xz = np.where((Xz > cut) | (Xz < -cut), np.nan, Xz)
        Xz = interp_rows(Xz) if fixed else interp_aligned(Xz)
    return Xz


def traindata(datas, fixed, robust=True, iqr_factor=4, savgol=True, savgol_length=5, savgol_order=2, test_fraction=0.1):
    """-> (z_train, z_test, [per-file cleaned arrays]) all (F', N) float64."""
    parts = [zscore_clean(d, robust, iqr_factor, fixed) for d in datas]
    pos = np.cumsum([0] + [p.shape[0] for p in parts])
    X = np.concatenate(parts, axis=0)                # (N_total, F)
    if not fixed:
        sd = np.std(X.T, axis=1)
        order = np.sort(sd)
        if order[0] == order[1]:
            a = np.where(sd == order[0])[0]
            a1, a2 = a[0], a[1]
        else:
            a1, a2 = int(np.where(sd == order[0])[0][0]), int(np.where(sd == order[1])[0][0])
        hi, lo = max(a1, a2), min(a1, a2)
        X = np.delete(np.delete(X, hi, 1), lo, 1)
    X = X.T
    Xm = savgol_rows(X, savgol_length, savgol_order) if savgol else X
    n_test = int(Xm.shape[1] * test_fraction)
    return Xm[:, n_test:], Xm[:, :n_test], [Xm[:, pos[i]:pos[i + 1]] for i in range(len(datas))]
