"""Training-set preparation on the device: the drop-in for ``vame/model/create_training.py``.

Mirrors the reference's function names and arguments (``traindata_aligned``, ``traindata_fixed``, ``create_trainset``;
create_training.py:94-186, :189-246, :249-291) and its on-disk formats: reads ``data/<file>/<file>-PE-seq.npy`` ((F, N)
float64), writes ``data/train/train_seq.npy``, ``data/train/test_seq.npy`` and ``data/<file>/<file>-PE-seq-clean.npy``
(float64), so the reference's ``train_model`` / ``pose_segmentation`` read them unchanged.

All passes over the data - z-score, IQR (device radix sort), outlier marking, both interpolation variants, marker standard
deviations, Savitzky-Golay smoothing - are CUDA kernels behind the C-ABI (``vame_trainset_*``, csrc/trainset.cu), in float64
like the reference.  Host work is limited to file I/O, the choice of the two anchor markers from F standard deviations and the
Savitzky-Golay tables (a window-sized least-squares problem).  ``check_parameter=True`` (matplotlib inspection plots) is not part of
the path.  No CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib as L
from ._lib import VameB200Error


def savgol_tables(window_length, polyorder):
    """Interior FIR coefficients and the edge matrices of ``scipy.signal.savgol_filter(mode='interp')``: the first / last
    ``window_length // 2`` outputs are the least-squares polynomial through the first / last ``window_length`` samples."""
    window_length, polyorder = int(window_length), int(polyorder)
    if window_length % 2 != 1 or window_length < 3 or polyorder >= window_length:
        raise ValueError("savgol_length must be odd, >= 3 and larger than savgol_order")
    half = window_length // 2
    pos = np.arange(-half, window_length - half, dtype=np.float64)
    A = pos[:, None] ** np.arange(polyorder + 1)[None, :]
    coeffs = np.linalg.pinv(A)[0]
    t = (np.arange(window_length, dtype=np.float64) - half) / max(half, 1)      # centred / scaled: same hat matrix, better conditioned
    V = t[:, None] ** np.arange(polyorder + 1)[None, :]
    hat = V @ np.linalg.pinv(V)
    return coeffs, np.ascontiguousarray(hat[:half]), np.ascontiguousarray(hat[window_length - half:])


def _dev(device):
    if not torch.cuda.is_available():
        raise VameB200Error("vame_b200.create_training needs a CUDA device (no CPU fallback)")
    return torch.device(device if device is not None else "cuda")


def zscore_clean(data, robust=True, iqr_factor=4, fixed=True, device=None):
    """Per-file stage (create_training.py:106-137 / :204-232).  data: (F, N) float64 numpy array or CUDA tensor.
    Returns (X_z as an (F, N) float64 CUDA tensor, stats dict with mean / std / iqr / outliers / unfilled)."""
    dev = _dev(device)
    lib = L.lib()
    x = torch.as_tensor(data, dtype=torch.float64).to(dev).contiguous()
    F, N = x.shape
    out = torch.empty_like(x)
    stats = torch.zeros(5, dtype=torch.float64, device=dev)
    nb = lib.vame_trainset_workspace_bytes(int(N), int(F))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.vame_trainset_zscore_clean(L.ptr(x), int(N), int(F), int(bool(robust)), float(iqr_factor), int(bool(fixed)),
                                               L.ptr(out), L.ptr(stats), L.ptr(ws), ws.numel(), L.cur_stream()),
                "vame_trainset_zscore_clean")
    s = stats.cpu().tolist()
    n = float(F * N)
    info = {"mean": s[0] / n, "std": float(np.sqrt(s[1] / n)), "iqr": s[2], "outliers": int(s[3]), "unfilled": int(s[4])}
    if info["unfilled"]:
        # np.interp raises on a frame (fixed) / an array (aligned) without any valid sample; the reference would stop here too
        raise ValueError("array of sample points is empty (%d entries could not be interpolated)" % info["unfilled"])
    return out, info


def row_std(x):
    """np.std(X, axis=1) of an (F, N) float64 CUDA tensor -> (F,) float64 CUDA tensor."""
    lib = L.lib()
    out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        L.check(lib.vame_trainset_row_std(L.ptr(x), int(x.shape[1]), int(x.shape[0]), L.ptr(out), L.cur_stream()), "vame_trainset_row_std")
    return out


def savgol_filter(x, window_length, polyorder):
    """scipy.signal.savgol_filter(x, window_length, polyorder) along the last axis of an (F, N) float64 CUDA tensor."""
    lib = L.lib()
    coeffs, head, tail = (torch.from_numpy(a).to(x.device) for a in savgol_tables(window_length, polyorder))
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(lib.vame_trainset_savgol(L.ptr(x), int(x.shape[1]), int(x.shape[0]), int(window_length), L.ptr(coeffs), L.ptr(head),
                                         L.ptr(tail), L.ptr(out), L.cur_stream()), "vame_trainset_savgol")
    return out


def trainset_arrays(datas, fixed, robust=True, iqr_factor=4, savgol=True, savgol_length=5, savgol_order=2, test_fraction=0.1,
                    device=None, verbose=False, names=None):
    """The arithmetic of traindata_fixed / traindata_aligned on a list of (F, N) arrays.
    Returns (z_train, z_test, [cleaned array per file]) as float64 numpy arrays of shape (F', N)."""
    parts, pos = [], [0]
    for i, d in enumerate(datas):
        if verbose:
            print("z-scoring of file %s" % (names[i] if names else i))
        xz, info = zscore_clean(d, robust, iqr_factor, fixed, device)
        if verbose and robust:
            print("IQR value: %.2f, IQR cutoff: %.2f" % (info["iqr"], iqr_factor * info["iqr"]))
        parts.append(xz)
        pos.append(pos[-1] + xz.shape[1])
    X = torch.cat(parts, dim=1)                                      # (F, N_total)
    if not fixed:
        # the two markers with the smallest standard deviation are the alignment anchors (create_training.py:148-171)
        sd = row_std(X).cpu().numpy()
        order = np.sort(sd)
        if order[0] == order[1]:
            a = np.where(sd == order[0])[0]
            a1, a2 = int(a[0]), int(a[1])
        else:
            a1, a2 = int(np.where(sd == order[0])[0][0]), int(np.where(sd == order[1])[0][0])
        keep = [f for f in range(X.shape[0]) if f not in (a1, a2)]
        X = X[keep].contiguous()
    Xm = savgol_filter(X, savgol_length, savgol_order) if savgol else X
    n_test = int(Xm.shape[1] * test_fraction)
    Xh = Xm.cpu().numpy()
    return Xh[:, n_test:], Xh[:, :n_test], [Xh[:, pos[i]:pos[i + 1]] for i in range(len(datas))]


def _run(cfg, files, testfraction, savgol_filter_flag, check_parameter, fixed):
    if check_parameter:
        raise NotImplementedError("check_parameter=True (matplotlib inspection plots) is outside the B200 path; "
                                  "use the reference's create_trainset for it")
    datas = [np.load(os.path.join(cfg["project_path"], "data", f, f + "-PE-seq.npy")) for f in files]
    z_train, z_test, cleans = trainset_arrays(datas, fixed, cfg["robust"] == True, cfg["iqr_factor"], bool(savgol_filter_flag),  # noqa: E712
                                              cfg["savgol_length"], cfg["savgol_order"], testfraction, verbose=True, names=files)
    np.save(os.path.join(cfg["project_path"], "data", "train", "train_seq.npy"), z_train)
    np.save(os.path.join(cfg["project_path"], "data", "train", "test_seq.npy"), z_test)
    for f, c in zip(files, cleans):
        np.save(os.path.join(cfg["project_path"], "data", f, f + "-PE-seq-clean.npy"), c)
    print("Lenght of train data: %d" % z_train.shape[1])
    print("Lenght of test data: %d" % z_test.shape[1])


def traindata_aligned(cfg, files, testfraction, num_features, savgol_filter, check_parameter):
    """create_training.py:94-186."""
    _run(cfg, files, testfraction, savgol_filter, check_parameter, fixed=False)


def traindata_fixed(cfg, files, testfraction, num_features, savgol_filter, check_parameter):
    """create_training.py:189-246."""
    _run(cfg, files, testfraction, savgol_filter, check_parameter, fixed=True)


def _read_config(path):
    import yaml
    with open(path) as fh:
        return yaml.safe_load(fh)


def create_trainset(config, check_parameter=False):
    """create_training.py:249-291 (the interactive file selection of all_data == 'No' is kept)."""
    cfg = _read_config(config)
    os.makedirs(os.path.join(cfg["project_path"], "data", "train"), exist_ok=True)
    files = []
    if cfg["all_data"] == "No":
        for file in cfg["video_sets"]:
            if input("Do you want to train on " + file + "? yes/no: ") == "yes":
                files.append(file)
    else:
        files = list(cfg["video_sets"])
    print("Creating training dataset...")
    if cfg["robust"] == True:   # noqa: E712
        print("Using robust setting to eliminate outliers! IQR factor: %d" % cfg["iqr_factor"])
    if cfg["egocentric_data"] == False:   # noqa: E712
        print("Creating trainset from the vame.egocentrical_alignment() output ")
        traindata_aligned(cfg, files, cfg["test_fraction"], cfg["num_features"], cfg["savgol_filter"], check_parameter)
    else:
        print("Creating trainset from the vame.csv_to_numpy() output ")
        traindata_fixed(cfg, files, cfg["test_fraction"], cfg["num_features"], cfg["savgol_filter"], check_parameter)
    if not check_parameter:
        print("A training and test set has been created. Next step: vame.train_model()")
